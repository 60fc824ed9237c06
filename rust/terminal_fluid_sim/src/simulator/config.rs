// UNCOMPILED.  Field-for-field the reference's SimConfig (src/simulator/config.rs:1-24).
#[derive(Clone)]
pub struct SimConfig {
    pub gravity: f32,
    pub wind_speed: f32,
    pub smoke_size: f32,
    pub density: f32,
}

impl Default for SimConfig {
    fn default() -> Self {
        Self { gravity: 0.0, wind_speed: 50.0, smoke_size: 0.25, density: 1000.0 }
    }
}
