// UNCOMPILED.  Same facade as the reference's src/lib.rs:1-3.
mod simulator;
pub use simulator::{FluidSim, SimConfig};
