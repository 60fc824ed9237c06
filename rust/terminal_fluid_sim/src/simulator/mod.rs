// UNCOMPILED.  Mirrors src/simulator/mod.rs of the reference.
mod config;
mod simulator;
pub use config::SimConfig;
pub use simulator::FluidSim;
