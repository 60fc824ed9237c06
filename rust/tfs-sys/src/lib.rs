//! Raw bindings of include/tfs.h.  UNCOMPILED (no Rust toolchain in the build image); kept in
//! one-to-one correspondence with the header by tests/test_abi.py::test_rust_bindings_cover_the_header.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_long, c_void};

#[repr(C)]
pub struct tfs_sim {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct tfs_config {
    pub gravity: f32,
    pub wind_speed: f32,
    pub smoke_size: f32,
    pub density: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct tfs_options {
    pub iterations: i32,
    pub omega: f32,
    pub mode: i32,
    pub device: i32,
    pub kernel: i32,
    pub temporal_block: i32,
    pub rank: i32,
    pub world: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct tfs_projection_info {
    pub kernel: i32,
    pub kg: i32,
    pub nw: i32,
    pub passes: i32,
}

pub const TFS_OK: c_int = 0;
pub const TFS_ERR_OUT_OF_RANGE: c_int = 2;
pub const TFS_FIELD_P: c_int = 2;
pub const TFS_FIELD_S: c_int = 3;
pub const TFS_FIELD_M: c_int = 4;
pub const TFS_PEER_BLOB_BYTES: usize = 512;

extern "C" {
    pub fn tfs_version() -> *const c_char;
    pub fn tfs_last_error() -> *const c_char;
    pub fn tfs_device_count(count: *mut c_int) -> c_int;
    pub fn tfs_default_config(cfg: *mut tfs_config);
    pub fn tfs_default_options(opt: *mut tfs_options);
    pub fn tfs_create(width: usize, height: usize, cfg: *const tfs_config, opt: *const tfs_options, out: *mut *mut tfs_sim) -> c_int;
    pub fn tfs_destroy(sim: *mut tfs_sim) -> c_int;
    pub fn tfs_resize(sim: *mut tfs_sim, width: usize, height: usize) -> c_int;
    pub fn tfs_restart(sim: *mut tfs_sim) -> c_int;
    pub fn tfs_set_config(sim: *mut tfs_sim, cfg: *const tfs_config) -> c_int;
    pub fn tfs_get_config(sim: *const tfs_sim, cfg: *mut tfs_config) -> c_int;
    pub fn tfs_set_options(sim: *mut tfs_sim, opt: *const tfs_options) -> c_int;
    pub fn tfs_get_options(sim: *const tfs_sim, opt: *mut tfs_options) -> c_int;
    pub fn tfs_get_size(sim: *const tfs_sim, width: *mut usize, height: *mut usize) -> c_int;
    pub fn tfs_set_block(sim: *mut tfs_sim, x: usize, y: usize, value: c_int) -> c_int;
    pub fn tfs_upload_blocks(sim: *mut tfs_sim, mask: *const u8, offset: usize, n: usize) -> c_int;
    pub fn tfs_step(sim: *mut tfs_sim, dt: f32) -> c_int;
    pub fn tfs_phase_gravity(sim: *mut tfs_sim, dt: f32) -> c_int;
    pub fn tfs_phase_project(sim: *mut tfs_sim, dt: f32) -> c_int;
    pub fn tfs_phase_advect(sim: *mut tfs_sim, dt: f32) -> c_int;
    pub fn tfs_sync(sim: *mut tfs_sim) -> c_int;
    pub fn tfs_stream_idle(sim: *mut tfs_sim, idle: *mut c_int) -> c_int;
    pub fn tfs_read_view(sim: *mut tfs_sim, field: c_int, x0: usize, y0: usize, w: usize, h: usize, dst: *mut c_void) -> c_int;
    pub fn tfs_read_field(sim: *mut tfs_sim, field: c_int, dst: *mut c_void, n: usize) -> c_int;
    pub fn tfs_write_field(sim: *mut tfs_sim, field: c_int, src: *const c_void, n: usize) -> c_int;
    pub fn tfs_pressure_minmax(sim: *mut tfs_sim, min_p: *mut f32, max_p: *mut f32) -> c_int;
    pub fn tfs_render_view(sim: *mut tfs_sim, x0: usize, area_w: usize, area_h: usize, minmax_in: *const f32, dst: *mut c_void, min_p: *mut f32, max_p: *mut f32) -> c_int;
    pub fn tfs_divergence_linf(sim: *mut tfs_sim, out: *mut f32) -> c_int;
    pub fn tfs_scene_disc(sim: *mut tfs_sim, cx: c_long, cy: c_long, r: c_long) -> c_int;
    pub fn tfs_scene_lattice(sim: *mut tfs_sim, pitch: c_long, r: c_long) -> c_int;
    pub fn tfs_scene_random(sim: *mut tfs_sim, seed: u64, threshold: u64) -> c_int;
    pub fn tfs_fill_random_velocity(sim: *mut tfs_sim, seed: u64, amp: f32) -> c_int;
    pub fn tfs_host_alloc(ptr: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn tfs_host_free(ptr: *mut c_void) -> c_int;
    pub fn tfs_timer_begin(sim: *mut tfs_sim) -> c_int;
    pub fn tfs_timer_end(sim: *mut tfs_sim, elapsed_ms: *mut f32) -> c_int;
    pub fn tfs_set_profiling(sim: *mut tfs_sim, enabled: c_int) -> c_int;
    pub fn tfs_get_phase_ms(sim: *mut tfs_sim, gravity_ms: *mut f32, project_ms: *mut f32, advect_ms: *mut f32, steps: *mut u64) -> c_int;
    pub fn tfs_launch_count(sim: *const tfs_sim, launches: *mut u64) -> c_int;
    pub fn tfs_get_last_projection(sim: *const tfs_sim, info: *mut tfs_projection_info) -> c_int;
    pub fn tfs_checksum(sim: *mut tfs_sim, out: *mut u64) -> c_int;
    pub fn tfs_selftest_division(device: c_int, mismatches: *mut u64) -> c_int;
    pub fn tfs_peer_export(sim: *mut tfs_sim, blob: *mut c_void) -> c_int;
    pub fn tfs_peer_attach(sim: *mut tfs_sim, left_blob: *const c_void, right_blob: *const c_void) -> c_int;
    pub fn tfs_slab_bounds(width: usize, rank: c_int, world: c_int, c_lo: *mut usize, c_hi: *mut usize) -> c_int;
    pub fn tfs_get_slab(sim: *const tfs_sim, c_lo: *mut usize, c_hi: *mut usize, stored_lo: *mut usize, stored_hi: *mut usize) -> c_int;
}
