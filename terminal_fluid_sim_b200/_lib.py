"""ctypes binding of libtfs_cuda.so -- the same symbols a Rust `tfs-sys` crate would bind
(include/tfs.h).  Fails loudly if the library is missing: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtfs_cuda.so")

PEER_BLOB_BYTES = 512

# every symbol include/tfs.h declares (tests check the .so exports exactly these)
SYMBOLS = [
    "tfs_version", "tfs_last_error", "tfs_device_count", "tfs_default_config", "tfs_default_options",
    "tfs_create", "tfs_destroy", "tfs_resize", "tfs_restart", "tfs_set_config", "tfs_get_config",
    "tfs_set_options", "tfs_get_options", "tfs_get_size", "tfs_set_block", "tfs_upload_blocks",
    "tfs_step", "tfs_phase_gravity", "tfs_phase_project", "tfs_phase_advect", "tfs_sync",
    "tfs_read_view", "tfs_read_field", "tfs_write_field", "tfs_pressure_minmax", "tfs_render_view",
    "tfs_divergence_linf",
    "tfs_scene_disc", "tfs_scene_lattice", "tfs_scene_random", "tfs_fill_random_velocity",
    "tfs_host_alloc", "tfs_host_free", "tfs_timer_begin", "tfs_timer_end", "tfs_set_profiling",
    "tfs_get_phase_ms", "tfs_launch_count", "tfs_selftest_division", "tfs_peer_export", "tfs_peer_attach",
    "tfs_slab_bounds", "tfs_get_slab", "tfs_get_last_projection", "tfs_checksum", "tfs_stream_idle",
]


class TfsConfig(C.Structure):
    _fields_ = [("gravity", C.c_float), ("wind_speed", C.c_float),
                ("smoke_size", C.c_float), ("density", C.c_float)]


class TfsOptions(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("omega", C.c_float), ("mode", C.c_int32),
                ("device", C.c_int32), ("kernel", C.c_int32), ("temporal_block", C.c_int32),
                ("rank", C.c_int32), ("world", C.c_int32)]


class TfsProjectionInfo(C.Structure):
    _fields_ = [("kernel", C.c_int32), ("kg", C.c_int32), ("nw", C.c_int32), ("passes", C.c_int32)]


class TfsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tfs error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """dlopen the in-tree library and declare argument types."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m terminal_fluid_sim_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, sz, f32, i32, u64 = C.c_void_p, C.c_size_t, C.c_float, C.c_int, C.c_uint64
    pf32 = C.POINTER(C.c_float)
    L.tfs_version.restype = C.c_char_p
    L.tfs_last_error.restype = C.c_char_p
    L.tfs_device_count.argtypes = [C.POINTER(C.c_int)]
    L.tfs_default_config.argtypes = [C.POINTER(TfsConfig)]
    L.tfs_default_config.restype = None
    L.tfs_default_options.argtypes = [C.POINTER(TfsOptions)]
    L.tfs_default_options.restype = None
    L.tfs_create.argtypes = [sz, sz, C.POINTER(TfsConfig), C.POINTER(TfsOptions), C.POINTER(vp)]
    L.tfs_destroy.argtypes = [vp]
    L.tfs_resize.argtypes = [vp, sz, sz]
    L.tfs_restart.argtypes = [vp]
    L.tfs_set_config.argtypes = [vp, C.POINTER(TfsConfig)]
    L.tfs_get_config.argtypes = [vp, C.POINTER(TfsConfig)]
    L.tfs_set_options.argtypes = [vp, C.POINTER(TfsOptions)]
    L.tfs_get_options.argtypes = [vp, C.POINTER(TfsOptions)]
    L.tfs_get_size.argtypes = [vp, C.POINTER(sz), C.POINTER(sz)]
    L.tfs_set_block.argtypes = [vp, sz, sz, i32]
    L.tfs_upload_blocks.argtypes = [vp, vp, sz, sz]
    L.tfs_step.argtypes = [vp, f32]
    L.tfs_phase_gravity.argtypes = [vp, f32]
    L.tfs_phase_project.argtypes = [vp, f32]
    L.tfs_phase_advect.argtypes = [vp, f32]
    L.tfs_sync.argtypes = [vp]
    L.tfs_read_view.argtypes = [vp, i32, sz, sz, sz, sz, vp]
    L.tfs_read_field.argtypes = [vp, i32, vp, sz]
    L.tfs_write_field.argtypes = [vp, i32, vp, sz]
    L.tfs_pressure_minmax.argtypes = [vp, pf32, pf32]
    L.tfs_render_view.argtypes = [vp, sz, sz, sz, pf32, vp, pf32, pf32]
    L.tfs_divergence_linf.argtypes = [vp, pf32]
    L.tfs_scene_disc.argtypes = [vp, C.c_long, C.c_long, C.c_long]
    L.tfs_scene_lattice.argtypes = [vp, C.c_long, C.c_long]
    L.tfs_scene_random.argtypes = [vp, u64, u64]
    L.tfs_fill_random_velocity.argtypes = [vp, u64, f32]
    L.tfs_host_alloc.argtypes = [C.POINTER(vp), sz]
    L.tfs_host_free.argtypes = [vp]
    L.tfs_timer_begin.argtypes = [vp]
    L.tfs_timer_end.argtypes = [vp, pf32]
    L.tfs_set_profiling.argtypes = [vp, i32]
    L.tfs_get_phase_ms.argtypes = [vp, pf32, pf32, pf32, C.POINTER(u64)]
    L.tfs_launch_count.argtypes = [vp, C.POINTER(u64)]
    L.tfs_selftest_division.argtypes = [i32, C.POINTER(u64)]
    L.tfs_peer_export.argtypes = [vp, vp]
    L.tfs_peer_attach.argtypes = [vp, vp, vp]
    L.tfs_slab_bounds.argtypes = [sz, i32, i32, C.POINTER(sz), C.POINTER(sz)]
    L.tfs_get_slab.argtypes = [vp, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)]
    L.tfs_get_last_projection.argtypes = [vp, C.POINTER(TfsProjectionInfo)]
    L.tfs_checksum.argtypes = [vp, C.POINTER(u64)]
    L.tfs_stream_idle.argtypes = [vp, C.POINTER(C.c_int)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise TfsError(rc, load().tfs_last_error().decode(errors="replace"))
