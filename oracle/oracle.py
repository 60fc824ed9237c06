"""ctypes harness around oracle/liboracle.so (fluid_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of fluid_oracle.c.  The product package
(terminal_fluid_sim_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


class OrcConfig(C.Structure):
    _fields_ = [("gravity", C.c_float), ("wind_speed", C.c_float),
                ("smoke_size", C.c_float), ("density", C.c_float)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "fluid_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, sz, f32, i32, u64 = C.c_void_p, C.c_size_t, C.c_float, C.c_int, C.c_uint64
        L.orc_create.restype = vp
        L.orc_create.argtypes = [sz, sz, C.POINTER(OrcConfig)]
        L.orc_destroy.argtypes = [vp]
        L.orc_resize.argtypes = [vp, sz, sz]
        L.orc_restart.argtypes = [vp]
        L.orc_set_config.argtypes = [vp, C.POINTER(OrcConfig)]
        L.orc_set_block.argtypes = [vp, sz, sz, i32]
        L.orc_add_gravity.argtypes = [vp, f32]
        L.orc_make_incompressible.argtypes = [vp, f32, i32, f32]
        L.orc_make_incompressible_redblack.argtypes = [vp, f32, i32, f32]
        L.orc_move_velocity.argtypes = [vp, f32]
        L.orc_step.argtypes = [vp, f32, i32, f32, i32]
        L.orc_divergence_linf.restype = f32
        L.orc_divergence_linf.argtypes = [vp]
        L.orc_cell_flags.argtypes = [vp, vp]
        L.orc_render_view.argtypes = [vp, sz, sz, vp, C.POINTER(f32), C.POINTER(f32)]
        L.orc_scene_disc.argtypes = [vp, C.c_long, C.c_long, C.c_long]
        L.orc_scene_lattice.argtypes = [vp, C.c_long, C.c_long]
        L.orc_scene_random.argtypes = [vp, u64, u64]
        L.orc_fill_random_velocity.argtypes = [vp, u64, f32]
        for name in ("orc_u", "orc_v", "orc_p", "orc_s", "orc_m"):
            getattr(L, name).restype = vp
            getattr(L, name).argtypes = [vp]
        L.orc_width.restype = sz
        L.orc_width.argtypes = [vp]
        L.orc_height.restype = sz
        L.orc_height.argtypes = [vp]
        _lib = L
    return _lib


class OracleSim:
    """Mirrors FluidSim (simulator.rs:28-340) with dt / iterations / omega injected."""

    def __init__(self, width, height, gravity=0.0, wind_speed=50.0, smoke_size=0.25, density=1000.0,
                 iterations=50, omega=1.9, mode=0):
        self.L = lib()
        cfg = OrcConfig(gravity, wind_speed, smoke_size, density)
        self.h = self.L.orc_create(width, height, C.byref(cfg))
        if not self.h:
            raise ValueError("width and height must be >= 1")
        self.iterations, self.omega, self.mode = iterations, omega, mode

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    # -- geometry ---------------------------------------------------------
    @property
    def width(self):
        return self.L.orc_width(self.h)

    @property
    def height(self):
        return self.L.orc_height(self.h)

    def get_size(self):
        return (self.width, self.height)

    def _view(self, name, dtype):
        n = self.width * self.height
        ptr = getattr(self.L, name)(self.h)
        ctype = C.c_float if dtype == np.float32 else C.c_uint8
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,))

    # live views into the oracle's memory (invalid after step/resize: re-fetch)
    @property
    def u(self):
        return self._view("orc_u", np.float32)

    @property
    def v(self):
        return self._view("orc_v", np.float32)

    @property
    def p(self):
        return self._view("orc_p", np.float32)

    @property
    def s(self):
        return self._view("orc_s", np.float32)

    @property
    def m(self):
        return self._view("orc_m", np.uint8)

    # -- reference API ----------------------------------------------------
    def resize(self, w, h):
        if self.L.orc_resize(self.h, w, h):
            raise ValueError("bad size")

    def restart_sim(self):
        self.L.orc_restart(self.h)

    def set_config(self, gravity, wind_speed, smoke_size, density):
        cfg = OrcConfig(gravity, wind_speed, smoke_size, density)
        self.L.orc_set_config(self.h, C.byref(cfg))

    def set_block(self, x, y):
        if self.L.orc_set_block(self.h, x, y, 1):
            raise IndexError("index out of bounds")

    def unset_block(self, x, y):
        if self.L.orc_set_block(self.h, x, y, 0):
            raise IndexError("index out of bounds")

    def step(self, dt):
        self.L.orc_step(self.h, dt, self.iterations, self.omega, self.mode)

    def add_gravity(self, dt):
        self.L.orc_add_gravity(self.h, dt)

    def make_incompressible(self, dt):
        if self.mode == 0:
            self.L.orc_make_incompressible(self.h, dt, self.iterations, self.omega)
        else:
            self.L.orc_make_incompressible_redblack(self.h, dt, self.iterations, self.omega)

    def move_velocity(self, dt):
        self.L.orc_move_velocity(self.h, dt)

    def divergence_linf(self):
        return float(self.L.orc_divergence_linf(self.h))

    def cell_flags(self):
        out = np.empty(self.width * self.height, np.uint8)
        self.L.orc_cell_flags(self.h, out.ctypes.data)
        return out

    def render_view(self, area_w, area_h):
        """render_sim (sim_renderer.rs:23-83): (bytes[area_h, area_w, 2, 4], min_p, max_p)"""
        out = np.empty((area_h, area_w, 2, 4), np.uint8)
        mn, mx = C.c_float(), C.c_float()
        if self.L.orc_render_view(self.h, area_w, area_h, out.ctypes.data, C.byref(mn), C.byref(mx)):
            raise IndexError("attempt to subtract with overflow")
        return out, mn.value, mx.value

    # -- scenes -----------------------------------------------------------
    def scene_disc(self, cx, cy, r):
        self.L.orc_scene_disc(self.h, cx, cy, r)

    def scene_lattice(self, pitch, r):
        self.L.orc_scene_lattice(self.h, pitch, r)

    def scene_random(self, seed, frac):
        self.L.orc_scene_random(self.h, seed, frac_to_threshold(frac))

    def fill_random_velocity(self, seed, amp):
        self.L.orc_fill_random_velocity(self.h, seed, amp)


def frac_to_threshold(frac: float) -> int:
    """threshold = floor(frac * 2^64) computed in exact integer arithmetic from a decimal
    fraction with <= 6 digits (0.02 -> 2 * 2^64 // 100)."""
    num = round(frac * 1_000_000)
    return (num << 64) // 1_000_000
