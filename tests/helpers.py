"""Shared helpers for the parity tests: golden loading, bitwise comparison, scene setup."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DT = float(np.float32(1.0 / 60.0))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    params = json.loads(bytes(z["params"]).decode())
    return z, params


def bits_equal(a, b):
    """Bitwise equality of two f32 arrays; NaNs compare equal to NaNs of any payload
    (x86 and CUDA generate different default-NaN bit patterns)."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    if a.shape != b.shape:
        return False
    ai, bi = a.view(np.uint32), b.view(np.uint32)
    return bool(np.all((ai == bi) | (np.isnan(a) & np.isnan(b))))


def first_mismatch(a, b, H):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    bad = np.nonzero((a.view(np.uint32) != b.view(np.uint32)) & ~(np.isnan(a) & np.isnan(b)))[0]
    if bad.size == 0:
        return "equal"
    k = int(bad[0])
    return f"{bad.size} mismatches, first at idx {k} (x={k // H}, y={k % H}): {a[k]!r} vs {b[k]!r}"


def sim_kwargs(params):
    return {k: params[k] for k in ("gravity", "wind_speed", "smoke_size", "density", "iterations", "omega")}
