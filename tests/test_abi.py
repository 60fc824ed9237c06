"""CPU tests of the boundary: the shared library loads, exports exactly what include/tfs.h
declares, and fails loudly (no CPU fallback) when no device is present."""
import ctypes as C
import os
import re
import subprocess

import pytest

import terminal_fluid_sim_b200 as tfs
from terminal_fluid_sim_b200 import _lib, build as tfs_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    tfs_build.build()
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "tfs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tfs_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r" T (tfs_[a-z_0-9]+)", out))
    assert exported == set(header_symbols())
    for name in header_symbols():
        assert getattr(lib, name) is not None


def test_defaults_match_reference_constants(lib):
    cfg, opt = _lib.TfsConfig(), _lib.TfsOptions()
    lib.tfs_default_config(C.byref(cfg))
    lib.tfs_default_options(C.byref(opt))
    assert (cfg.gravity, cfg.wind_speed, cfg.smoke_size, cfg.density) == (0.0, 50.0, 0.25, 1000.0)  # config.rs:16-24
    assert opt.iterations == 50 and abs(opt.omega - 1.9) < 1e-7      # simulator.rs:152-153
    assert opt.mode == tfs.MODE_WAVEFRONT_EXACT and opt.world == 1
    assert b"sm_100a" in lib.tfs_version()


def test_argument_validation_without_device(lib):
    h = C.c_void_p()
    assert lib.tfs_create(0, 10, None, None, C.byref(h)) == 1          # TFS_ERR_INVALID_ARG
    assert b"width and height" in lib.tfs_last_error()
    assert lib.tfs_create(10, 10, None, None, None) == 1
    assert lib.tfs_step(None, 0.1) == 1
    assert lib.tfs_destroy(None) == 0


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(tfs.TfsError) as e:
        tfs.FluidSim(16, 16)
    assert e.value.code == 6 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package or include/ may mention it."""
    for base in ("terminal_fluid_sim_b200", "include", "rust"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".rs")):
                    text = open(os.path.join(dirpath, f), errors="replace").read()
                    assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_cpp_mirror_compiles(lib, tmp_path):
    """include/fluid_sim.hpp (the C++ host mirror of impl FluidSim) compiles and links against the .so."""
    src = tmp_path / "t.cpp"
    src.write_text('#include "fluid_sim.hpp"\nint main() { tfs_options o; tfs_default_options(&o); return o.iterations == 50 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           _lib.LIB_PATH, "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH)])
    assert subprocess.call([str(exe)]) == 0


def test_rust_bindings_cover_the_header():
    """rust/tfs-sys (uncompiled: no Rust toolchain here) declares exactly the header's symbols."""
    src = open(os.path.join(ROOT, "rust", "tfs-sys", "src", "lib.rs")).read()
    declared = sorted(set(re.findall(r"pub fn (tfs_[a-z_0-9]+)\(", src)))
    assert declared == header_symbols()
