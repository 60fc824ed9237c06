"""terminal_fluid_sim_b200 -- B200-native (sm_100a CUDA) per-frame simulator step of
terminal-fluid-sim behind the reference's FluidSim / SimConfig API.  The compute lives in
libtfs_cuda.so (csrc/, C ABI in include/tfs.h); importing FluidSim without the built library
raises ImportError: there is no CPU fallback."""
from .fluid_sim import (FIELD_FLAGS, FIELD_M, FIELD_P, FIELD_S, FIELD_U, FIELD_V, KERNEL_AUTO,  # noqa: F401
                        KERNEL_DIAGONAL, KERNEL_SYSTOLIC, MODE_RED_BLACK, MODE_WAVEFRONT_EXACT,
                        FluidSim, PinnedBuffer, SimConfig, SlabGroup, checksum_of_fields, combine_checksums, selftest_division,
                        slab_bounds, PROJ_NONE, PROJ_SYSTOLIC_BLOCK_RING, PROJ_SYSTOLIC_GENERIC, PROJ_DIAGONAL,
                        PROJ_REDBLACK_FUSED, PROJ_REDBLACK_TWO_LAUNCH, PROJ_SMALL_FUSED)
from ._lib import TfsError  # noqa: F401

__all__ = ["FluidSim", "SimConfig", "PinnedBuffer", "SlabGroup", "TfsError", "selftest_division", "slab_bounds"]
