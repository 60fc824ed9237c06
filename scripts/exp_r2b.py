"""Round-2 experiment batch B: fused red-black kernel after the L2 prefetch."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
from exp_r2a import run
run("c3 rb fused", [4096, 4096, 100, 9.8, 1, 10, "none"])
run("c3 rb fused lx64", [4096, 4096, 100, 9.8, 1, 10, "none"], {"TFS_RB_CHUNK": "64"})
run("c4 rb", [16384, 16384, 50, 0.0, 1, 3, "lattice"])
run("c5 rb", [32768, 32768, 200, 0.0, 1, 3, "random"])
