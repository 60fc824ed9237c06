"""Round-2 experiment batch C: restructured red-black kernel, prefetch distances."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
from exp_r2a import run
for pf in ("0", "4", "8", "12"):
    run("c3 rb pf" + pf, [4096, 4096, 100, 9.8, 1, 10, "none"], {"TFS_RB_PF": pf})
for pf in ("0", "8"):
    run("c4 rb pf" + pf, [16384, 16384, 50, 0.0, 1, 3, "lattice"], {"TFS_RB_PF": pf})
    run("c5 rb pf" + pf, [32768, 32768, 200, 0.0, 1, 2, "random"], {"TFS_RB_PF": pf})
