import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
from exp_r2a import run
for fast in ("1", "0"):
    for pf in ("0", "4"):
        run("c3 rb fast%s pf%s" % (fast, pf), [4096, 4096, 100, 9.8, 1, 10, "none"], {"TFS_RB_PF": pf, "TFS_RB_FAST": fast})
