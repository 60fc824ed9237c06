import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
from exp_r2a import run
run("c5 rb general-only kernel", [32768, 32768, 200, 0.0, 1, 2, "random"], {"TFS_RB_FAST": "2"})
run("c5 rb general-only kernel pf0", [32768, 32768, 200, 0.0, 1, 2, "random"], {"TFS_RB_FAST": "2", "TFS_RB_PF": "0"})
run("c3 rb general-only kernel", [4096, 4096, 100, 9.8, 1, 10, "none"], {"TFS_RB_FAST": "2"})
run("c3 rb default", [4096, 4096, 100, 9.8, 1, 10, "none"])
