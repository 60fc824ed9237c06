import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
from exp_r2a import run
run("c1 rb", [100, 100, 50, 0.0, 1, 50, "none"])
run("c2 rb", [1024, 1024, 40, 0.0, 1, 20, "disc"])
run("2048 rb", [2048, 2048, 40, 0.0, 1, 20, "none"])
run("c3 rb", [4096, 4096, 100, 9.8, 1, 10, "none"])
run("c3 rb again", [4096, 4096, 100, 9.8, 1, 10, "none"])
run("c5 rb fast0 pf0", [32768, 32768, 200, 0.0, 1, 2, "random"], {"TFS_RB_FAST": "0", "TFS_RB_PF": "0"})
run("c5 rb fast0 pf4", [32768, 32768, 200, 0.0, 1, 2, "random"], {"TFS_RB_FAST": "0", "TFS_RB_PF": "4"})
run("c5 rb default", [32768, 32768, 200, 0.0, 1, 2, "random"])
