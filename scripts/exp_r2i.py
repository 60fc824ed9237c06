import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
from exp_r2a import run
for v in ("0", "10", "7", "8", "9", "11"):
    run("c3 advect variant " + v, [4096, 4096, 4, 9.8, 0, 20, "none"], {"TFS_ADV_VARIANT": v})
for v in ("0", "8"):
    run("8192 advect variant " + v, [8192, 8192, 4, 9.8, 0, 10, "none"], {"TFS_ADV_VARIANT": v})
run("c3 exact G8", [4096, 4096, 100, 9.8, 0, 10, "none"])
run("c3 exact G4", [4096, 4096, 100, 9.8, 0, 10, "none"], {"TFS_SYS_G": "4"})
run("c5 exact G4", [32768, 32768, 200, 0.0, 0, 2, "random"], {"TFS_SYS_G": "4"})
run("c4 rb", [16384, 16384, 50, 0.0, 1, 3, "lattice"])
run("c3 rb", [4096, 4096, 100, 9.8, 1, 10, "none"])
