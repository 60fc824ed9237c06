import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
from exp_r2a import run
for v in ("0", "5", "6"):
    run("c3 advect variant " + v, [4096, 4096, 4, 9.8, 0, 20, "none"], {"TFS_ADV_VARIANT": v})
    run("8192 advect variant " + v, [8192, 8192, 4, 9.8, 0, 10, "none"], {"TFS_ADV_VARIANT": v})
for wpc in ("4", "2", "1"):
    run("c3 rb wpc" + wpc, [4096, 4096, 100, 9.8, 1, 10, "none"], {"TFS_RB_WPC": wpc})
    run("c4 rb wpc" + wpc, [16384, 16384, 50, 0.0, 1, 3, "lattice"], {"TFS_RB_WPC": wpc})
run("c3 rb wpc1 lx64", [4096, 4096, 100, 9.8, 1, 10, "none"], {"TFS_RB_WPC": "1", "TFS_RB_CHUNK": "64"})
run("c3 rb wpc1 lx96", [4096, 4096, 100, 9.8, 1, 10, "none"], {"TFS_RB_WPC": "1", "TFS_RB_CHUNK": "96"})
