import sys, os, shutil
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
from exp_r2a import run
pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "terminal_fluid_sim_b200")
for variant in ("inline", "gencall", "inline"):
    shutil.copy(os.path.join(pkg, "lib_%s.so.bin" % variant), os.path.join(pkg, "libtfs_cuda.so"))
    print("=== ", variant, flush=True)
    run("c3 rb", [4096, 4096, 100, 9.8, 1, 10, "none"])
    run("c4 rb", [16384, 16384, 50, 0.0, 1, 3, "lattice"])
    run("c5 rb", [32768, 32768, 200, 0.0, 1, 2, "random"])
    run("c3 rb nonfull", [4096, 4096, 100, 9.8, 1, 10, "none"], {"TFS_RB_FAST": "2"})
