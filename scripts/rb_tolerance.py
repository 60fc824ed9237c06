"""Measure how far the red-black order drifts from the reference's lexicographic order on BASELINE configs 2 and 3
(CPU oracle, both orders): L-inf of u, v, smoke and the ratio of the divergence residuals.  The tolerances in
tests/test_gpu_parity.py are these numbers + 20 %.   python scripts/rb_tolerance.py [c2|c3]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.oracle import OracleSim
DT = float(np.float32(1.0 / 60.0))
which = sys.argv[1:] or ["c2", "c3"]
for name in which:
    W, K, g, steps = (1024, 40, 0.0, 20) if name == "c2" else (4096, 100, 9.8, 3)
    a, b = OracleSim(W, W, gravity=g, iterations=K, mode=0), OracleSim(W, W, gravity=g, iterations=K, mode=1)
    if name == "c2":
        a.scene_disc(256, 512, 128); b.scene_disc(256, 512, 128)
    t0 = time.time()
    for st in range(1, steps + 1):
        a.step(DT); b.step(DT)
        print(name, "step", st, "Linf u %.6g v %.6g s %.6g  residual lex %.6g rb %.6g  (%.0f s)" % (
            np.max(np.abs(a.u - b.u)), np.max(np.abs(a.v - b.v)), np.max(np.abs(a.s - b.s)), a.divergence_linf(),
            b.divergence_linf(), time.time() - t0), flush=True)
