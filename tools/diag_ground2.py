"""Diagnostic: per-state GPU-vs-oracle error of the cfg-4 test members over the first instants (uses FLIGHTBATCH_LIB)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import flightbatch as fb
from oracle import orc
import ground_util as gu
import test_gpu_ground as tg
orc.build()
n = 64; idx = list(range(0, n, 2))
sim, oracles, idx = tg.build_cfg4(fb, orc, n, idx)
T = 0
for chunk in gu.CHUNKS[:6]:
    sim.step(chunk * 0.01)
    for o in oracles: o.step(chunk)
    T += chunk
    xg = sim.x[:, idx]; xo = np.array([o.x for o in oracles]).T
    e = np.abs(xg - xo) / np.maximum(np.abs(xo), 1.0)
    memb = e.max(axis=0)
    print("t=%d steps: median member err %.2e, max %.2e; per-state max:" % (T, np.median(memb), memb.max()), " ".join("%.0e" % v for v in e.max(axis=1)))
    print("    takeoff half median %.2e, approach half median %.2e" % (np.median(memb[: len(idx) // 2]), np.median(memb[len(idx) // 2:])))
