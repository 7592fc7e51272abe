"""Diagnostic (GPU box): cfg-4 style takeoff / landing members, GPU vs oracle error by state group over time."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import flightbatch as fb
from oracle import orc
import bench
orc.build()
import test_gpu_ground as T
n = 32
sim, oracles, idx = T.build_cfg4(fb, orc, n, list(range(n)))
nx = sim.nx
t = 0
for chunk in [1, 9, 40, 50] + [100] * 24:
    sim.step(chunk * 0.01); t += chunk
    for o in oracles: o.step(chunk)
    xg = sim.x
    xo = np.array([o.x for o in oracles]).T
    d = np.abs(xg[:, idx] - xo)
    st = sim.status(); so = np.array([o.status for o in oracles])
    wow_g = np.array([sim.get_output(k)[idx] for k in ("wow_left", "wow_right", "wow_nose")])
    wow_o = np.array([[o.y(fb.OUTPUT_NAMES[k]) for o in oracles] for k in ("wow_left", "wow_right", "wow_nose")])
    stg, eng = sim.discrete()
    eo = np.array([o.discrete[1] for o in oracles])
    print(t, "aero %.1e ldg %.1e fuel %.1e eng %.1e quat %.1e h %.1e w %.1e v %.1e" % (d[0:2].max(), d[2:8].max(), d[8].max(), d[9:12].max(), d[12:20].max(), d[20].max(), d[21:24].max(), d[24:27].max()),
          "status mism", int((st[idx] != so).sum()), "wow mism", int((wow_g != wow_o).any(axis=0).sum()), "eng mism", int((eng[idx] != eo).sum()),
          "n_wow", int(wow_o.any(axis=0).sum()), "status", np.bincount(so, minlength=7).tolist(), "worst member", int(d[24:27].max(axis=0).argmax()))
