"""Diagnostic (GPU box): RHS of the CUDA path vs the oracle at oracle-trimmed states, per component."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
import flightbatch as fb
from oracle import orc
from test_gpu_parity import cfg2_params, COPY_FIELDS, SEED
np.set_printoptions(precision=3, linewidth=250)
variant, kin = int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = 8
rng = np.random.default_rng(SEED)
p = cfg2_params(n, rng)
ac = [fb.Cessna172Sv0, fb.Cessna172Xv0][variant]([fb.WA, None, fb.NED][kin]())
sim = fb.BatchSimulation(fb.SimpleWorld(ac), n=n, dt=0.01, t_end=100.0)
ors = []
X = np.zeros((sim.nx, n))
for i in range(n):
    o = orc.Sim(variant, kin, 0.01)
    ok, ts, c = o.init_trim(**{k: float(v[i]) for k, v in p.items()})
    X[:, i] = o.x
    ors.append(o)
for f in COPY_FIELDS:
    sim.set_input(f, np.array([o.get_input(fb.INPUT_NAMES[f]) for o in ors]))
sim.init_state(X, stall=np.zeros(n, np.uint8), eng_state=np.full(n, 2, np.uint8))
xd = sim.xdot
for i in range(3):
    print("gpu   ", xd[:, i])
    print("oracle", ors[i].xdot)
    print("diff  ", xd[:, i] - ors[i].xdot)
