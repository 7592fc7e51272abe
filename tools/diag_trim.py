import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
import flightbatch as fb
from oracle import orc
from test_gpu_parity import cfg2_params, SEED
np.set_printoptions(precision=4, linewidth=250)
n = 8
rng = np.random.default_rng(SEED)
p = cfg2_params(n, rng)
sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=0.01, t_end=100.0)
ok, ts, cost = sim.init(fb.C172.TrimParameters(**p))
print("ok", ok); print("cost", cost); print("ts", ts.T); print("status", sim.status())
print("x", sim.x[:, 0])
o = orc.Sim(0, 0, 0.01); print(o.init_trim(**{k: float(v[0]) for k, v in p.items()})); print(o.x)
from test_gpu_parity import oracle_from_gpu
ors = oracle_from_gpu(fb, orc, sim, 0, 0, 0.01, [0, 1])
xd = sim.xdot
for i, o2 in zip([0, 1], ors):
    print("gpu xdot   ", xd[:, i])
    print("oracle xdot", o2.xdot)
    L_dyn = 21
    v = sim.x[L_dyn + 3:L_dyn + 6, i]
    r = np.concatenate([xd[L_dyn + 3:L_dyn + 6, i] / np.linalg.norm(v), xd[L_dyn:L_dyn + 3, i], [xd[9, i] / (2700 * np.pi / 30)]])
    print("host cost from gpu xdot", np.sum(r * r), "gpu cost", cost[i])
