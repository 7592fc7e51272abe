"""Diagnostic (GPU box): magnitude of RHS and stepping differences GPU vs oracle."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
import flightbatch as fb
from oracle import orc
from test_gpu_parity import cfg2_params, oracle_from_gpu, SEED
np.set_printoptions(precision=2, linewidth=250)
variant, kin = int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = 32
rng = np.random.default_rng(SEED + 1)
p = cfg2_params(n, rng)
ac = [fb.Cessna172Sv0, fb.Cessna172Xv0][variant]([fb.WA, None, fb.NED][kin]())
sim = fb.BatchSimulation(fb.SimpleWorld(ac), n=n, dt=0.01, t_end=100.0, steps_per_launch=50)
sim.init(fb.C172.TrimParameters(**p))
ele_f = fb.U["FB_U_ELEVATOR"] if variant == 0 else fb.U["FB_U_ACT_ELEVATOR"]
sim.set_input(ele_f, sim.get_input(ele_f) + 0.1)
idx = [0, 8, 16, 24]
ors = oracle_from_gpu(fb, orc, sim, variant, kin, 0.01, idx)
xdg = sim.xdot
for i, o in zip(idx, ors):
    d = xdg[:, i] - o.xdot
    print("rhs abs diff", i, d)
    print("rhs oracle   ", i, o.xdot)
done = 0
for nsteps in [1, 9, 90, 900]:
    sim.step(nsteps * 0.01); done += nsteps
    xg = sim.x
    for i, o in zip(idx[:2], ors[:2]):
        o.step(nsteps)
        print(f"steps {done:5d} member {i}: abs diff", xg[:, i] - o.x)
print("state", ors[0].x)
print("---- RHS parity at the evolved (off-trim) state")
ors2 = oracle_from_gpu(fb, orc, sim, variant, kin, 0.01, idx)
xdg = sim.xdot
for i, o in zip(idx, ors2):
    print("rhs abs diff", i, xdg[:, i] - o.xdot)
    print("rhs oracle  ", i, o.xdot)
