"""Diagnostic (GPU box): GPU-vs-oracle error growth for a member settling on its landing gear."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import flightbatch as fb
from oracle import orc
import bench
from parity_util import scaled_err, state_floors
orc.build()
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
h = 1.9 + float(bench.geoid_height(0.0, 0.0))
ac = [fb.Cessna172Sv0, fb.Cessna172Xv0, fb.Cessna172Xv1][variant](fb.WA())
sim = fb.BatchSimulation(fb.SimpleWorld(ac), n=4, dt=0.01, t_end=100.0, steps_per_launch=1)
sim.init(fb.C172.Init(h_e=h))
o = orc.Sim(variant=variant, kin=0, dt=0.01, Dt=0.01)
o.init(h_e=h)
if variant == 2:
    for k, v in (("lon_mode_req", 7), ("lat_mode_req", 2), ("lon_throttle_axis", 0.1), ("lon_elevator_axis", 0.3), ("lat_aileron_axis", 0.2), ("lat_rudder_axis", 0.4)):
        sim.set_input(k, float(v)); o.set_input(fb.INPUT_NAMES[k], v)
fl = state_floors(sim.nx, variant, 0)
xd = sim.xdot; o.rhs()
print("rhs err", np.max(np.abs(xd[:, 0] - o.xdot) / np.maximum(np.abs(o.xdot), fl)))
for c in range(sim.nx):
    print(c, "%.15e %.15e %.2e" % (xd[c, 0], o.xdot[c], abs(xd[c, 0] - o.xdot[c])))
for f in ("xi_left", "xi_right", "xi_nose", "fdmp_left", "fdmp_right", "fdmp_nose", "h_o", "h_e", "lat", "lon"):
    print(f, "%.15e %.15e" % (sim.get_output(f)[0], o.y(fb.OUTPUT_NAMES[f])))
if len(sys.argv) > 2 and sys.argv[2] == "x": sys.exit(0)
t = 0
for chunk in [1] * 30 + [10] * 20:
    sim.step(chunk * 0.01); t += chunk
    o.step(chunk)
    xg = sim.x[:, 0]
    err = scaled_err(xg, o.x, fl)
    k = int(err.argmax())
    print(t, "err %.2e @%d gpu %.12e orc %.12e" % (err.max(), k, xg[k], o.x[k]), "wow", [sim.get_output(f)[0] for f in ("wow_left", "wow_right", "wow_nose")],
          [o.y(fb.OUTPUT_NAMES[f]) for f in ("wow_left", "wow_right", "wow_nose")], "ldg x", xg[2:8], "act", xg[12:16] if variant else "")
