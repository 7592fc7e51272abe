"""Diagnostic (GPU box): per-chunk GPU-vs-oracle error of the mode-sequence test."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import flightbatch as fb
from oracle import orc
import test_gpu_c172x as T
from parity_util import scaled_err, state_floors
orc.build()
n, dt, Dt = 16, 0.01, float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
rng = np.random.default_rng(T.SEED + 12)
p = T.cfg3_params(n, rng)
idx = [0]
sim, oracles = T.make_pair(fb, orc, 2, 0, n, p, idx, dt, Dt, steps_per_launch=37)
U = fb.U
h0 = sim.get_input("lon_h_ref"); eas0 = sim.get_input("lon_eas_ref"); chi0 = sim.get_input("lat_chi_ref")
ev = []
for i in range(n):
    s = 10
    for lon_mode, lat_mode in (("sas", "sas"), ("thr_q", "p_beta"), ("thr_theta", "phi_beta"), ("thr_EAS", "chi_beta"),
                               ("EAS_q", "sas"), ("EAS_theta", "p_beta"), ("EAS_clm", "chi_beta"), ("EAS_alt", "phi_beta"),
                               ("direct", "direct"), ("EAS_alt", "chi_beta")):
        ev += [(i, s, U["FB_U_LON_MODE_REQ"], T.LON[lon_mode]), (i, s + 3, U["FB_U_LAT_MODE_REQ"], T.LAT[lat_mode])]
        s += 150
m, s, f, v = zip(*ev)
sim.schedule_inputs(m, s, f, v)
o = oracles[0]
for (mi, si, fi, vi) in ev:
    if mi == 0: o.add_event(si, fi, vi)
fl = state_floors(sim.nx, 1, 0)
t = 0
for chunk in [9, 1, 1, 1, 1, 1, 1, 1, 1, 1, 20] + [10] * 140:
    sim.step(chunk * dt); t += chunk
    o.step(chunk)
    err = scaled_err(sim.x[:, 0], o.x, fl)
    outs = [sim.get_output(k)[0] for k in T.CTL_OUTPUTS]
    oo = [o.y(fb.OUTPUT_NAMES[k]) for k in T.CTL_OUTPUTS]
    flag = "" if err.max() < 1e-9 else "  <<<<"
    print(t, "err %.2e @%d" % (err.max(), err.argmax()), "gpu", ["%.9g" % a for a in outs], "orc", ["%.9g" % a for a in oo], flag)
    if err.max() > 1e-6: break
