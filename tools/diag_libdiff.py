"""Diagnostic: derivative and short-horizon state of the same ensemble under two builds of the library.
usage: diag_libdiff.py run <out.npz>   (uses FLIGHTBATCH_LIB)   |   diag_libdiff.py cmp a.npz b.npz"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if sys.argv[1] == "run":
    import flightbatch as fb
    n = 256
    rng = np.random.default_rng(5)
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=0.01, t_end=100.0)
    lat = np.deg2rad(rng.uniform(-60, 60, n)); lon = rng.uniform(-np.pi, np.pi, n)
    ok, ts, cost = sim.init(fb.C172.TrimParameters(EAS=rng.uniform(30, 55, n), h_e=rng.uniform(100, 3000, n), lat=lat, lon=lon, psi_nb=rng.uniform(-3, 3, n),
                                                   m_pilot=rng.uniform(0, 100, n), m_baggage=rng.uniform(0, 100, n)))
    x0 = sim.x.copy(); xd0 = sim.xdot.copy()
    sim.set_input("elevator", sim.get_input("elevator") + 0.05)
    sim.step(0.01); x1 = sim.x.copy()
    sim.step(0.99); x100 = sim.x.copy()
    sim.step(9.0); x1000 = sim.x.copy()
    np.savez(sys.argv[2], x0=x0, xd0=xd0, x1=x1, x100=x100, x1000=x1000)
else:
    a = np.load(sys.argv[2]); b = np.load(sys.argv[3])
    for k in ("x0", "xd0", "x1", "x100", "x1000"):
        d = np.abs(a[k] - b[k]); s = np.maximum(np.abs(a[k]), 1e-3)
        r = d / s
        print(k, "max abs %.3e  max scaled %.3e  at row %d" % (d.max(), r.max(), np.unravel_index(r.argmax(), r.shape)[0]), " per-row max scaled:", " ".join("%.1e" % v for v in r.max(axis=1)))
