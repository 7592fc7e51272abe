"""GPU parity: fb_linearize (batched trim + finite-difference Jacobians, BASELINE.json configs[4] at test size) against the
oracle's linearize on the same trim points. Forward differences with eps ~ 1.5e-8 amplify rounding differences between two
correct right-hand sides by 1/eps, so the Jacobians agree to ~1e-6 of each row's scale, not to 1e-9; x0, u0, y0 (no
differencing) agree to 1e-9."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_envelope_sweep_linearize_matches_oracle(fb, orc, variant):
    EAS, h, flaps = np.meshgrid(np.linspace(28, 55, 6), np.linspace(50, 3050, 4), [0.0, 1.0], indexing="ij")
    EAS, h, flaps = EAS.ravel(), h.ravel(), flaps.ravel()
    keep = ~((flaps == 1.0) & (EAS > 42))          # full flap is power-limited at high speed
    EAS, h, flaps = EAS[keep], h[keep], flaps[keep]
    n = len(EAS)
    rng = np.random.default_rng(5)
    m_pilot = rng.uniform(50, 100, n)
    ac = [fb.Cessna172Sv0, fb.Cessna172Xv0, fb.Cessna172Xv1][variant](fb.NED())
    sim = fb.BatchSimulation(fb.SimpleWorld(ac), n=n, dt=0.01, t_end=1.0)
    ok, ts, cost = sim.init(fb.C172.TrimParameters(EAS=EAS, h_e=h, flaps=flaps, m_pilot=m_pilot))
    assert ok.all()
    x_before = sim.x
    g = sim.linearize()
    assert np.array_equal(sim.x, x_before)                                    # the ensemble is left at its trim
    worst = 0.0
    for i in range(0, n, 3):
        o = orc.Sim(variant=variant, kin=2, dt=0.01)
        ok_o, _, _ = o.init_trim(EAS=EAS[i], h_e=h[i], flaps=flaps[i], m_pilot=m_pilot[i])
        assert ok_o
        r = o.linearize()
        for k in ("x0", "u0", "y0"):
            assert np.all(np.abs(g[k][i] - r[k]) <= 1e-9 * np.maximum(np.abs(r[k]), 1e-2)), (k, i)
        assert np.abs(g["xdot0"][i] - r["xdot0"]).max() < 1e-7
        for k in ("A", "B", "C", "D"):
            scale = np.maximum(np.abs(r[k]).max(axis=1, keepdims=True), 1.0)
            err = np.abs(g[k][i] - r[k]) / scale
            worst = max(worst, float(err.max()))
            assert err.max() < 2e-5, (k, i, np.unravel_index(err.argmax(), err.shape), float(err.max()))
        # the dynamics the Jacobians describe are the same: eigenvalues of A agree
        eg, eo = np.sort_complex(np.linalg.eigvals(g["A"][i])), np.sort_complex(np.linalg.eigvals(r["A"]))
        assert np.abs(eg - eo).max() < 1e-4 * max(1.0, np.abs(eo).max())
    print("linearize worst row-scaled Jacobian error", worst)
    sim.close()


def test_linearize_needs_ned(fb):
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=4, dt=0.01, t_end=1.0)
    sim.init(fb.C172.TrimParameters())
    with pytest.raises(fb.FlightBatchError, match="NED"):
        sim.linearize()
    sim.close()


def test_cfg5_lattice_of_1296_points(fb, orc):
    """BASELINE.json configs[4] at 1,296 lattice points (the tools/bench_configs.py cfg5 lattice, 6 values per axis): every 9th
    point against the oracle — trim state and operating point <= 1e-9, Jacobians to forward-difference noise, infeasible
    lattice corners flagged identically."""
    m = 6
    EAS, h, fl, mass = np.meshgrid(np.linspace(25, 55, m), np.linspace(50, 3050, m), np.linspace(0, 1, m), np.linspace(0, 100, m), indexing="ij")
    EAS, h, fl, mass = (a.ravel() for a in (EAS, h, fl, mass))
    fl = np.where(EAS > 42, 0.0, fl)
    n = EAS.size
    assert n == 1296
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.NED())), n=n, dt=0.01, t_end=1.0)
    ok, ts, cost = sim.init(fb.C172.TrimParameters(EAS=EAS, h_e=h, flaps=fl, m_pilot=mass, m_copilot=mass))
    g = sim.linearize()
    worst_ts, worst_j, n_fail = 0.0, 0.0, 0
    for i in range(0, n, 9):
        o = orc.Sim(variant=0, kin=2, dt=0.01)
        ok_o, ts_o, _ = o.init_trim(EAS=EAS[i], h_e=h[i], flaps=fl[i], m_pilot=mass[i], m_copilot=mass[i])
        assert ok_o == bool(ok[i]), (i, EAS[i], h[i], fl[i])
        if not ok_o:
            n_fail += 1
            continue
        e = np.abs(ts[:, i] - ts_o) / np.maximum(np.abs(ts_o), 1e-2)
        worst_ts = max(worst_ts, float(e.max()))
        assert e.max() < 1e-9, (i, e)
        r = o.linearize()
        for k in ("x0", "u0", "y0"):
            assert np.all(np.abs(g[k][i] - r[k]) <= 1e-9 * np.maximum(np.abs(r[k]), 1e-2)), (k, i)
        for k in ("A", "B", "C", "D"):
            err = np.abs(g[k][i] - r[k]) / np.maximum(np.abs(r[k]).max(axis=1, keepdims=True), 1.0)
            worst_j = max(worst_j, float(err.max()))
            assert err.max() < 2e-5, (k, i, float(err.max()))
    print(f"cfg5 lattice: {n} points, {n_fail} of the sampled {len(range(0, n, 9))} infeasible (slow, low, flapless corners); worst trim-state error {worst_ts:.2e}, "
          f"worst row-scaled Jacobian error {worst_j:.2e}")
    sim.close()
