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
