"""GPU: batched LinearizedSS (lib/FlightPhysics/src/linearization.jl:24-111,157-192; README.md:142-204 example 3) —
the warp-per-member linear RK4 against the oracle's LinearizedSS, and the linear model beside the nonlinear one."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SEED = 20261018


@pytest.mark.parametrize("variant", [0, 1])
def test_lss_matches_oracle_and_tracks_the_nonlinear_model(fb, orc, variant):
    n, dt = 48, 0.01
    rng = np.random.default_rng(SEED + 80 + variant)
    EAS, h = rng.uniform(35, 55, n), rng.uniform(300, 2500, n)
    ac = [fb.Cessna172Sv0, fb.Cessna172Xv0][variant](fb.NED())
    sim = fb.BatchSimulation(fb.SimpleWorld(ac), n=n, dt=dt, t_end=100.0)
    ok, _, _ = sim.init(fb.C172.TrimParameters(EAS=EAS, h_e=h))
    assert ok.all()
    lin = sim.linearize()
    lss = fb.LinearizedSS(sim)                                   # Jacobians go device to device
    assert (lss.nx, lss.nu, lss.ny) == ((16, 4, 33) if variant == 0 else (20, 4, 38))
    assert np.array_equal(lss.x, lin["x0"].T) and np.array_equal(lss.u, lin["u0"].T)      # X(lss) = x0, U(lss) = u0
    assert np.allclose(lss.y, lin["y0"].T, rtol=0, atol=1e-12 * np.maximum(np.abs(lin["y0"].T), 1.0))
    # the same model from host arrays gives the same object
    lss2 = fb.LinearizedSS(model=lin, dt=dt)
    # README example 3: +0.1 elevator step (input 2 of UStateSpace: throttle, aileron, elevator, rudder), 5 s
    u = lin["u0"].T.copy(); u[2] += 0.1
    lss.u = u; lss2.u = u
    lss.step(5.0); lss2.step(5.0)
    xg = lss.x
    assert np.array_equal(xg, lss2.x)
    for i in range(0, n, 5):
        xo = orc.lss_step(lin["A"][i], lin["B"][i], lin["x0"][i], lin["u0"][i], lin["xdot0"][i], lin["x0"][i], u[:, i], dt, 500)
        err = np.abs(xg[:, i] - xo) / np.maximum(np.abs(xo), 1e-3)
        assert err.max() < 1e-10, (i, int(err.argmax()), float(err.max()))
    # outputs: y = y0 + C dx + D du
    yg = lss.y
    for i in range(0, n, 5):
        yo = lin["y0"][i] + lin["C"][i] @ (xg[:, i] - lin["x0"][i]) + lin["D"][i] @ (u[:, i] - lin["u0"][i])
        assert np.max(np.abs(yg[:, i] - yo) / np.maximum(np.abs(yo), 1e-3)) < 1e-10
    # the nonlinear ensemble under the same step stays close to its linearisation over the first seconds (pitch angle,
    # XStateSpace component 4 = theta): the check the reference's README plot makes by eye
    ele = "elevator" if variant == 0 else "act_elevator"
    sim.set_input(ele, sim.get_input(ele) + 0.1)
    sim.step(2.0)
    lss3 = fb.LinearizedSS(model=lin, dt=dt); lss3.u = u; lss3.step(2.0)
    th_nl = sim.get_output("theta"); th_lin = lss3.x[4]; th0 = lin["x0"][:, 4]
    d_nl, d_lin = th_nl - th0, th_lin - th0
    assert (np.abs(d_nl) > 0.02).all()
    assert np.median(np.abs(d_lin - d_nl) / np.abs(d_nl)) < 0.15
    for o in (lss, lss2, lss3):
        o.close()
    sim.close()
