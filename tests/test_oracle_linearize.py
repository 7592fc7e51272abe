"""CPU tests of the oracle's linearize(aircraft, trim) (lib/FlightPhysics/src/aircraftbase.jl:292-341, linearization.jl:55-111;
state-space vectors lib/FlightApps/src/c172/c172s/c172s.jl:269-412 and c172x/c172x.jl:332-498). The reference only asserts
`@test_nowarn linearize(...)` (test_c172s.jl:47), so these are consistency checks of the restatement: structure of the
vectors, agreement of the forward-difference Jacobians with independent central differences, and README example 3
(README.md:142-204): the linear model reproduces the nonlinear response to a small elevator step."""
import numpy as np
import pytest


@pytest.mark.parametrize("variant,dims", [(0, (16, 4, 33)), (1, (20, 4, 38)), (2, (20, 4, 38))])
def test_structure_of_the_linear_model(orc, variant, dims):
    o = orc.Sim(variant=variant, kin=2, dt=0.01)
    ok, ts, cost = o.init_trim()
    assert ok
    r = o.linearize()
    nx, nu, ny = dims
    assert r["A"].shape == (nx, nx) and r["B"].shape == (nx, nu) and r["C"].shape == (ny, nx) and r["D"].shape == (ny, nu)
    assert np.abs(r["xdot0"][[0, 1, 2, 6, 7, 8, 14]]).max() < 1e-7          # trimmed: rates, accelerations, engine speed at rest
    assert np.allclose(r["y0"][:15], r["x0"][:15], rtol=0, atol=1e-12)       # YStateSpace starts with the state block
    assert np.allclose(r["D"][-4:], np.eye(4), atol=1e-6)                    # the control inputs are passed through to y
    assert np.allclose(r["C"][:15, :15], np.eye(15), atol=1e-5)
    ev = np.linalg.eigvals(r["A"])
    assert ev.real.max() < 1e-3                                              # no unstable mode at the default trim point
    sp = ev[np.argmax(np.abs(ev.imag))]
    assert 3.0 < abs(sp.imag) < 10.0 and sp.real < -2.0                      # short-period pair
    o.linearize()                                                            # the member is restored: repeatable
    assert np.array_equal(o.linearize()["A"], r["A"])


def test_forward_differences_agree_with_central_differences(orc, fb):
    o = orc.Sim(variant=0, kin=2, dt=0.01)
    o.init_trim(EAS=45.0, h_e=2000.0, flaps=0.0)
    r = o.linearize()
    x_idx = {0: 21, 1: 22, 2: 23, 6: 21 + 0}   # column -> index in the full state: p q r at dyn(18..20)
    nxf = o.nx
    dyn = nxf - 6
    st, en = o.discrete
    x_full = o.x
    for col, full in ((0, dyn), (1, dyn + 1), (2, dyn + 2), (6, dyn + 3), (8, dyn + 5), (4, 12 + 1)):
        h = 1e-6 * max(1.0, abs(x_full[full]))
        fp, fm = [], []
        for sgn, acc in ((+1, fp), (-1, fm)):
            x = x_full.copy(); x[full] += sgn * h
            o.init_state(x, stall=st, eng_state=en)
            acc.append(o.xdot.copy())
        cd = (fp[0] - fm[0]) / (2 * h)
        rows = [(0, dyn), (1, dyn + 1), (2, dyn + 2), (6, dyn + 3), (7, dyn + 4), (8, dyn + 5), (14, 9)]
        for r_ss, r_full in rows:
            assert abs(r["A"][r_ss, col] - cd[r_full]) <= 2e-5 * max(1.0, abs(cd[r_full])), (col, r_ss, r["A"][r_ss, col], cd[r_full])
    o.init_state(x_full, stall=st, eng_state=en)


def test_linear_model_reproduces_a_small_elevator_step(orc, fb):
    """README.md:142-204 (example 3), scaled to a 0.01 elevator increment so that the comparison is in the linear range."""
    dt, T = 0.01, 5.0
    o = orc.Sim(variant=0, kin=2, dt=dt)
    o.init_trim()
    r = o.linearize()
    A, B = r["A"], r["B"]
    du = np.array([0.0, 0.0, 0.01, 0.0])
    o.set_input(fb.U["FB_U_ELEVATOR"], o.get_input(fb.U["FB_U_ELEVATOR"]) + 0.01)
    n = int(round(T / dt))
    dx = np.zeros(A.shape[0])
    f = lambda v: A @ v + B @ du
    theta_lin, theta_nl = [], []
    th0 = r["x0"][4]
    for _ in range(n):
        k1 = f(dx); k2 = f(dx + dt / 2 * k1); k3 = f(dx + dt / 2 * k2); k4 = f(dx + dt * k3)
        dx = dx + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        o.step(1)
        theta_lin.append(dx[4]); theta_nl.append(o.y(fb.Y["FB_Y_THETA"]) - th0)
    theta_lin, theta_nl = np.array(theta_lin), np.array(theta_nl)
    assert np.abs(theta_nl).max() > 5e-3                                       # the step does something
    assert np.abs(theta_lin - theta_nl).max() < 0.05 * np.abs(theta_nl).max()   # and the linear model follows it


def test_linearized_ss_known_answers_of_the_reference(orc):
    """lib/FlightPhysics/test/test_linearization.jl:14-45 transcribed: a LinearizedSS built from x0 = (1, 0.5, 0.3, 5), u0 = (0.1,
    0.2), y0 = (0.3, 0.8, 2, 3, -9.8) with xdot0 = x0, A = x0 x0', B = x0 u0', C = y0 x0', D = y0 u0', evaluated at x = 2 x0,
    u = 3 u0, must give xdot = xdot0 + A (x - x0) + B (u - u0) and y = y0 + C (x - x0) + D (u - u0) — the reference asserts
    `all(.==)`, i.e. its non-allocating update order reproduces the one-line formula exactly."""
    x0 = np.array([1.0, 0.5, 0.3, 5.0]); u0 = np.array([0.1, 0.2]); y0 = np.array([0.3, 0.8, 2.0, 3.0, -9.8])
    A, B, C, D = np.outer(x0, x0), np.outer(x0, u0), np.outer(y0, x0), np.outer(y0, u0)
    x, u = 2 * x0, 3 * u0
    xd, y = orc.lss_rhs(A, B, C, D, x0.copy(), x0, u0, y0, x, u)
    # Julia's `xdot0 + A * (x - x0) + B * (u - u0)` is left-associative: (xdot0 + A dx) + B du
    assert np.array_equal(xd, (x0 + A @ (x - x0)) + B @ (u - u0))
    assert np.array_equal(y, (y0 + C @ (x - x0)) + D @ (u - u0))
    # at the linearisation point the model returns xdot0 and y0 themselves
    xd, y = orc.lss_rhs(A, B, C, D, x0.copy(), x0, u0, y0, x0, u0)
    assert np.array_equal(xd, x0) and np.array_equal(y, y0)
    # and the RK4 stepping of the same model (orc.lss_step) is consistent with it: one small step ~ x + dt f(x)
    x1 = orc.lss_step(A, B, x0, u0, x0.copy(), x, u, 1e-6, 1)
    f0, _ = orc.lss_rhs(A, B, C, D, x0.copy(), x0, u0, y0, x, u)
    assert np.allclose((x1 - x) / 1e-6, f0, rtol=1e-4)
