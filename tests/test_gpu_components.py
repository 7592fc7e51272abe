"""GPU parity of the non-default components (variant FB_C172XA2): a FlyByWireActuation of seven Actuator2
units (lib/FlightApps/src/c172/c172x/c172x.jl:68-106) and a Propeller(VariablePitch) (lib/FlightPhysics/src/propellers.jl:
341-400), through the C ABI against the CPU oracle. Bars as everywhere: continuous state <= 1e-9 scaled, discrete state exact."""
import numpy as np
import pytest

from parity_util import scaled_err, state_floors

pytestmark = pytest.mark.gpu
SEED = 20261018
VAR_A2 = 4


def _params(n, rng):
    import bench
    lat = np.deg2rad(rng.uniform(-60, 60, n)); lon = rng.uniform(-np.pi, np.pi, n)
    return dict(EAS=rng.uniform(38, 47, n), h_e=rng.uniform(300, 2500, n) + bench.geoid_height(lat, lon), lat=lat, lon=lon,
                psi_nb=rng.uniform(-np.pi, np.pi, n), fuel_load=rng.uniform(0.2, 1, n), m_pilot=rng.uniform(50, 100, n),
                m_baggage=rng.uniform(0, 80, n))


@pytest.mark.parametrize("kin,pitch", [(0, (-0.1, 0.2, 11)), (2, None), (1, (-0.1, 0.2, 11))])
def test_actuator2_and_variable_pitch_parity(fb, orc, kin, pitch):
    """Trim, right-hand side, outputs and 8 s of stepping with actuator command doublets and blade-pitch changes scheduled as
    input events (the second-order units ring at 5 Hz: w_n dt = 0.31, well inside what RK4 resolves)."""
    n, dt, steps = 96, 0.01, 800
    rng = np.random.default_rng(SEED + 40 + kin)
    p = _params(n, rng)
    prop = fb.VariablePitch(*pitch) if pitch else fb.FixedPitch()
    ac = fb.Cessna172Xv0A2([fb.WA, fb.ECEF, fb.NED][kin](), propeller=prop)
    sim = fb.BatchSimulation(fb.SimpleWorld(ac), n=n, dt=dt, t_end=100.0, steps_per_launch=64)
    assert sim.nx == {0: 41, 1: 40, 2: 38}[kin]
    u0 = rng.uniform(0.0, 0.4, n) if pitch else np.zeros(n)
    sim.set_input("prop_pitch", u0)                       # the propeller input is set before the trim, which then balances against it
    ok, ts, cost = sim.init(fb.C172.TrimParameters(**p))
    assert ok.all() and cost.max() <= 1e-16
    U, Y = fb.U, fb.Y
    idx = list(range(0, n, 6))
    fl = state_floors(sim.nx, VAR_A2, kin)
    xg, xdg = sim.x, sim.xdot
    assert (xg[12:26:2] == 0.0).all()                     # actuator velocities at rest, positions = commands (c172x.jl:268-274)
    assert np.array_equal(xg[13:26:2], np.array([sim.get_input(U["FB_U_ACT_THROTTLE"] + a) for a in range(7)]))
    oracles = []
    for i in idx:
        o = orc.Sim(variant=VAR_A2, kin=kin, dt=dt, prop_pitch=pitch)
        o.set_input(U["FB_U_PROP_PITCH"], u0[i])
        ok_o, ts_o, _ = o.init_trim(**{k: float(v[i]) for k, v in p.items()})
        assert ok_o
        assert np.max(scaled_err(xg[:, i], o.x, fl)) < 1e-9, i            # trim point
        st, en = o.discrete
        o.want_y_extra(True)
        o.init_state(xg[:, i], stall=st, eng_state=en)                   # right-hand side at the SAME state
        xdo = o.xdot
        assert (np.abs(xdg[:, i] - xdo) / np.maximum(np.abs(xdo), fl)).max() < 1e-9, i
        oracles.append(o)
    names = ["theta", "eas", "alpha", "prop_thrust", "prop_j", "prop_dbeta", "eng_omega", "mass", "f_z"]
    out = {k: sim.get_output(k) for k in names}
    for i, o in zip(idx, oracles):
        for k in names:
            a, b = out[k][i], o.y(fb.OUTPUT_NAMES[k])
            assert abs(a - b) <= 1e-9 * max(abs(b), 1e-2), (k, i, a, b)
    if pitch:
        assert np.allclose(out["prop_dbeta"], pitch[0] + (pitch[1] - pitch[0]) * u0, rtol=0, atol=1e-15)
    # ---- stepping: elevator / aileron / flap command doublets through the second-order units, blade pitch stepped mid-run
    ele_f, ail_f, flp_f, pp_f = U["FB_U_ACT_ELEVATOR"], U["FB_U_ACT_AILERON"], U["FB_U_ACT_FLAPS"], U["FB_U_PROP_PITCH"]
    ele0, ail0 = sim.get_input(ele_f), sim.get_input(ail_f)
    ev = []
    for i in range(n):
        ev += [(i, 100, ele_f, ele0[i] + 0.1), (i, 200, ele_f, ele0[i] - 0.1), (i, 300, ele_f, ele0[i]),
               (i, 350, ail_f, ail0[i] + 0.1), (i, 450, ail_f, ail0[i] - 0.1), (i, 550, ail_f, ail0[i]),
               (i, 600, flp_f, 0.5)]
        if pitch:
            ev += [(i, 250, pp_f, u0[i] + 0.3), (i, 500, pp_f, 7.0 if i % 2 else 0.1)]   # 7.0 saturates to 1 (Ranged)
    ev.sort(key=lambda e: (e[0], e[1]))
    sim.schedule_inputs(*zip(*ev))
    for o, i in zip(oracles, idx):
        for (mi, si, fi, vi) in ev:
            if mi == i:
                o.add_event(si, fi, vi)
    sim.step(steps * dt)
    for o in oracles:
        o.step(steps)
    xg = sim.x
    stg, eng = sim.discrete()
    assert (sim.status() == 0).all()
    worst = 0.0
    for i, o in zip(idx, oracles):
        assert o.status == 0
        err = scaled_err(xg[:, i], o.x, fl)
        worst = max(worst, float(err.max()))
        assert err.max() < 1e-9, (i, int(err.argmax()), float(err.max()))
        assert o.discrete == (int(stg[i]), int(eng[i]))
    if pitch:
        want = np.where(np.arange(n) % 2 == 1, 1.0, 0.1)
        assert np.array_equal(sim.get_input("prop_pitch"), want)
        assert np.allclose(sim.get_output("prop_dbeta"), pitch[0] + (pitch[1] - pitch[0]) * want, rtol=0, atol=1e-15)
    assert np.abs(xg[12:26:2]).max() < 1e-3 and np.abs(xg[12 + 9] - 0.5).max() < 1e-3     # the units have settled on their commands
    print(f"Actuator2 / {'VariablePitch' if pitch else 'FixedPitch'} kin {kin}: worst scaled state error {worst:.3e}")
    sim.close()


def test_actuator2_on_the_ground_and_custom_parameters(fb, orc):
    """The variant runs in the complete (ground-capable) integrator only: members dropped onto the runway with the brakes
    commanded through their second-order units, against the oracle over the first 1.5 s (the struts load at ~0.4 s; the bound
    is the airborne one until then and 1e-5 after: the order of the ground-contact rounding spread that
    tests/test_gpu_ground.py measures for the same strut code), with non-default actuator parameters (w_n = 20 rad/s, zeta = 1)."""
    import bench
    n, dt = 32, 0.01
    rng = np.random.default_rng(SEED + 47)
    g = float(bench.geoid_height(0.0, 0.0))
    h = g + 1.9 + rng.uniform(0.3, 1.0, n)
    ac = fb.Cessna172Xv0A2(fb.WA(), actuator=fb.Actuator2(omega_n=20.0, zeta=1.0))
    sim = fb.BatchSimulation(fb.SimpleWorld(ac), n=n, dt=dt, t_end=100.0, steps_per_launch=50)
    sim.init(fb.C172.Init(h_e=h))
    br = rng.uniform(0.2, 1.0, n)
    sim.set_input("act_brake_left", br); sim.set_input("act_brake_right", br)
    oracles = []
    for i in range(n):
        o = orc.Sim(variant=VAR_A2, kin=0, dt=dt, act2=(20.0, 1.0))
        o.init(h_e=h[i])
        o.set_input(fb.U["FB_U_ACT_BRAKE_LEFT"], br[i]); o.set_input(fb.U["FB_U_ACT_BRAKE_RIGHT"], br[i])
        oracles.append(o)
    fl = state_floors(sim.nx, VAR_A2, 0)
    for k, tol in ((20, 1e-9), (130, 1e-5)):
        sim.step(k * dt)
        xg = sim.x
        wow = sim.get_output("wow_left")
        for i, o in enumerate(oracles):
            o.step(k)
            assert sim.status()[i] == o.status == 0
            assert wow[i] == o.y(fb.Y["FB_Y_WOW_LEFT"])
            assert scaled_err(xg[:, i], o.x, fl).max() < tol, (k, i, float(scaled_err(xg[:, i], o.x, fl).max()))
    assert wow.all()
    # brake units follow the critically damped closed form p = cmd (1 - (1 + w t) exp(-w t))
    t = 1.5
    assert np.allclose(sim.x[12 + 11], br * (1 - (1 + 20.0 * t) * np.exp(-20.0 * t)), rtol=0, atol=1e-6)
    sim.close()


def test_linearize_is_refused_for_the_component_variant(fb):
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Xv0A2(fb.NED())), n=4, dt=0.01, t_end=1.0)
    sim.init(fb.C172.TrimParameters())
    with pytest.raises(fb.FlightBatchError, match="not offered"):
        sim.linearize()
    sim.close()


def test_variable_pitch_needs_a_matching_blob(fb):
    with pytest.raises(fb.FlightBatchError, match="prop.pitch3|prop.tables3"):
        fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Xv0A2(fb.WA(), propeller=fb.VariablePitch(0.0, 0.3))), n=4, dt=0.01, t_end=1.0)
