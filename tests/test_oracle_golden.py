"""CPU tests: the oracle (oracle/, a C++ restatement of the reference) against the reference's own golden values.
Every expected number below is transcribed from the reference's tests / data, with the file:line it comes from."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "trim_points.json")))


# ---- trim: the reference's stored trim solutions (gain-schedule files) -------------------------
@pytest.mark.parametrize("pt", GOLD["points"], ids=lambda p: f"EAS{p['EAS']}_h{p['h_e']}")
def test_stored_reference_trim_is_an_equilibrium_of_the_oracle(orc, pt):
    # The reference stopped BOBYQA at cost <= 1e-16 (c172.jl:927): its stored trim state must sit below that
    # value in the oracle too, i.e. the oracle's f_ode! reproduces the reference's at these 28 states.
    s = orc.Sim(variant=1, kin=2, dt=0.01)   # C172X.Vehicle{NED}, as in c172x_design.jl:106-118
    cost = s.trim_cost_at(pt["trim_state"], EAS=pt["EAS"], h_e=pt["h_e"], flaps=pt["flaps"])
    assert cost <= 1.0e-16 * 1.02, cost


@pytest.mark.parametrize("pt", GOLD["points"][::3], ids=lambda p: f"EAS{p['EAS']}_h{p['h_e']}")
def test_oracle_trim_solver_converges_to_the_stored_trim(orc, pt):
    s = orc.Sim(variant=1, kin=2, dt=0.01)
    ok, ts, cost = s.init_trim(EAS=pt["EAS"], h_e=pt["h_e"], flaps=pt["flaps"])
    assert ok and cost < 1e-20
    # the reference's own solution carries its optimiser's stop tolerance (cost 1e-16 ~ 1e-7 in the weakest direction)
    assert np.max(np.abs(ts - np.array(pt["trim_state"]))) < 1e-6
    x = s.x   # state layout: aero 2, ldg 6, fuel 1, engine 3, act 7, NED kin 6, dyn 6
    assert abs(x[19 + 1] - pt["x_lon"][1]) < 1e-6            # theta
    assert abs(x[9] / (2700 * np.pi / 30) - pt["x_lon"][5]) < 1e-6   # n_eng


# ---- picture-level pin for config 1: README example 3, assets/elevator_step_response.png --------------------------
def test_config1_elevator_step_matches_readme_plot(orc, fb):
    o = orc.Sim(variant=0, kin=2, dt=0.02)   # Cessna172Sv0(NED()), Simulation default dt = 0.02 (sim.jl:188)
    ok, ts, c = o.init_trim()
    th = [o.y(fb.Y["FB_Y_THETA"])]
    o.step(50)
    o.set_input(fb.U["FB_U_ELEVATOR"], o.get_input(fb.U["FB_U_ELEVATOR"]) + 0.1)
    for _ in range(450):
        o.step(1)
        th.append(o.y(fb.Y["FB_Y_THETA"]))
    th = np.array(th)
    assert abs(th[0] - 0.0236) < 2e-4                        # theta_trim ~ 0.0236 rad (also te2te.h5 x_trim)
    assert abs(th.max() - 0.225) < 3e-3                      # peak ~ 0.225 rad ...
    assert abs((1 + 0.02 * th.argmax()) - 7.5) < 0.3         # ... at t ~ 7.5 s
    assert abs(th[-1] - 0.19) < 3e-3                         # ~ 0.19 rad at t = 10 s


# ---- kinematics: lib/FlightPhysics/test/test_kinematics.jl:37-41,75-93 ----------------------------------------------
def test_kinematic_mechanisations_agree(orc):
    ic = dict(lat=np.pi / 3, lon=-np.pi / 6, h_orth=12354.0, w_wb_b=(0.1, 0.1, -0.2), v_eb_n=(100, 10, -4))
    r0 = {k: orc.kin_sim(k, 0.02, 0, **ic) for k in (0, 1, 2)}      # after init (:49-62)
    for k in (0, 2):
        for f in ("q_nb", "n_e", "h_e", "v_eb_b", "w_eb_b"):
            assert np.allclose(r0[1][f], r0[k][f], rtol=1.5e-8, atol=0)
    r = {k: orc.kin_sim(k, 0.02, 1000, **ic) for k in (0, 1, 2)}    # Simulation(...; t_end = 20), default dt (:75-93)
    for a, b in ((1, 0), (0, 2)):
        for f in ("q_nb", "v_eb_n", "h_e"):
            assert np.allclose(r[a][f], r[b][f], rtol=1.5e-8, atol=0), (a, b, f)


# ---- dynamics: lib/FlightPhysics/test/test_dynamics.jl:21-66 -----------------------------------------------------------
def _qrot(q, v):
    qi = np.array(q[1:]); v = np.asarray(v, float)
    return v + 2 * np.cross(qi, q[0] * v + np.cross(qi, v))


def _q_from_euler(psi, theta, phi):
    def mul(a, b):
        return np.array([a[0] * b[0] - a[1:] @ b[1:], *(a[0] * b[1:] + b[0] * a[1:] + np.cross(a[1:], b[1:]))])
    rz = np.array([np.cos(psi / 2), 0, 0, np.sin(psi / 2)]); ry = np.array([np.cos(theta / 2), 0, np.sin(theta / 2), 0])
    rx = np.array([np.cos(phi / 2), np.sin(phi / 2), 0, 0])
    return mul(mul(rz, ry), rx)


def test_dynamics_known_answers(orc):
    ic = dict(lat=0.4, lon=0.5, h_orth=1000.0, w_wb_b=(1.0, 1.0, 0.0), q_nb=_q_from_euler(1.0, 0.5, 0.4), v_eb_n=(100, 0, 0))
    d = orc.dyn_test(1.0, (1, 1, 1), (0, 0, 0), (1, 2, 1), (0, 0, 0), **ic)          # Ob = Oc (:38-50)
    q_nb_inv = np.array([d["q_nb"][0], *(-d["q_nb"][1:])])
    g_b = _qrot(q_nb_inv, [0, 0, d["gravity"]])
    assert np.allclose(d["wdot"], 0, atol=1e-12)                                      # all(ω̇ .≈ 0)
    assert np.allclose(d["vdot"], np.array([1, 2, 1]) + g_b, rtol=1.5e-8)             # v̇ ≈ F + q_nb'(g_n)
    assert np.allclose(d["a_eb_b"], d["vdot"], rtol=1.5e-8)
    w_ie = np.array([0, 0, 7.292115e-05])
    q_en_inv = np.array([d["q_en"][0], *(-d["q_en"][1:])])
    G_n = np.array([0, 0, d["gravity"]]) + _qrot(q_en_inv, np.cross(w_ie, np.cross(w_ie, d["r_eb_e"])))
    assert np.allclose(d["a_ib_b"], np.array([1, 2, 1]) + _qrot(q_nb_inv, G_n), rtol=1.5e-8)
    d = orc.dyn_test(1.0, (1, 1, 1), (1, 0, 0), (0, 0, 1), (0, 0, 0), **ic)          # CoM 1 m ahead (:53-63)
    assert abs(d["wdot"][1] - 1) < 1.5e-8


# ---- landing gear: lib/FlightPhysics/test/test_landing_gear.jl:98-209 ----------------------------------------------------
def test_friction_coefficients(orc):
    l = orc.lib()
    assert l.orc_get_mu(0, 0, 0.0075) == pytest.approx(0.025)      # :98
    assert l.orc_get_mu(1, 0, 0.0075) == pytest.approx(0.5)        # :99
    assert l.orc_get_mu(1, 1, 1e-5) == pytest.approx(0.25)         # :100
    assert l.orc_get_mu(1, 2, 10.0) == pytest.approx(0.025)        # :101


def test_landing_gear_unit_known_answers(orc):
    t = orc.ldg_test
    assert t(h_orth=1.1)["wow"] is False                                                           # :118-122
    r = t(h_orth=0.9)                                                                              # :133-140
    assert r["wow"] and r["xi"] == pytest.approx(-0.1, rel=1e-7) and r["F_dmp"] == pytest.approx(2500, rel=1e-7)
    assert tuple(r["mu_eff"]) == (0, 0) and tuple(r["f_c"][:2]) == (0, 0) and r["f_c"][2] < 0
    r = t(h_orth=0.9, q_nb=_q_from_euler(0, 0, np.pi / 12))                                        # :143-146
    assert r["xi"] > -0.1 and r["F_dmp"] < 2500
    r = t(h_orth=0.9, v_eb_n=(0, 0, 1))                                                            # :149-152
    assert r["xi_dot"] < 0 and r["F_dmp"] == pytest.approx(3500, rel=1e-6)
    r = t(h_orth=0.9, v_eb_n=(1e-4, 0, 0))                                                         # :155-162
    assert np.allclose(r["v_xy"], [1e-4, 0], atol=1e-6) and r["mu_max"][0] <= r["mu_roll"] and r["mu_eff"][1] == 0
    assert r["mu_eff"][0] < 0 and abs(r["mu_eff"][0]) < r["mu_max"][0] and r["xdot"][0] < 0
    r = t(h_orth=0.9, v_eb_n=(0, -1e-4, 0))                                                        # :165-168
    assert r["F_b"][1] > 0 and r["xdot"][1] > 0
    r = t(h_orth=0.9, v_eb_n=(0, -1, 0))                                                           # :171-174
    assert r["sat_out"][1] == 1 and r["xdot"][1] == 0
    assert np.allclose(t(h_orth=0.9, v_eb_n=(10, 0, 1))["v_xy"], [10, 0], atol=1e-5)                # :177-179
    assert np.allclose(t(h_orth=0.9, w_wb_b=(1, 0, 0))["v_xy"], [0, -0.9], atol=1e-5)               # :182-184
    assert t(h_orth=0.9, q_nb=_q_from_euler(0, 0, np.pi / 12), v_eb_n=(1e-4, 0, 0))["F_b"][1] < 0  # :187-189
    r = t(h_orth=0.9, v_eb_n=(10, 0, 1), steer_input=0.5)                                          # :192-196
    assert r["v_xy"][0] < 10 and r["v_xy"][1] < 0
    r = t(h_orth=0.9, v_eb_n=(1e-4, 0, 0), brake=1.0)                                              # :199-202
    assert r["mu_max"][0] > r["mu_roll"]


# ---- piston engine charts: lib/FlightPhysics/test/test_piston.jl:57-126 -----------------------------------------------------
def test_engine_lookup_chart_points(orc):
    l = orc.lib()
    inHg, p_std, w_r, P = 3386.389, 101325.0, 2700.0, 200.0
    ft = 0.3048
    E = lambda which, n, a, b=0.0: l.orc_engine_lookup(0.15, 1.4, which, n, a, b)
    h2d = lambda h_ft: l.orc_h2delta(h_ft * ft)
    for rpm, map_, h in ((1800, 20, 9500), (2700, 22, 7000), (2100, 16, 15250), (2300, 12, 22000)):       # :70-73
        assert abs(E(0, rpm / w_r, inHg * map_ / p_std) - h2d(h)) < 0.1
    for rpm, map_, hp in ((1800, 20, 71), (2050, 24, 113), (2400, 17, 85), (2400, 28.8, 176)):             # :81-84
        assert abs(E(1, rpm / w_r, inHg * map_ / p_std) * P - hp) < 1
    for rpm, h, hp in ((1800, 3e3, 108), (2300, 2.4e3, 153), (2500, 10e3, 129), (2000, 20e3, 65)):         # :94-97
        assert abs(E(2, rpm / w_r, h2d(h)) * P - hp) < 3
    assert E(4, 0.15, 0.0, 1.0) == pytest.approx(0, abs=1e-12)                                              # :110
    assert E(4, 0.15, E(3, 0.15, 1.0), 1.0) == pytest.approx(0, abs=1e-12)                                  # :111
    assert E(4, 0.075, 0.5, 1.0) == pytest.approx(0, abs=1e-12)                                             # :112
    assert E(4, 0.225, 0.5, 1.0) > E(4, 0.225, 0.3, 1.0)                                                    # :115
    assert 71 < E(4, 1800 / w_r, inHg * 20 / p_std, h2d(3e3)) * P < 84                                       # :118
    assert 131 < E(4, 2310 / w_r, inHg * 23.6 / p_std, h2d(2.4e3)) * P < 139                                 # :119
    assert 102 < E(4, 2500 / w_r, inHg * 18 / p_std, h2d(10e3)) * P < 119                                    # :120


# ---- geodesy: lib/FlightPhysics/test/test_geodesy.jl:181-192 (Cartesian <-> Geographic torture test) --------------------------
def test_geodesy_roundtrip(orc):
    l = orc.lib()
    rng = np.random.default_rng(1)
    out = np.zeros(7)
    import ctypes as C
    for _ in range(1000):
        lat, lon, h = rng.uniform(-np.pi / 2, np.pi / 2), rng.uniform(-np.pi, np.pi), rng.uniform(-900, 80000)
        l.orc_geo_roundtrip(lat, lon, h, out.ctypes.data_as(C.POINTER(C.c_double)))
        n = np.array([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)])
        assert np.allclose(out[3:6], n, atol=1e-9) and abs(out[6] - h) < 1e-6


# ---- propeller / engine tables: independent numpy build (tools/build_tables.py) vs the oracle's C++ build --------------------
def test_table_blob_matches_oracle_generated_tables(orc):
    import struct
    blob = open(os.path.join(HERE, "..", "flight.jl_b200", "data", "c172_tables.bin"), "rb").read()
    n = struct.unpack_from("<I", blob, 8)[0]
    sec = {}
    for i in range(n):
        nm, dt, nd, d0, d1, d2, d3, off, nb = struct.unpack_from("<24sII4IQQ", blob, 16 + 64 * i)
        sec[nm.rstrip(b"\0").decode()] = np.frombuffer(blob, dtype="<f4" if dt else "<f8", count=nb // (4 if dt else 8), offset=off)
    w = orc.world()
    assert np.allclose(w.prop_table().ravel(), sec["prop.tables"], rtol=1e-10, atol=1e-13)
    a, b, c = w.engine_tables()
    assert np.allclose(a.ravel(), sec["eng.mu_wot"], atol=1e-14) and np.allclose(b.ravel(), sec["eng.pi_std"], atol=1e-14)
    assert np.allclose(c.ravel(), sec["eng.pi_wot"], atol=1e-14)


def test_algorithmic_flop_count_of_record(orc):
    """The op-counting build fixes A_step: bench.py's constant must equal the measured count."""
    import bench
    lc = orc.lib(counting=True)
    s = orc.Sim(0, 0, 0.01, counting=True)
    s.init_trim(EAS=45.0, h_e=1500.0)
    orc.ops_reset(lc); s.rhs(); r = orc.ops_get(lc)["total"]
    orc.ops_reset(lc); s.step(100); t = orc.ops_get(lc)["total"] / 100
    a_step = 4 * r + (t - 5 * r)     # 4 RHS + RK combine + f_step! (the 5th reference evaluation is redundant, App. B)
    assert abs(a_step - bench.A_STEP_FLOP) <= 0.01 * bench.A_STEP_FLOP, (r, t, a_step)


def test_algorithmic_flop_count_on_the_ground(orc):
    """The same count for a member standing on all three wheels with the engine running (cfg 4's roofline numerator)."""
    import bench
    import flightbatch as fb
    lc = orc.lib(counting=True)
    s = orc.Sim(0, 0, 0.01, counting=True)
    s.init(h_e=1.9 + float(bench.geoid_height(0.0, 0.0)), engine_state=2.0, n_eng=0.3, throttle=0.2)
    s.step(300)
    assert [s.y(fb.OUTPUT_NAMES[k]) for k in ("wow_left", "wow_right", "wow_nose")] == [1.0, 1.0, 1.0]
    orc.ops_reset(lc); s.rhs(); r = orc.ops_get(lc)["total"]
    orc.ops_reset(lc); s.step(100); t = orc.ops_get(lc)["total"] / 100
    a_step = 4 * r + (t - 5 * r)
    assert abs(a_step - bench.A_STEP_FLOP_GROUND) <= 0.01 * bench.A_STEP_FLOP_GROUND, (r, t, a_step)


# ---- rotation algebra: lib/FlightPhysics/test/test_attitude.jl:29-82 (RQuat), :182-240 (REuler) ------------------------------
def test_rotation_algebra_known_answers(orc):
    """Composition, inversion, quaternion and Euler-angle kinematics with the primitives every right-hand side uses, on the
    reference's own test values: transformation by a composed rotation = successive transformations (:66, :199), inverse (:67),
    q' = 1/2 q (x) (0, w) and its inverse (:76-78), Euler rates at the reference's six sample attitudes (:203-209), the
    Euler -> omega round trip (:213-215), Euler -> quaternion -> Euler over the 1,000-point angle lattice (:225-233)."""
    o = orc.attitude_test()
    assert np.abs(o[0:3]).max() < 1e-14 and np.abs(o[3:6]).max() < 1e-14 and np.abs(o[33:36]).max() < 1e-14
    assert np.abs(o[6:10]).max() == 0.0                       # the reference asserts equality here (:76)
    assert np.abs(o[10:13]).max() < 1e-13
    assert np.allclose(o[13:16], [1.0, 0.0, 0.0], rtol=0, atol=1e-10)          # dt(REuler(phi = pi/2), (0, 1, 0)) = (1, 0, 0)
    assert np.allclose(o[16:19], [0.0, -1.0, 0.0], rtol=0, atol=1e-10)         # dt(REuler(phi = pi), (0, 1, 0)) = (0, -1, 0)
    assert np.allclose(o[19:22], [0.0, 0.0, 1.0], rtol=0, atol=1e-10)          # dt(REuler(theta = pi/3, phi = pi/4), (1, 0, 0)) = (0, 0, 1)
    assert np.abs(o[22:25]).max() < 1e-14                                      # Euler rates do not depend on the heading
    assert np.abs(o[25:28]).max() < 1e-13
    assert o[28] < 1e-9                                                         # the reference's isapprox on the lattice
    assert np.abs(o[29:33]).max() == 0.0                                        # RQuat(REuler()) = (1, 0, 0, 0)


# ---- discrete compensators: lib/FlightPhysics/test/test_control.jl:88-146 (Integrator), :254-372 (PID) ------------------------
def test_discrete_integrator_reference_sequence(orc):
    """test_control.jl:88-146 transcribed: bounds (-1, 2), Dt = 0.01; 2 s of input -1 saturate low and halt; input +1 releases,
    2 s later saturates high and halts; reversing releases; an external saturation of the input's sign halts, of the other
    sign does not; f_init! zeroes the state."""
    m = orc.DiscreteIntegrator(lo=-1.0, hi=2.0)
    y = m.run(-1.0, 0, 0.01, 200)
    assert y["x0"] <= -1 and abs(y["output"] + 1.0) < 1e-12 and y["sat_out"] == -1 and y["halted"]
    y = m.run(1.0, 0, 0.01, 200)
    assert y["sat_out"] == 0 and not y["halted"]
    y = m.run(1.0, 0, 0.01, 200)
    assert y["sat_out"] == 1 and y["halted"]
    y = m.run(-1.0, 0, 0.01, 100)
    assert y["sat_out"] == 0 and not y["halted"]
    y = m.run(-1.0, +1, 0.01, 100)        # sat_ext = -sign(input)
    assert not y["halted"]
    y = m.run(-1.0, -1, 0.01, 100)        # sat_ext = sign(input)
    assert y["halted"]
    m.init()
    y = m.run(0.0, 0, 0.01, 1)
    assert y["x1"] == 0 and y["output"] == 0 and y["sat_out"] == 0 and not y["halted"]


def test_discrete_pid_reference_sequence(orc):
    """test_control.jl:254-372 transcribed: k_p = k_i = 1, k_d = 0.1, Dt = 0.01, unit input for 1 s -> y_p = 1, y_i ~ 1, output ~ 2;
    with bounds (-1, 1) it saturates and the integrator halts; reversing saturates low; a small input releases; external
    saturation halts only with the integrator input's sign; as a filtered derivative (k_d = 1, tau_f = 0.2) the output
    decays to zero within 5 s; and the discretisation follows the continuous PID k_p + k_i / s + k_d s / (tau_f s + 1),
    whose unit-step response k_p + k_i t + (k_d / tau_f) exp(-t / tau_f) the reference simulates as a LinearizedSS, to 1e-3
    at t = 2 with Dt = 1e-4."""
    m = orc.DiscretePID(k_p=1.0, k_i=1.0, k_d=0.1)
    y = m.run(1.0, 0, 0.01, 100)
    assert y["y_p"] == 1.0 and abs(y["y_i"] - 1.0) < 1e-9 and abs(y["out_free"] - 2.0) < 1e-6 and abs(y["output"] - 2.0) < 1e-6
    assert y["sat_out"] == 0 and not y["int_halted"]
    m.lo, m.hi = -1.0, 1.0
    y = m.run(1.0, 0, 0.01, 100)
    assert y["out_free"] > 2.0 and abs(y["output"] - 1.0) < 1e-12 and y["sat_out"] == 1 and y["int_halted"]
    y = m.run(-1.0, 0, 0.01, 200)
    assert y["sat_out"] == -1 and y["int_halted"]
    y = m.run(0.1, 0, 0.01, 100)
    assert y["sat_out"] == 0 and not y["int_halted"]
    y = m.run(0.1, -int(np.sign(y["u_i"])), 0.01, 1)
    assert not y["int_halted"]
    y = m.run(0.1, int(np.sign(y["u_i"])), 0.01, 1)
    assert y["int_halted"]
    m.init()
    y = m.run(0.0, 0, 0.01, 1)
    assert y["y_i"] == 0 and y["y_d"] == 0 and y["y_p"] == 0 and not y["int_halted"]
    # filtered derivative
    d = orc.DiscretePID(k_p=0.0, k_i=0.0, k_d=1.0, tau_f=0.2)
    y = d.run(1.0, 0, 0.01, 1)
    assert y["y_d"] > 0.0
    y = d.run(1.0, 0, 0.01, 500)
    assert abs(y["y_d"]) < 1e-6
    # numerical correctness of the discretisation against the continuous transfer function
    k_p, k_i, k_d, tau_f = 1.0, 1.0, 0.2, 0.1
    c = orc.DiscretePID(k_p=k_p, k_i=k_i, k_d=k_d, tau_f=tau_f)
    y = c.run(1.0, 0, 1e-4, 20000)
    assert abs(y["output"] - (k_p + k_i * 2.0 + k_d / tau_f * np.exp(-2.0 / tau_f))) < 1e-3


# ---- continuous PI with anti-windup: lib/FlightPhysics/test/test_control.jl:36-84 --------------------------------------------------
def test_continuous_pi_reference_sequence(orc):
    """test_control.jl:36-84 transcribed for the two lanes of PIContinuous{2} (k_p = k_i = 1, default dt = 0.02): lane 1 bounded to
    (-1, 1) saturates at once and its integrator halts (|y_i| < 0.1, output -1); lane 2 unbounded integrates to y_i = -2,
    output = -3 after 2 s; an external saturation halts lane 2 only when it has the sign of the integrator's input."""
    inf = float("inf")
    x1, y1 = orc.pi_run(1.0, 1.0, 0.0, -1.0, 1.0, 0.0, -1.0, 0, 0.02, 100)
    assert y1["sat_out"] == -1 and y1["int_halted"] and abs(y1["y_i"]) < 0.1 and abs(y1["output"] + 1.0) < 1e-12
    x2, y2 = orc.pi_run(1.0, 1.0, 0.0, -inf, inf, 0.0, -1.0, 0, 0.02, 100)
    assert y2["sat_out"] == 0 and not y2["int_halted"] and abs(y2["y_i"] + 2.0) < 1e-2 and abs(y2["output"] + 3.0) < 1e-2
    x2, y2 = orc.pi_run(1.0, 1.0, 0.0, -inf, inf, x2, -1.0, -int(np.sign(y2["u_i"])), 0.02, 50)
    assert not y2["int_halted"]
    x2b, y2 = orc.pi_run(1.0, 1.0, 0.0, -inf, inf, x2, -1.0, int(np.sign(y2["u_i"])), 0.02, 50)
    assert y2["int_halted"] and x2b == x2 and x2 != 0


# ---- wander-azimuth geodesy: lib/FlightPhysics/test/test_geodesy.jl:20-112 ---------------------------------------------------------
def test_wander_azimuth_geodesy_known_answers(orc):
    """The local-tangent-frame rotation of an n-vector has zero wander angle (:45); composing it with Rz(psi_nw) gives back psi_nw
    (:50) and the same n-vector (:55); LatLon -> NVector -> ltf -> o Rz(psi) -> NVector -> LatLon is the identity over the
    reference's 1,000-point torture lattice including the poles and the date line (:84-92); orthometric -> ellipsoidal ->
    orthometric altitude is exact (:108)."""
    o = orc.geodesy_test()
    assert abs(o[0]) < 1e-15 and abs(o[1]) < 1e-15 and np.abs(o[2:5]).max() < 1e-15
    assert o[5] < 1e-9
    assert o[6] == 0.0 and abs(o[7]) < 1e-15


# ---- trim constraint and model-level tests: test_aircraft_base.jl:16-40, test_c172s.jl:20-68, test_c172x.jl:18-67 -------------------
def test_theta_constraint_recovers_the_flight_path_angle(orc):
    """test_aircraft_base.jl:16-40: theta_nb from theta_constraint(v_wb_b, gamma_wb_n, phi_nb) must give back gamma_wb_n when the
    aerodynamic velocity is rotated to NED by REuler(psi_nb, theta_nb, phi_nb) (the reference's isapprox)."""
    assert abs(orc.theta_constraint_test()) < 1e-12


def test_default_trim_succeeds_and_ground_init_loads_the_struts(orc, fb):
    """test_c172s.jl / test_c172x.jl: `f_init!(vehicle, C172.TrimParameters())` reports success for the mechanical and the
    fly-by-wire vehicle (:36-38 / :33-34), and `KinInit(h = elevation + 1.8) |> C172.Init` puts the aircraft on its wheels
    (`ldg.left.strut.wow == true`, :58-68 / :57-67: the reference uses this state for full landing-gear code coverage)."""
    import bench
    for variant in (0, 1):
        s = orc.Sim(variant=variant, kin=0, dt=0.01)
        ok, ts, cost = s.init_trim()
        assert ok and cost <= 1e-16
        g = orc.Sim(variant=variant, kin=0, dt=0.01)
        g.init(h_e=float(bench.geoid_height(0.0, 0.0)) + 1.8)          # HorizontalTerrain: elevation 0 (orthometric) at NVector()
        assert g.y(fb.Y["FB_Y_WOW_LEFT"]) == 1.0 and g.y(fb.Y["FB_Y_WOW_RIGHT"]) == 1.0 and g.status == 0


# ---- quaternion algebra: lib/FlightPhysics/test/test_quaternions.jl:60-105 -----------------------------------------------------------
def test_hamilton_product_identities(orc):
    """Associativity (:94), distributivity (:95), inverse (:96), identity (:93), normalisation (:39), conjugation (:79), and the
    non-commutativity the reference insists on (`q1 \\ q2 != q2 / q1`, :99)."""
    o = orc.quat_test()
    assert o[0] < 1e-12 and o[1] < 1e-12 and o[2] < 1e-15 and o[3] == 0.0 and o[4] < 1e-15 and o[5] == 0.0
    assert o[6] > 0.1
