"""CPU tests of the non-default components: Actuator2 (lib/FlightApps/src/c172/c172x/c172x.jl:68-106) and
the VariablePitch propeller (lib/FlightPhysics/src/propellers.jl:341-400) in the oracle, pinned by what the reference holds for
them: the closed-form response of the second-order lag it states, and its own propeller tests
(lib/FlightPhysics/test/test_propellers.jl:85-120 Lookup, :149-169 VariablePitch) transcribed."""
import os
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VAR_A2 = 4


def _blob_section(name):
    path = os.path.join(ROOT, "flight.jl_b200", "data", "c172_tables.bin")
    b = open(path, "rb").read()
    n = struct.unpack_from("<I", b, 8)[0]
    for i in range(n):
        nm, dt, nd, d0, d1, d2, d3, off, nb = struct.unpack_from("<24sII4IQQ", b, 16 + 64 * i)
        if nm.rstrip(b"\0").decode() == name:
            return np.frombuffer(b, dtype="<f8", count=nb // 8, offset=off).reshape([d0, d1, d2, d3][:nd])
    return None


def test_actuator2_layout_and_init(orc, fb):
    """X(::Actuator2) = (v, p) per unit (c172x.jl:84): seven more states than Cessna172Xv0; f_init! assigns the command to p
    (c172x.jl:268-274) and leaves v at zero, so a trimmed aircraft has a stationary actuation."""
    s = orc.Sim(variant=VAR_A2, kin=0, dt=0.01)
    s1 = orc.Sim(variant=1, kin=0, dt=0.01)
    assert s.nx == s1.nx + 7 == 41
    ok, ts, cost = s.init_trim(EAS=45.0, h_e=1500.0)
    ok1, ts1, _ = s1.init_trim(EAS=45.0, h_e=1500.0)
    assert ok and ok1 and cost < 1e-16
    assert np.allclose(ts, ts1, rtol=0, atol=1e-12)        # same trim point: the actuators are at rest in both
    act = s.x[12:26]
    assert (act[0::2] == 0.0).all()
    cmds = [s.get_input(fb.U["FB_U_ACT_THROTTLE"] + i) for i in range(7)]
    assert np.array_equal(act[1::2], cmds)
    assert np.array_equal(act[1::2], s1.x[12:19])
    assert np.abs(s.xdot[12:26]).max() == 0.0              # c172x.jl:324 (all actuator derivatives vanish at trim)
    # everything outside the actuator block is the Xv0 state
    assert np.allclose(np.delete(s.x, np.s_[12:26]), np.delete(s1.x, np.s_[12:19]), rtol=0, atol=1e-12)


def test_actuator2_step_response_matches_the_closed_form(orc, fb):
    """x'' + 2 zeta w_n x' + w_n^2 x = w_n^2 cmd (c172x.jl:100-102) with the defaults w_n = 5 * 2 pi, zeta = 0.6 (:71): the
    underdamped step response p(t) = cmd + e0 exp(-zeta w_n t) (cos w_d t + zeta w_n / w_d sin w_d t). RK4 at dt = 0.01
    (w_n dt = 0.31) follows it to the scheme's fourth-order accuracy, and the position overshoots as the reference's comment
    says it does."""
    dt = 0.002
    s = orc.Sim(variant=VAR_A2, kin=0, dt=dt)
    assert s.init_trim(EAS=45.0, h_e=1500.0)[0]
    f = fb.U["FB_U_ACT_FLAPS"]          # flaps: slowest to act on the airframe, and nothing feeds back into an actuator anyway
    p0 = s.x[12 + 2 * 4 + 1]
    s.set_input(f, p0 + 0.3)
    wn, zt = 5 * 2 * np.pi, 0.6
    wd = wn * np.sqrt(1 - zt * zt)
    worst, peak = 0.0, 0.0
    for k in range(1, 101):
        s.step(1)
        t = k * dt
        ref = (p0 + 0.3) - 0.3 * np.exp(-zt * wn * t) * (np.cos(wd * t) + zt * wn / wd * np.sin(wd * t))
        worst = max(worst, abs(s.x[12 + 2 * 4 + 1] - ref))
        peak = max(peak, s.x[12 + 2 * 4 + 1])
        vref = 0.3 * np.exp(-zt * wn * t) * (wn * wn / wd) * np.sin(wd * t)
        assert abs(s.x[12 + 2 * 4] - vref) < 1e-5
    assert worst < 2e-7, worst
    assert peak > p0 + 0.3 + 0.02        # overshoot of a zeta = 0.6 unit: 9.5 %
    # custom parameters reach the model
    s2 = orc.Sim(variant=VAR_A2, kin=0, dt=dt, act2=(20.0, 1.0))
    assert s2.init_trim(EAS=45.0, h_e=1500.0)[0]
    s2.set_input(f, p0 + 0.3); s2.step(100)
    t = 100 * dt
    assert abs(s2.x[12 + 2 * 4 + 1] - ((p0 + 0.3) - 0.3 * np.exp(-20.0 * t) * (1 + 20.0 * t))) < 1e-7   # critically damped


def test_actuator_position_output_is_clamped_and_saturates_on_command(orc, fb):
    """Actuator2Y.pos = R(x.p): the overshooting state is clamped to the unit's range (c172x.jl:58-66,96); commands saturate
    on assignment (Ranged)."""
    s = orc.Sim(variant=VAR_A2, kin=0, dt=0.01)
    assert s.init_trim(EAS=45.0, h_e=1500.0)[0]
    s.set_input(fb.U["FB_U_ACT_FLAPS"], 7.0)
    assert s.get_input(fb.U["FB_U_ACT_FLAPS"]) == 1.0
    top = 0.0
    for _ in range(40):
        s.step(1); top = max(top, s.x[12 + 2 * 4 + 1])
    assert top > 1.0                      # the state overshoots the range ...
    s2 = orc.Sim(variant=VAR_A2, kin=0, dt=0.01)
    assert s2.init_trim(EAS=45.0, h_e=1500.0)[0]
    x = s.x.copy(); x[12 + 2 * 4 + 1] = 1.2
    s2.init_state(x); a = s2.xdot.copy()
    x[12 + 2 * 4 + 1] = 1.0
    s2.init_state(x); b = s2.xdot.copy()
    # ... but the airframe sees the clamped position: only the actuator's own derivative differs
    d = np.nonzero(a != b)[0]
    assert list(d) == [12 + 2 * 4]


def test_variable_pitch_propeller_reference_tests(orc, fb):
    """test_propellers.jl:149-169 transcribed: with Propeller(VariablePitch(-0.1, 0.2)) the input 0 gives the lower blade-pitch
    bound, 1 the upper one, and more pitch gives more thrust at the same state (v = 50 m/s, omega = 300 rad/s there; the
    trimmed cruise here)."""
    s = orc.Sim(variant=VAR_A2, kin=0, dt=0.01, prop_pitch=(-0.1, 0.2))
    assert s.init_trim(EAS=45.0, h_e=1500.0)[0]
    x = s.x.copy(); x[9] = 300.0
    Y = fb.Y
    s.set_input(fb.U["FB_U_PROP_PITCH"], 0.0); s.init_state(x); s.xdot
    assert s.y(Y["FB_Y_PROP_DBETA"]) == -0.1
    F0 = s.y(Y["FB_Y_PROP_THRUST"])
    s.set_input(fb.U["FB_U_PROP_PITCH"], 1.0); s.init_state(x); s.xdot
    assert abs(s.y(Y["FB_Y_PROP_DBETA"]) - 0.2) < 1e-15
    F1 = s.y(Y["FB_Y_PROP_THRUST"])
    assert F1 > F0 > 0
    s.set_input(fb.U["FB_U_PROP_PITCH"], 3.0)            # Ranged(0, 0, 1) saturates on assignment (propellers.jl:385)
    assert s.get_input(fb.U["FB_U_PROP_PITCH"]) == 1.0


def test_variable_pitch_lookup_reference_tests(orc, fb):
    """test_propellers.jl:85-120 transcribed for the 2-blade C172 propeller: Flat extrapolation on every axis, and a
    fixed-pitch lookup equals the variable-pitch one on the knot delta_beta = 0 (`lookup_fp(0.1, 0.3, 0) == lookup(0.1, 0.3, 0)`)."""
    s3 = orc.Sim(variant=VAR_A2, kin=0, dt=0.01, prop_pitch=(0.0, 0.1, 11))
    t3 = s3.prop3()
    assert t3.shape == (21, 21, 11, 6)
    fp = orc.world().prop_table().reshape(6, 21, 21)     # the FixedPitch table of the C172 (delta_beta = 0)
    assert np.array_equal(t3[:, :, 0, :], np.moveaxis(fp, 0, -1))
    # through the model: same state, FixedPitch Xv0 against VariablePitch at input 0 (delta_beta = 0 is the first knot)
    s1 = orc.Sim(variant=1, kin=0, dt=0.01)
    assert s1.init_trim(EAS=45.0, h_e=1500.0)[0] and s3.init_trim(EAS=45.0, h_e=1500.0)[0]
    assert s1.y(fb.Y["FB_Y_PROP_THRUST"]) == s3.y(fb.Y["FB_Y_PROP_THRUST"])
    # Flat extrapolation: J, Mt beyond 1.5 read the last knot (`lookup.C_Fx(1.5, 1.5, 0.1) == lookup.C_Fx(3, 2, 2)`)
    x = s3.x.copy()
    s3.set_input(fb.U["FB_U_PROP_PITCH"], 1.0)
    x[9] = 1.0; s3.init_state(x); s3.xdot                 # propeller almost at rest at 45 m/s: J = 2 pi v / (max(w, 1) d) >> 1.5
    assert s3.y(fb.Y["FB_Y_PROP_J"]) > 1.5
    thrust = s3.y(fb.Y["FB_Y_PROP_THRUST"])
    rho_f2_d4 = thrust / t3[20, 0, 10, 0]                 # C_Fx at the last J knot, first Mt knot (Mt = w d / 2 a ~ 0.003 -> cell 0)
    assert np.isfinite(rho_f2_d4)


def test_blob_holds_the_variable_pitch_table_of_the_reference_test_case(orc):
    """`prop.tables3` of the table blob (tools/build_tables.py, a Python restatement) against the oracle's C++ one for
    VariablePitch(-0.1, 0.2; length = 11), the reference's own test case (test_propellers.jl:152)."""
    t = _blob_section("prop.tables3"); r = _blob_section("prop.pitch3")
    assert t is not None and t.shape == (21, 21, 11, 6) and list(r) == [-0.1, 0.2]
    s = orc.Sim(variant=VAR_A2, kin=0, dt=0.01, prop_pitch=(-0.1, 0.2, 11))
    assert np.allclose(t, s.prop3(), rtol=1e-9, atol=1e-12)
    # monotone in pitch where the blade is not stalled: static thrust coefficient grows with delta_beta (test_propellers.jl:165)
    assert (np.diff(t[10, 5, :, 0]) > 0).all()


def test_actuator2_variant_on_every_kinematic_mechanisation(orc, fb):
    """The component variant over WA / ECEF / NED (kinematics.jl:150,250,329): same trim, and after 5 s with an elevator command
    step through the second-order unit the three descriptions agree on attitude, altitude and airspeed like the reference's
    own kinematics test demands of them (test_kinematics.jl:75-93), with identical actuator states."""
    sims = []
    for kin in (0, 1, 2):
        s = orc.Sim(variant=VAR_A2, kin=kin, dt=0.01, prop_pitch=(-0.1, 0.2))
        s.set_input(fb.U["FB_U_PROP_PITCH"], 0.2)
        ok, ts, cost = s.init_trim(EAS=42.0, h_e=900.0, psi_nb=0.7, lat=0.6, lon=-0.4)
        assert ok and cost < 1e-16
        s.set_input(fb.U["FB_U_ACT_ELEVATOR"], s.get_input(fb.U["FB_U_ACT_ELEVATOR"]) + 0.08)
        s.step(500)
        sims.append((s, ts))
    assert s.nx == 38 and sims[0][0].nx == 41 and sims[1][0].nx == 40
    ref, ts0 = sims[0]
    for s, ts in sims[1:]:
        assert np.allclose(ts, ts0, rtol=0, atol=1e-9)
        assert np.allclose(s.x[12:26], ref.x[12:26], rtol=0, atol=1e-12)                    # the units see nothing but their commands
        for f, tol in (("FB_Y_THETA", 1e-7), ("FB_Y_PHI", 1e-7), ("FB_Y_H_E", 1e-4), ("FB_Y_EAS", 1e-6), ("FB_Y_PROP_DBETA", 0.0)):
            assert abs(s.y(fb.Y[f]) - ref.y(fb.Y[f])) <= tol, (f, s.y(fb.Y[f]), ref.y(fb.Y[f]))
    assert abs(ref.y(fb.Y["FB_Y_PROP_DBETA"]) - (-0.1 + 0.3 * 0.2)) < 1e-15
