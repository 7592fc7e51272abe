"""GPU parity tests for the fly-by-wire variants with the gain-scheduled control laws in the loop
(Cessna172Xv1 / Xv2: lib/FlightApps/src/c172/c172x/control/c172x_ctl.jl, c172x1.jl, c172x2.jl).
The CUDA path (through the C ABI) is compared with the CPU oracle member by member on the same seeded inputs:
continuous state and commands <= 1e-9 scaled error, control modes / altitude-tracking state / stall / engine
state bit-exact at every comparison instant."""
import numpy as np
import pytest

from parity_util import scaled_err, state_floors

pytestmark = pytest.mark.gpu

SEED = 20261018
LON = dict(direct=0, sas=1, thr_q=2, thr_theta=3, thr_EAS=4, EAS_q=5, EAS_theta=6, EAS_clm=7, EAS_alt=8)
LAT = dict(direct=0, sas=1, p_beta=2, phi_beta=3, chi_beta=4)
CTL_INPUTS = ["lon_mode_req", "lon_throttle_axis", "lon_throttle_offset", "lon_elevator_axis", "lon_elevator_offset", "lon_q_ref",
              "lon_theta_ref", "lon_eas_ref", "lon_clm_ref", "lon_h_ref", "lat_mode_req", "lat_aileron_axis", "lat_aileron_offset",
              "lat_rudder_axis", "lat_rudder_offset", "lat_p_ref", "lat_beta_ref", "lat_phi_ref", "lat_chi_ref"]
CTL_OUTPUTS = ["lon_mode", "lat_mode", "alt_state", "throttle_cmd", "elevator_cmd", "aileron_cmd", "rudder_cmd"]


def cfg3_params(n, rng):
    import bench
    lat = np.deg2rad(rng.uniform(-60, 60, n)); lon = rng.uniform(-np.pi, np.pi, n)
    return dict(EAS=rng.uniform(40, 55, n), h_e=rng.uniform(50, 3050, n) + bench.geoid_height(lat, lon), psi_nb=rng.uniform(-np.pi, np.pi, n),
                lat=lat, lon=lon, fuel_load=rng.uniform(0.1, 1, n), m_pilot=rng.uniform(0, 100, n), m_copilot=rng.uniform(0, 100, n),
                m_lpass=rng.uniform(0, 100, n), m_rpass=rng.uniform(0, 100, n), m_baggage=rng.uniform(0, 100, n))


def make_pair(fb, orc, variant, kin, n, p, idx, dt, Dt, wind=None, **simkw):
    ac = {2: fb.Cessna172Xv1, 3: fb.Cessna172Xv2}[variant]([fb.WA, fb.ECEF, fb.NED][kin]())
    sim = fb.BatchSimulation(fb.SimpleWorld(ac), n=n, dt=dt, Dt=Dt, t_end=1000.0, **simkw)
    oracles = []
    for i in idx:
        o = orc.Sim(variant=variant, kin=kin, dt=dt, Dt=Dt)
        oracles.append(o)
    if wind is not None:
        for k, f in enumerate(("wind_n", "wind_e")):
            sim.set_input(f, wind[k])
            for o, i in zip(oracles, idx):
                o.set_input(fb.INPUT_NAMES[f], wind[k][i])
    ok, ts, cost = sim.init(fb.C172.TrimParameters(**p))
    assert ok.all()
    for o, i in zip(oracles, idx):
        ok_o, _, _ = o.init_trim(**{k: float(v[i]) for k, v in p.items()})
        assert ok_o
    return sim, oracles


def compare(fb, sim, oracles, idx, variant, kin, tol=1e-9):
    xg = sim.x
    stg, eng = sim.discrete()
    fl = state_floors(sim.nx, 1, kin)
    outs = {k: sim.get_output(k) for k in CTL_OUTPUTS}
    worst = 0.0
    for o, i in zip(oracles, idx):
        err = scaled_err(xg[:, i], o.x, fl)
        worst = max(worst, float(err.max()))
        assert err.max() < tol, (i, int(err.argmax()), float(err.max()))
        assert o.discrete == (int(stg[i]), int(eng[i]))
        for k in ("lon_mode", "lat_mode", "alt_state"):
            assert outs[k][i] == o.y(fb.OUTPUT_NAMES[k]), (k, i, outs[k][i], o.y(fb.OUTPUT_NAMES[k]))
        for k in ("throttle_cmd", "elevator_cmd", "aileron_cmd", "rudder_cmd"):
            a, b = outs[k][i], o.y(fb.OUTPUT_NAMES[k])
            assert abs(a - b) <= tol * max(abs(b), 1e-2), (k, i, a, b)
    return worst


@pytest.mark.parametrize("variant,kin", [(2, 0), (3, 0), (2, 2)])
def test_avionics_init_matches_oracle(fb, orc, variant, kin):
    """f_init!(avionics, vehicle) after the trim: inputs made consistent with the vehicle, LQR trackers run once."""
    n = 32
    rng = np.random.default_rng(SEED + 10)
    p = cfg3_params(n, rng)
    idx = list(range(0, n, 4))
    sim, oracles = make_pair(fb, orc, variant, kin, n, p, idx, 0.01, 0.01)
    ins = {k: sim.get_input(k) for k in CTL_INPUTS}
    for o, i in zip(oracles, idx):
        for k in CTL_INPUTS:
            a, b = ins[k][i], o.get_input(fb.INPUT_NAMES[k])
            assert abs(a - b) <= 1e-9 * max(abs(b), 1e-2), (k, i, a, b)
    compare(fb, sim, oracles, idx, variant, kin)
    # the first derivative after init: assign!(vehicle, avionics) then f_ode!
    xd = sim.xdot
    fl = state_floors(sim.nx, 1, kin)
    for o, i in zip(oracles, idx):
        o.rhs()
        err = np.abs(xd[:, i] - o.xdot) / np.maximum(np.abs(o.xdot), fl)
        assert err.max() < 1e-7, (i, int(err.argmax()), float(err.max()))   # near-zero trim residuals: absolute floor dominates
    sim.close()


def test_cfg3_turning_climb_parity(fb, orc):
    """BASELINE.json configs[2] at test size: Cessna172Xv2/WA, EAS_clm + φ_β, random constant wind, dt = Δt = 0.01."""
    n, steps = 128, 1500
    rng = np.random.default_rng(SEED + 11)
    p = cfg3_params(n, rng)
    wind = (rng.uniform(-10, 10, n), rng.uniform(-10, 10, n))
    clm = rng.uniform(0.5, 3, n); phi = np.deg2rad(rng.uniform(-30, 30, n))
    idx = list(range(0, n, 8))
    sim, oracles = make_pair(fb, orc, 3, 0, n, p, idx, 0.01, 0.01, wind=wind, steps_per_launch=100)
    for name, val in (("lon_mode_req", np.full(n, float(LON["EAS_clm"]))), ("lat_mode_req", np.full(n, float(LAT["phi_beta"]))),
                      ("lon_clm_ref", clm), ("lat_phi_ref", phi)):
        sim.set_input(name, val)
        for o, i in zip(oracles, idx):
            o.set_input(fb.INPUT_NAMES[name], val[i])
    worst = 0.0
    for chunk in (1, 99, 400, 1000):       # compare right after the mode switch, then along the climb
        sim.step(chunk * 0.01)
        for o in oracles:
            assert o.step(chunk) == 0
        assert (sim.status() == 0).all()
        worst = max(worst, compare(fb, sim, oracles, idx, 3, 0))
    assert (sim.get_output("lon_mode") == LON["EAS_clm"]).all() and (sim.get_output("lat_mode") == LAT["phi_beta"]).all()
    # the climb is actually being flown: climb rate within 0.2 m/s of its reference after 15 s
    assert np.max(np.abs(-sim.get_output("v_d") - clm)) < 0.2
    assert sim.counters()["rhs_evals"] <= 4.2 * n * steps   # commands are patched into k1, not re-evaluated
    print("cfg3 worst scaled state error", worst)
    sim.close()


def test_mode_sequence_parity(fb, orc):
    """Every longitudinal / lateral mode, switched by scheduled inputs (bumpless-transfer resets, altitude acquire/hold
    hysteresis, χ tracking through ±π), Δt = 2 dt so that the periodic update skips every other step."""
    n, dt, Dt = 64, 0.01, 0.02
    rng = np.random.default_rng(SEED + 12)
    p = cfg3_params(n, rng)
    idx = list(range(0, n, 4))
    sim, oracles = make_pair(fb, orc, 2, 0, n, p, idx, dt, Dt, steps_per_launch=37)
    U = fb.U
    h0 = sim.get_input("lon_h_ref"); eas0 = sim.get_input("lon_eas_ref"); chi0 = sim.get_input("lat_chi_ref")
    ev = []
    for i in range(n):
        s = 10
        for lon_mode, lat_mode in (("sas", "sas"), ("thr_q", "p_beta"), ("thr_theta", "phi_beta"), ("thr_EAS", "chi_beta"),
                                   ("EAS_q", "sas"), ("EAS_theta", "p_beta"), ("EAS_clm", "chi_beta"), ("EAS_alt", "phi_beta"),
                                   ("direct", "direct"), ("EAS_alt", "chi_beta")):
            ev += [(i, s, U["FB_U_LON_MODE_REQ"], LON[lon_mode]), (i, s + 3, U["FB_U_LAT_MODE_REQ"], LAT[lat_mode])]
            s += 150
        ev += [(i, 200, U["FB_U_LON_Q_REF"], 0.01), (i, 260, U["FB_U_LON_Q_REF"], 0.0), (i, 200, U["FB_U_LAT_P_REF"], 0.02),
               (i, 250, U["FB_U_LAT_P_REF"], -0.02), (i, 300, U["FB_U_LAT_P_REF"], 0.0),
               (i, 330, U["FB_U_LON_THETA_REF"], 0.08), (i, 340, U["FB_U_LAT_PHI_REF"], 0.3),
               (i, 470, U["FB_U_LON_EAS_REF"], eas0[i] - 3.0), (i, 480, U["FB_U_LAT_CHI_REF"], chi0[i] + 2.5),
               (i, 500, U["FB_U_LON_THROTTLE_AXIS"], 0.9), (i, 930, U["FB_U_LON_CLM_REF"], 2.0),
               (i, 1080, U["FB_U_LON_H_REF"], h0[i] + 40.0), (i, 1090, U["FB_U_LAT_BETA_REF"], 0.03),
               (i, 1400, U["FB_U_LON_H_REF"], h0[i] + 25.0)]
    m, s, f, v = zip(*ev)
    sim.schedule_inputs(m, s, f, v)
    for o, i in zip(oracles, idx):
        for (mi, si, fi, vi) in ev:
            if mi == i:
                o.add_event(si, fi, vi)
    worst = 0.0
    seen_lon, seen_lat, seen_alt = set(), set(), set()
    t = 0
    for chunk in [11, 3, 100] + [50] * 28:
        sim.step(chunk * dt)
        t += chunk
        for o in oracles:
            assert o.step(chunk) == 0
        assert (sim.status() == 0).all()
        worst = max(worst, compare(fb, sim, oracles, idx, 2, 0))
        seen_lon |= set(sim.get_output("lon_mode").astype(int)); seen_lat |= set(sim.get_output("lat_mode").astype(int))
        seen_alt |= set(sim.get_output("alt_state").astype(int))
    assert seen_lon == set(range(9)) and seen_lat == set(range(5)) and seen_alt == {0, 1}, (seen_lon, seen_lat, seen_alt)
    print("mode sequence worst scaled state error", worst)
    sim.close()


def test_control_laws_are_overridden_on_the_ground(fb, orc):
    """test_c172x1.jl:55-86: with weight on wheels every requested mode collapses to direct."""
    import bench
    n = 8
    h = 1.9 + float(bench.geoid_height(0.0, 0.0))
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Xv1(fb.WA())), n=n, dt=0.01, t_end=10.0)
    sim.init(fb.C172.Init(h_e=h))
    o = orc.Sim(variant=2, kin=0, dt=0.01, Dt=0.01)
    o.init(h_e=h)
    for k, v in (("lon_mode_req", LON["EAS_clm"]), ("lat_mode_req", LAT["p_beta"]), ("lon_throttle_axis", 0.1),
                 ("lon_elevator_axis", 0.3), ("lat_aileron_axis", 0.2), ("lat_rudder_axis", 0.4)):
        sim.set_input(k, float(v))
        o.set_input(fb.INPUT_NAMES[k], v)
    sim.step(0.01)
    assert o.step(1) == 0
    assert (sim.get_output("wow_left") == 1).all() or (sim.get_output("wow_nose") == 1).all()
    assert (sim.get_output("lon_mode") == 0).all() and (sim.get_output("lat_mode") == 0).all()
    assert (sim.get_input("act_throttle") == 0.1).all() and (sim.get_input("act_elevator") == 0.3).all()
    assert (sim.get_input("act_rudder") == 0.4).all() and (sim.get_input("act_aileron") == 0.2).all()
    sim.step(1.0); assert o.step(100) == 0
    # Ground contact is ill-conditioned in the reference's own formulation: the strut compression is the difference of
    # two ECEF position vectors of magnitude 6.4e6 m (landinggear.jl:262-270), so one ulp (9.3e-10 m) of rounding noise
    # is a relative 1e-7..1e-6 of a millimetre-scale compression (here the aircraft is dropped with the wheels just
    # touching). Two correct implementations therefore agree to ~1e-6 on the gear-driven states, not to 1e-9. (Past
    # ~1.1 s this member sits at rest with wheel speeds hovering around the reference's 1e-3 m/s friction-direction
    # switch (landinggear.jl:436-452), a discontinuity that turns that 1e-6 into O(1) differences: see DESIGN.md §6.)
    fl = state_floors(sim.nx, 1, 0)
    err = scaled_err(sim.x[:, 0], o.x, fl)
    assert err.max() < 1e-5, (int(err.argmax()), float(err.max()))
    assert [sim.get_output(k)[0] for k in ("wow_left", "wow_right", "wow_nose")] == [o.y(fb.OUTPUT_NAMES[k]) for k in ("wow_left", "wow_right", "wow_nose")]
    sim.close()
