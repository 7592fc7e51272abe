"""GPU-vs-oracle parity on the branches round 1 left untested (VERDICT r1 "What's weak" #6): the ISA layers above the
tropopause and the airborne -> complete hand-over they trigger (atmosphere.jl:99-135), non-standard sea-level temperature
and pressure over their Ranged bounds (atmosphere.jl:67-84), a vertical wind component (atmosphere.jl:165), fuel exhaustion
and the engine stop input (piston.jl:428-453, c172.jl:607-641), the altitude bound (geodesy.jl:218-221) and the fly-by-wire
control laws on ECEF kinematics (kinematics.jl:250-320)."""
import numpy as np
import pytest

from parity_util import scaled_err, state_floors

pytestmark = pytest.mark.gpu
SEED = 20261018


def _check(fb, sim, oracles, variant, kin, tol=1e-9):
    xg = sim.x
    stg, eng = sim.discrete()
    status = sim.status()
    fl = state_floors(sim.nx, variant, kin)
    worst = 0.0
    for i, o in enumerate(oracles):
        assert status[i] == o.status, (i, status[i], o.status)
        assert o.discrete == (int(stg[i]), int(eng[i])), (i, o.discrete, stg[i], eng[i])
        err = scaled_err(xg[:, i], o.x, fl)
        worst = max(worst, float(err.max()))
        assert err.max() < tol, (i, int(err.argmax()), float(err.max()))
    return worst


@pytest.mark.parametrize("variant", [0, 2])
def test_tropopause_crossing_and_upper_isa_layers(fb, orc, variant):
    """Members launched at 9-13 km (and two at 20 km) climbing or descending through 11 km / 20 km: the airborne kernel
    carries the troposphere only and hands a member to the complete kernel above it; T, p, the engine's density ratio and
    every state must follow the oracle through the layer change. (The complete kernel anchors its ISA pressure on the
    step-start state, a tropospheric formula: a step that straddles 11 km must not reuse an anchor across the layer change —
    this test caught exactly that when the complete C172S kernel got the anchored shape.) Cessna172Sv0 and Cessna172Xv1."""
    h = np.array([9000.0, 10900.0, 10960.0, 10995.0, 11005.0, 11050.0, 11200.0, 12000.0, 13000.0, 19990.0, 20030.0])
    vD = np.array([-15.0, -60.0, -40.0, -20.0, 15.0, 20.0, 25.0, 5.0, -5.0, -30.0, 20.0])
    n = h.size
    kw = dict(h_e=h, vN=np.full(n, 60.0), vD=vD, n_eng=np.full(n, 0.85), engine_state=np.full(n, 2.0), throttle=np.full(n, 1.0), mixture=0.5)
    ac = {0: fb.Cessna172Sv0, 2: fb.Cessna172Xv1}[variant]
    sim = fb.BatchSimulation(fb.SimpleWorld(ac(fb.WA())), n=n, dt=0.01, Dt=0.01, t_end=100.0, steps_per_launch=40)
    sim.init(fb.C172.Init(**kw))
    oracles = []
    for i in range(n):
        o = orc.Sim(variant=variant, kin=0, dt=0.01, Dt=0.01)
        o.init(**{k: float(np.broadcast_to(v, (n,))[i]) for k, v in kw.items()})
        oracles.append(o)
    crossed = np.zeros(n, dtype=bool)
    h0 = sim.get_output("h_o")
    for _ in range(5):
        sim.step(2.0)
        for o in oracles:
            o.step(200)
        worst = _check(fb, sim, oracles, variant, 0)
        h1 = sim.get_output("h_o")
        for lvl in (11000.0, 20000.0):
            crossed |= (np.minimum(h0, h1) < lvl) & (np.maximum(h0, h1) > lvl)
    assert crossed.sum() >= 7
    for f in ("tas", "eas", "mach", "eng_map", "eng_p_shaft"):
        a = sim.get_output(f); b = np.array([o.y(fb.OUTPUT_NAMES[f]) for o in oracles])
        assert np.all(np.abs(a - b) <= 1e-9 * np.maximum(np.abs(b), 1.0)), f
    print("upper-atmosphere worst scaled state error", worst)
    sim.close()


@pytest.mark.parametrize("variant,kin", [(0, 0), (2, 0)])
def test_sea_level_conditions_and_vertical_wind(fb, orc, variant, kin):
    """T_sl and p_sl swept over their Ranged bounds (+-50 K, +-10 kPa; assignments beyond them saturate), plus a wind with a
    vertical component: trim and 8 s of flight against the oracle."""
    n = 24
    rng = np.random.default_rng(SEED + 70)
    T_sl = 288.15 + np.concatenate([np.linspace(-50, 50, n - 4), [-80.0, 80.0, 0.0, 0.0]])
    p_sl = 101325.0 + np.concatenate([rng.uniform(-10000, 10000, n - 4), [0.0, 0.0, -15000.0, 15000.0]])
    wind = rng.uniform(-8, 8, (3, n)); wind[2] = rng.uniform(-3, 3, n)
    EAS, h = rng.uniform(38, 52, n), rng.uniform(300, 2500, n)
    ac = {0: fb.Cessna172Sv0, 2: fb.Cessna172Xv1}[variant]
    sim = fb.BatchSimulation(fb.SimpleWorld(ac(fb.WA())), n=n, dt=0.01, Dt=0.01, t_end=100.0)
    U = fb.U
    fields = [("t_sl", T_sl), ("p_sl", p_sl), ("wind_n", wind[0]), ("wind_e", wind[1]), ("wind_d", wind[2])]
    for f, v in fields:
        sim.set_input(f, v)
    assert sim.get_input("t_sl").min() == 288.15 - 50 and sim.get_input("p_sl").max() == 101325.0 + 10000     # Ranged saturation
    ok, _, _ = sim.init(fb.C172.TrimParameters(EAS=EAS, h_e=h))
    oracles = []
    for i in range(n):
        o = orc.Sim(variant=variant, kin=kin, dt=0.01, Dt=0.01)
        for f, v in fields:
            o.set_input(fb.INPUT_NAMES[f], v[i])
        ok_o, _, _ = o.init_trim(EAS=EAS[i], h_e=h[i])
        assert ok_o == bool(ok[i])
        oracles.append(o)
    assert ok.sum() >= n - 4
    if variant == 2:
        sim.set_input("lon_mode_req", 7.0); sim.set_input("lat_mode_req", 3.0); sim.set_input("lon_clm_ref", 1.0); sim.set_input("lat_phi_ref", 0.2)
        for o in oracles:
            o.set_input(U["FB_U_LON_MODE_REQ"], 7.0); o.set_input(U["FB_U_LAT_MODE_REQ"], 3.0)
            o.set_input(U["FB_U_LON_CLM_REF"], 1.0); o.set_input(U["FB_U_LAT_PHI_REF"], 0.2)
    else:
        e0 = sim.get_input("elevator")
        sim.set_input("elevator", e0 + 0.05)
        for i, o in enumerate(oracles):
            o.set_input(U["FB_U_ELEVATOR"], e0[i] + 0.05)
    for _ in range(4):
        sim.step(2.0)
        for o in oracles:
            o.step(200)
        worst = _check(fb, sim, oracles, 1 if variant else 0, kin)
    print("sea-level / wind sweep worst scaled state error", worst)
    sim.close()


def test_fuel_exhaustion_and_engine_stop(fb, orc):
    """Half the members start with a few grams of usable fuel and run dry within seconds (fuel_available false -> engine off,
    piston.jl:428-453); the others get the stop input at t = 1.5 s. Engine state bit-exact, windmilling decay <= 1e-9."""
    n = 16
    rng = np.random.default_rng(SEED + 71)
    fuel = np.where(np.arange(n) % 2 == 0, rng.uniform(2e-5, 2e-4, n), 0.5)
    EAS, h = rng.uniform(40, 52, n), rng.uniform(500, 2000, n)
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=0.01, t_end=100.0, steps_per_launch=70)
    ok, _, _ = sim.init(fb.C172.TrimParameters(EAS=EAS, h_e=h, fuel_load=fuel))
    assert ok.all()
    stop_f = fb.U["FB_U_ENG_STOP"]
    odd = np.arange(1, n, 2)
    sim.schedule_inputs(odd, np.full(odd.size, 150), np.full(odd.size, stop_f), np.ones(odd.size))
    oracles = []
    for i in range(n):
        o = orc.Sim(variant=0, kin=0, dt=0.01)
        ok_o, _, _ = o.init_trim(EAS=EAS[i], h_e=h[i], fuel_load=fuel[i])
        assert ok_o
        if i % 2:
            o.add_event(150, stop_f, 1.0)
        oracles.append(o)
    seen = []
    for _ in range(6):
        sim.step(1.0)
        for o in oracles:
            o.step(100)
        worst = _check(fb, sim, oracles, 0, 0)
        seen.append(sim.discrete()[1].copy())
    assert (seen[0][odd] == fb.FB_ENG_RUNNING).all() and (seen[-1][odd] == fb.FB_ENG_OFF).all()
    assert (seen[-1][::2] == fb.FB_ENG_OFF).all() and any((s[::2] == fb.FB_ENG_RUNNING).any() for s in seen)   # ran dry on the way
    assert (sim.get_output("fuel_m_total")[::2] <= 1.0 + 1e-9).all()
    print("fuel / stop worst scaled state error", worst)
    sim.close()


def test_altitude_bound_terminates_like_the_reference(fb, orc):
    """Altitude{D}(h) refuses h < -1000 m (geodesy.jl:218-221). Over valid terrain (elevation >= -1000 m) a flying aircraft cannot
    cross the bound before it touches the ground, so the bound only meets bad initial conditions: a member initialised with an
    orthometric altitude below it (h_e = -985 m under a +17 m geoid) is frozen with FB_ST_BOUNDS where the reference's first
    f_ode! would have thrown; a member diving into terrain at -990 m ends as a GroundCrash; their neighbours keep flying."""
    n = 4
    h = np.array([-950.0, -900.0, 500.0, -985.0])
    vD = np.array([40.0, 10.0, 0.0, 30.0])
    kw = dict(h_e=h, vN=np.full(n, 50.0), vD=vD, n_eng=np.full(n, 0.8), engine_state=np.full(n, 2.0), throttle=np.full(n, 0.5))
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=0.01, t_end=100.0, steps_per_launch=30)
    sim.set_input("trn_elevation", -990.0)
    sim.init(fb.C172.Init(**kw))
    oracles = []
    for i in range(n):
        o = orc.Sim(variant=0, kin=0, dt=0.01)
        o.set_input(fb.U["FB_U_TRN_ELEVATION"], -990.0)
        o.init(**{k: float(v[i]) for k, v in kw.items()})
        oracles.append(o)
    hist = []
    for _ in range(8):
        sim.step(0.5)
        for o in oracles:
            o.step(50)
        st = sim.status(); so = np.array([o.status for o in oracles])
        assert np.array_equal(st, so), (st, so)
        hist.append(st.copy())
        xg = sim.x
        for i, o in enumerate(oracles):
            assert np.abs(xg[:, i] - o.x).max() <= 1e-9 * max(1.0, np.abs(o.x).max()), i     # frozen members stay where they stopped
    assert hist[-1][2] == fb.FB_ST_RUNNING and (hist[-1][[0, 3]] != fb.FB_ST_RUNNING).all()
    assert hist[-1][3] == fb.FB_ST_BOUNDS and hist[-1][0] == fb.FB_ST_CRASH_RATE and hist[0][0] == fb.FB_ST_RUNNING
    sim.close()


def test_fly_by_wire_control_laws_on_ecef_kinematics(fb, orc):
    """Cessna172Xv1 over the ECEF mechanisation (q_eb, n_e, h_e state; kinematics.jl:250-320) with the autopilot in
    EAS_clm / phi_beta: 12 s closed loop against the oracle, modes bit-exact."""
    n = 16
    rng = np.random.default_rng(SEED + 72)
    EAS, h = rng.uniform(40, 55, n), rng.uniform(300, 2800, n)
    lat, lon, psi = np.deg2rad(rng.uniform(-70, 70, n)), rng.uniform(-3.1, 3.1, n), rng.uniform(-3.1, 3.1, n)
    clm, phi = rng.uniform(0.5, 2.5, n), np.deg2rad(rng.uniform(-25, 25, n))
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Xv1(fb.ECEF())), n=n, dt=0.01, Dt=0.01, t_end=100.0)
    ok, _, _ = sim.init(fb.C172.TrimParameters(EAS=EAS, h_e=h, lat=lat, lon=lon, psi_nb=psi))
    assert ok.all()
    sim.set_input("lon_mode_req", 7.0); sim.set_input("lat_mode_req", 3.0); sim.set_input("lon_clm_ref", clm); sim.set_input("lat_phi_ref", phi)
    U = fb.U
    oracles = []
    for i in range(n):
        o = orc.Sim(variant=2, kin=1, dt=0.01, Dt=0.01)
        ok_o, _, _ = o.init_trim(EAS=EAS[i], h_e=h[i], lat=lat[i], lon=lon[i], psi_nb=psi[i])
        assert ok_o
        o.set_input(U["FB_U_LON_MODE_REQ"], 7.0); o.set_input(U["FB_U_LAT_MODE_REQ"], 3.0)
        o.set_input(U["FB_U_LON_CLM_REF"], clm[i]); o.set_input(U["FB_U_LAT_PHI_REF"], phi[i])
        oracles.append(o)
    for _ in range(4):
        sim.step(3.0)
        for o in oracles:
            o.step(300)
        worst = _check(fb, sim, oracles, 1, 1)
        for k in ("lon_mode", "lat_mode", "alt_state"):
            a = sim.get_output(k); b = np.array([o.y(fb.OUTPUT_NAMES[k]) for o in oracles])
            assert np.array_equal(a, b), k
    assert (sim.get_output("lon_mode") == 7).all() and (sim.get_output("lat_mode") == 3).all()
    print("C172X1 / ECEF worst scaled state error", worst)
    sim.close()
