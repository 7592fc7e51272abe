"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bars: continuous state <= 1e-9 scaled error (see parity_util.state_floors); Bool/enum state bit-exact."""
import numpy as np
import pytest

from parity_util import scaled_err, state_floors

pytestmark = pytest.mark.gpu

SEED = 20261018


def cfg2_params(n, rng):
    import bench
    EAS = rng.uniform(30, 55, n)
    lat = np.deg2rad(rng.uniform(-60, 60, n)); lon = rng.uniform(-np.pi, np.pi, n)
    h_e = rng.uniform(50, 3050, n) + bench.geoid_height(lat, lon)   # 50..3050 m above mean sea level
    flaps = rng.choice([0.0, 1 / 3, 2 / 3, 1.0], n)
    flaps[EAS < 35] = 1.0
    flaps[EAS > 42] = 0.0   # full flap at high speed cannot be trimmed level with the available power
    return dict(EAS=EAS, h_e=h_e, flaps=flaps, psi_nb=rng.uniform(-np.pi, np.pi, n),
                lat=lat, lon=lon, fuel_load=rng.uniform(0.1, 1, n),
                m_pilot=rng.uniform(0, 100, n), m_copilot=rng.uniform(0, 100, n), m_lpass=rng.uniform(0, 100, n),
                m_rpass=rng.uniform(0, 100, n), m_baggage=rng.uniform(0, 100, n))


COPY_FIELDS = ["aileron", "elevator", "rudder", "aileron_offset", "elevator_offset", "rudder_offset", "flaps", "brake_left",
               "brake_right", "throttle", "mixture", "eng_start", "eng_stop", "mixture_auto", "steering_engaged", "m_pilot",
               "m_copilot", "m_lpass", "m_rpass", "m_baggage", "t_sl", "p_sl", "wind_n", "wind_e", "wind_d", "trn_elevation",
               "trn_surface", "act_throttle", "act_aileron", "act_elevator", "act_rudder", "act_flaps", "act_brake_left",
               "act_brake_right"]


def oracle_from_gpu(fb, orc, sim, variant, kin, dt, idx, Dt=None):
    """Oracle members started from the GPU ensemble's CURRENT state and inputs (isolates the stepping path)."""
    x = sim.x
    st, en = sim.discrete()
    inputs = {f: sim.get_input(f) for f in COPY_FIELDS}
    out = []
    for i in idx:
        o = orc.Sim(variant=variant, kin=kin, dt=dt, Dt=Dt)
        for f, v in inputs.items():
            o.set_input(fb.INPUT_NAMES[f], v[i])
        o.init_state(x[:, i], stall=int(st[i]), eng_state=int(en[i]))
        out.append(o)
    return out


def oracle_members(orc, variant, kin, dt, params, idx, Dt=None):
    sims = []
    for i in idx:
        o = orc.Sim(variant=variant, kin=kin, dt=dt, Dt=Dt)
        ok, ts, cost = o.init_trim(**{k: float(np.asarray(v)[i]) if np.ndim(v) else float(v) for k, v in params.items()})
        assert ok
        sims.append(o)
    return sims


@pytest.mark.parametrize("variant,kin", [(0, 0), (0, 2), (1, 0), (1, 2), (0, 1), (1, 1)])
def test_trim_and_rhs_parity(fb, orc, variant, kin):
    n = 64
    rng = np.random.default_rng(SEED)
    p = cfg2_params(n, rng)
    ac = [fb.Cessna172Sv0, fb.Cessna172Xv0][variant]([fb.WA, fb.ECEF, fb.NED][kin]())
    sim = fb.BatchSimulation(fb.SimpleWorld(ac), n=n, dt=0.01, t_end=10.0)
    ok, ts, cost = sim.init(fb.C172.TrimParameters(**p))
    assert ok.all()
    assert cost.max() <= 1e-16
    xg, xdg = sim.x, sim.xdot
    fl = state_floors(sim.nx, variant, kin)
    for i in range(0, n, 4):
        o = oracle_members(orc, variant, kin, 0.01, p, [i])[0]
        # trim states agree (both converge far below the reference optimiser's 1e-16 stop value)
        assert np.max(scaled_err(xg[:, i], o.x, fl)) < 1e-9
        # RHS at the SAME state: feed the GPU state to the oracle
        st, en = o.discrete
        o.init_state(xg[:, i], stall=st, eng_state=en)
        xdo = o.xdot
        err = np.abs(xdg[:, i] - xdo) / np.maximum(np.abs(xdo), fl)
        assert err.max() < 1e-9, (i, err.argmax(), err.max())
    sim.close()


@pytest.mark.parametrize("variant,kin,steps", [(0, 0, 1000), (0, 2, 1000), (1, 0, 400), (0, 1, 1000)])
def test_step_parity_with_doublets(fb, orc, variant, kin, steps):
    n = 96
    dt = 0.01
    rng = np.random.default_rng(SEED + 1)
    p = cfg2_params(n, rng)
    ac = [fb.Cessna172Sv0, fb.Cessna172Xv0][variant]([fb.WA, fb.ECEF, fb.NED][kin]())
    sim = fb.BatchSimulation(fb.SimpleWorld(ac), n=n, dt=dt, t_end=100.0, steps_per_launch=64)
    ok, ts, cost = sim.init(fb.C172.TrimParameters(**p))
    assert ok.all()
    ele_f = fb.U["FB_U_ELEVATOR"] if variant == 0 else fb.U["FB_U_ACT_ELEVATOR"]
    ail_f = fb.U["FB_U_AILERON"] if variant == 0 else fb.U["FB_U_ACT_AILERON"]
    ele0, ail0 = sim.get_input(ele_f), sim.get_input(ail_f)
    # elevator doublet +-0.1 on [1,2) s and [2,3) s, aileron doublet on [4,5), [5,6)
    ev = []
    for i in range(n):
        ev += [(i, 100, ele_f, ele0[i] + 0.1), (i, 200, ele_f, ele0[i] - 0.1), (i, 300, ele_f, ele0[i]),
               (i, 400, ail_f, ail0[i] + 0.1), (i, 500, ail_f, ail0[i] - 0.1), (i, 600, ail_f, ail0[i])]
    m, s, f, v = zip(*ev)
    sim.schedule_inputs(m, s, f, v)
    idx = list(range(0, n, 8))
    oracles = oracle_from_gpu(fb, orc, sim, variant, kin, dt, idx)
    sim.step(steps * dt)
    xg = sim.x
    stg, eng = sim.discrete()
    assert (sim.status() == 0).all()
    fl = state_floors(sim.nx, variant, kin)
    worst = 0.0
    for i, o in zip(idx, oracles):
        for (mi, si, fi, vi) in ev:
            if mi == i:
                o.add_event(si, fi, vi)
        o.step(steps)
        assert o.status == 0
        err = scaled_err(xg[:, i], o.x, fl)
        worst = max(worst, err.max())
        assert err.max() < 1e-9, (i, int(err.argmax()), float(err.max()))
        st, en = o.discrete
        assert (st, en) == (int(stg[i]), int(eng[i]))
    print("worst scaled state error", worst)
    sim.close()


def test_infeasible_trim_is_flagged_like_the_oracle(fb, orc):
    # full flaps at 55 m/s EAS needs more power than the engine has: init! reports failure (c172.jl:934-940)
    n = 8
    EAS = np.array([54.9, 50.0, 54.2, 45.0, 55.0, 40.0, 35.0, 30.0])
    flaps = np.array([1.0, 0.0, 1.0, 0.0, 1.0, 0.0, 0.5, 1.0])
    h = np.full(n, 2700.0)
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=0.01, t_end=1.0)
    ok, ts, cost = sim.init(fb.C172.TrimParameters(EAS=EAS, flaps=flaps, h_e=h))
    st = sim.status()
    for i in range(n):
        o = orc.Sim(0, 0, 0.01)
        ok_o, _, _ = o.init_trim(EAS=EAS[i], flaps=flaps[i], h_e=h[i])
        assert bool(ok[i]) == ok_o
        assert (st[i] == fb.FB_ST_TRIM_FAILED) == (not ok_o)
    assert not ok.all() and ok.any()
    # the reference itself only warns and goes on from the best trim state found (c172.jl:934-940): release_trim_failed gives that
    x0 = sim.x.copy()
    sim.step(0.2)
    assert np.array_equal(sim.x[:, ~ok], x0[:, ~ok])                      # frozen until released
    assert sim.release_trim_failed() == int((~ok).sum()) and (sim.status() == 0).all()
    sim.step(0.2)
    x1 = sim.x
    assert (np.abs(x1[:, ~ok] - x0[:, ~ok]).max(axis=0) > 0).all() and np.isfinite(x1).all()
    sim.close()


def test_outputs_match_oracle(fb, orc):
    n = 32
    rng = np.random.default_rng(SEED + 2)
    p = cfg2_params(n, rng)
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=0.01, t_end=10.0)
    sim.init(fb.C172.TrimParameters(**p))
    sim.step(0.5)
    names = ["psi", "theta", "phi", "lat", "lon", "h_e", "h_o", "v_n", "v_e", "v_d", "v_gnd", "chi_gnd", "gamma_gnd", "p", "q", "r",
             "tas", "eas", "cas", "mach", "qbar", "alpha", "beta", "alpha_filt", "beta_filt", "stall", "c_l", "c_d",
             "wow_left", "wow_right", "wow_nose", "eng_omega", "eng_state", "eng_map", "eng_p_shaft", "eng_tau_shaft", "eng_mdot",
             "fuel_m_total", "prop_thrust", "prop_j", "mass", "f_x", "f_y", "f_z"]
    xg = sim.x
    stg, eng = sim.discrete()
    out = {k: sim.get_output(k) for k in names}
    for i in range(0, n, 4):
        o = orc.Sim(variant=0, kin=0, dt=0.01)
        o.init_trim(**{k: float(v[i]) for k, v in p.items()})
        o.want_y_extra(True)
        o.init_state(xg[:, i], stall=int(stg[i]), eng_state=int(eng[i]))
        for k in names:
            a, b = out[k][i], o.y(fb.OUTPUT_NAMES[k])
            assert abs(a - b) <= 1e-9 * max(abs(b), 1e-2), (k, a, b)
    sim.close()


def test_fstep_and_fperiodic_hooks_match_oracle(fb, orc):
    """The Modeling-protocol pieces in isolation: f_step!(world) on states that trip its branches
    (denormalised quaternions, stall hysteresis, engine start/stop) and f_periodic!(avionics) + assign! on a trimmed
    Cessna172Xv1, each against the oracle's callback bodies."""
    n = 64
    rng = np.random.default_rng(SEED + 90)
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=0.01, t_end=10.0)
    EAS = rng.uniform(35, 55, n)
    ok, _, _ = sim.init(fb.C172.TrimParameters(EAS=EAS, h_e=rng.uniform(500, 2500, n)))
    assert ok.all()
    x = sim.x.copy()
    L = sim.nx
    kinx, dyn = 12, L - 6
    x[kinx:kinx + 4] *= 1.0 + rng.choice([0.0, 3e-8, 1e-9], n)          # q_wb: some beyond the 1e-8 renormalisation threshold
    x[kinx + 4:kinx + 8] *= 1.0 - rng.choice([0.0, 5e-8], n)            # q_ew
    pitch_up = rng.uniform(0, 1, n) < 0.3                               # alpha > 0.36 -> stall sets
    x[dyn + 5] = np.where(pitch_up, x[dyn + 3] * 0.5, x[dyn + 5])
    x[9] = np.where(rng.uniform(0, 1, n) < 0.2, 20.0, x[9])             # engine below the stall speed -> off
    stall0 = (rng.uniform(0, 1, n) < 0.3).astype(np.uint8)
    eng0 = rng.choice([0, 1, 2], n).astype(np.uint8)
    start = (rng.uniform(0, 1, n) < 0.5).astype(float)
    sim.set_input("eng_start", start)
    sim.init_state(x, stall=stall0, eng_state=eng0)
    changed = sim.f_step()
    xg = sim.x; stg, eng = sim.discrete(); st = sim.status()
    inputs = {f: sim.get_input(f) for f in ("throttle", "elevator", "aileron", "rudder", "flaps", "mixture", "m_pilot", "m_copilot", "m_lpass", "m_rpass", "m_baggage")}
    n_changed = 0
    for i in range(0, n, 2):
        o = orc.Sim(variant=0, kin=0, dt=0.01)
        for f, v in inputs.items():
            o.set_input(fb.INPUT_NAMES[f], v[i])
        o.set_input(fb.U["FB_U_ENG_START"], start[i])
        o.init_state(x[:, i], stall=int(stall0[i]), eng_state=int(eng0[i]))
        assert o.f_step() == st[i]
        assert o.discrete == (int(stg[i]), int(eng[i])), i
        assert np.max(np.abs(xg[:, i] - o.x) / np.maximum(np.abs(o.x), 1e-3)) < 1e-12, i
        n_changed += int(changed[i])
    assert 0 < n_changed < n // 2 + 1 and set(eng.tolist()) == {0, 1, 2}
    sim.close()
    # ---- f_periodic!(avionics) + assign! in isolation, right after the trim (controller memory = f_init!)
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Xv1(fb.WA())), n=n, dt=0.01, Dt=0.01, t_end=10.0)
    h_e = rng.uniform(500, 2500, n)
    ok, _, _ = sim.init(fb.C172.TrimParameters(EAS=EAS, h_e=h_e))
    assert ok.all()
    lon_req = rng.choice([1, 4, 6, 7, 8], n).astype(float); lat_req = rng.choice([1, 2, 3, 4], n).astype(float)
    clm = rng.uniform(-2, 2, n); phi = rng.uniform(-0.4, 0.4, n)
    sim.set_input("lon_mode_req", lon_req); sim.set_input("lat_mode_req", lat_req)
    sim.set_input("lon_clm_ref", clm); sim.set_input("lat_phi_ref", phi)
    sim.f_periodic()
    names = ("act_throttle", "act_aileron", "act_elevator", "act_rudder")
    cmds = {k: sim.get_input(k) for k in names}
    lon_mode, lat_mode = sim.get_output("lon_mode"), sim.get_output("lat_mode")
    for i in range(0, n, 2):
        o = orc.Sim(variant=2, kin=0, dt=0.01, Dt=0.01)
        ok_o, _, _ = o.init_trim(EAS=float(EAS[i]), h_e=float(h_e[i]))
        assert ok_o
        for f, v in (("lon_mode_req", lon_req[i]), ("lat_mode_req", lat_req[i]), ("lon_clm_ref", clm[i]), ("lat_phi_ref", phi[i])):
            o.set_input(fb.INPUT_NAMES[f], float(v))
        assert o.f_periodic() == 0
        assert lon_mode[i] == o.y(fb.Y["FB_Y_LON_MODE"]) and lat_mode[i] == o.y(fb.Y["FB_Y_LAT_MODE"]), i
        for k in names:
            a, b = cmds[k][i], o.get_input(fb.INPUT_NAMES[k])
            assert abs(a - b) <= 1e-9 * max(abs(b), 1e-2), (k, i, a, b)
    sim.close()


def test_cfg1_single_member_elevator_step(fb, orc):
    """BASELINE.json configs[0], the reference's own CPU-runnable case (README example 3, `README.md:142-204`): ONE
    Cessna172Sv0 on NED kinematics, `C172.TrimParameters()` defaults, 1 s at trim, then `act.u.elevator += 0.1`, to t = 10 s at
    dt = 0.01 — an ensemble of one member (one thread of one warp), every logged sample of theta against the oracle, and the
    picture-level pins of the README plot (theta_trim ~ 0.0236 rad; pitch-up to ~0.2 rad within the 10 s)."""
    dt = 0.01
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.NED())), n=1, dt=dt, t_end=10.0, saveat=0.1, log_fields=("theta", "h_e", "eas"))
    ok, ts, cost = sim.init(fb.C172.TrimParameters())
    assert ok.all() and cost[0] <= 1e-16
    o = orc.Sim(variant=0, kin=2, dt=dt)
    ok_o, ts_o, _ = o.init_trim()
    assert ok_o and np.allclose(np.asarray(ts).reshape(7, -1)[:, 0], ts_o, rtol=0, atol=1e-11)
    theta0 = float(sim.get_output("theta")[0])
    assert abs(theta0 - 0.0236) < 5e-4
    sim.step(1.0); o.step(100)
    assert np.max(np.abs(sim.x[:, 0] - o.x) / np.maximum(np.abs(o.x), 1e-2)) < 1e-10      # still at trim
    e = fb.U["FB_U_ELEVATOR"]
    sim.set_input("elevator", sim.get_input("elevator") + 0.1)
    o.set_input(e, o.get_input(e) + 0.1)
    th_o = []
    for k in range(90):
        o.step(10); th_o.append(o.y(fb.Y["FB_Y_THETA"]))
    sim.run()
    assert sim.n_steps == 1000 and sim.status()[0] == fb.FB_ST_DONE          # run!(sim) ran to t_end (sim.jl:611)
    fl = state_floors(sim.nx, 0, 2)
    err = scaled_err(sim.x[:, 0], o.x, fl)
    assert err.max() < 1e-9, (int(err.argmax()), float(err.max()))
    th = fb.TimeSeries(sim).theta[:, 0]
    assert th.shape == (101,)
    assert np.max(np.abs(th[11:] - np.array(th_o))) < 1e-10          # samples at t = 1.1 .. 10.0
    assert abs(th[:11] - theta0).max() < 1e-5 and th.max() > 0.18     # flat for the first second (fuel burn moves the trim by 1e-6), then the pitch-up of the README plot
    print(f"cfg1: worst scaled state error after 10 s {err.max():.3e}, theta peak {th.max():.4f} rad")
    sim.close()


def test_empty_and_tiny_ensembles(fb):
    """Edge sizes: an ensemble of zero members is refused at construction (FB_ERR_ARG), ensembles of 1, 31, 33 members (one
    partial warp, one full + one partial) step and keep every member (ragged N: the padded lanes of the last warp stay out)."""
    with pytest.raises(fb.FlightBatchError, match="n_members"):
        fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=0, dt=0.01, t_end=1.0)
    ref = None
    for n in (1, 31, 33):
        sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=0.01, t_end=1.0)
        ok, _, _ = sim.init(fb.C172.TrimParameters(EAS=np.full(n, 47.0), h_e=np.full(n, 1200.0)))
        assert ok.all()
        sim.set_input("elevator", sim.get_input("elevator") + 0.05)
        sim.step(0.5)
        x = sim.x
        assert x.shape == (sim.nx, n) and (sim.status() == 0).all() and np.isfinite(x).all()
        if ref is None:
            ref = x[:, 0].copy()
        assert (x == ref[:, None]).all()          # identical members, identical trajectories whatever the ensemble size
        m = sim.stats("theta")
        assert m["count"] == n
        sim.close()
