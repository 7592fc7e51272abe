"""GPU parity on the ground-contact path (BASELINE.json configs[3] at test size): takeoff rolls from a drop onto the
gear with an engine start, and 3-degree approaches to touchdown and roll-out; landing gear struts, contact friction
regulators, braking / steering / surface types, PistonEngine state machine (off -> starting -> running), crash checks."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SEED = 20261018


def build_cfg4(fb, orc, n, idx, dt=0.01):
    """n members, first half takeoff (dropped from 1.9 + U[0, 1] m above the terrain, engine start, throttle ramp), second
    half on approach (trimmed at EAS 30-35, flaps 1, gamma = -3 deg, 3-12 m above the terrain). Returns (sim, oracles, idx)."""
    import bench
    rng = np.random.default_rng(SEED + 40)
    half = n // 2
    lat = np.deg2rad(rng.uniform(-60, 60, n)); lon = rng.uniform(-np.pi, np.pi, n)
    geoid = bench.geoid_height(lat, lon)
    elev = rng.uniform(0, 1500, n)
    surface = rng.integers(0, 3, n).astype(float)
    psi = rng.uniform(-np.pi, np.pi, n)
    U = fb.U
    common = [("trn_elevation", elev), ("trn_surface", surface)]
    # ---- takeoff members: C172.Init on the ground
    simA = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=dt, t_end=1000.0, steps_per_launch=50)
    for k, v in common:
        simA.set_input(k, v)
    # one handle holds both groups: initialise everyone on the ground, then overwrite the second half with approach trims
    h_gnd = elev + geoid + 1.9 + rng.uniform(0, 1, n)
    q0, q3 = np.cos(psi / 2), np.sin(psi / 2)
    brake = rng.uniform(0, 0.3, n) * (rng.uniform(0, 1, n) < 0.3)
    rudder = rng.uniform(-0.2, 0.2, n)
    simA.init(fb.C172.Init(q0=q0, q3=q3, lat=lat, lon=lon, h_e=h_gnd, mixture=0.5, brake_left=brake, brake_right=brake, rudder=rudder,
                           fuel_load=rng.uniform(0.2, 1, n)))
    xA = simA.x
    stA, enA = simA.discrete()
    inputsA = {f: simA.get_input(f) for f in ("throttle", "elevator", "aileron", "rudder", "flaps", "brake_left", "brake_right", "mixture", "m_pilot", "m_copilot", "m_lpass", "m_rpass", "m_baggage")}
    # ---- approach members: trim on a -3 deg path
    EAS = rng.uniform(30, 35, n)
    h_app = elev + geoid + rng.uniform(3, 12, n)
    ok, ts, cost = simA.init(fb.C172.TrimParameters(EAS=EAS, h_e=h_app, flaps=1.0, gamma_wb_n=np.deg2rad(-3.0), psi_nb=psi, lat=lat, lon=lon))
    assert ok[half:].all()
    xB = simA.x
    stB, enB = simA.discrete()
    inputsB = {f: simA.get_input(f) for f in inputsA}
    x = np.concatenate([xA[:, :half], xB[:, half:]], axis=1)
    st = np.concatenate([stA[:half], stB[half:]]); en = np.concatenate([enA[:half], enB[half:]])
    for f in inputsA:
        simA.set_input(f, np.concatenate([inputsA[f][:half], inputsB[f][half:]]))
    simA.init_state(x, stall=st, eng_state=en)
    # schedules: takeoff = starter on, throttle ramp, rotate; approach = flare + idle + brakes
    ev = []
    for i in range(half):
        ev += [(i, 0, U["FB_U_ENG_START"], 1.0), (i, 300, U["FB_U_ENG_START"], 0.0), (i, 300, U["FB_U_THROTTLE"], 0.3), (i, 500, U["FB_U_THROTTLE"], 1.0),
               (i, 500, U["FB_U_BRAKE_LEFT"], 0.0), (i, 500, U["FB_U_BRAKE_RIGHT"], 0.0), (i, 1500, U["FB_U_ELEVATOR"], 0.3)]
    for i in range(half, n):
        ev += [(i, 150, U["FB_U_THROTTLE"], 0.0), (i, 150, U["FB_U_ELEVATOR"], float(inputsB["elevator"][i]) + 0.15),
               (i, 900, U["FB_U_BRAKE_LEFT"], 0.5), (i, 900, U["FB_U_BRAKE_RIGHT"], 0.5)]
    m, s, f, v = zip(*ev)
    simA.schedule_inputs(m, s, f, v)
    oracles = []
    all_inputs = {f: simA.get_input(f) for f in ("aileron", "elevator", "rudder", "flaps", "brake_left", "brake_right", "throttle", "mixture", "eng_start",
                                                 "eng_stop", "mixture_auto", "steering_engaged", "m_pilot", "m_copilot", "m_lpass", "m_rpass", "m_baggage",
                                                 "trn_elevation", "trn_surface")}
    for i in idx:
        o = orc.Sim(variant=0, kin=0, dt=dt)
        for f, val in all_inputs.items():
            o.set_input(fb.INPUT_NAMES[f], val[i])
        for (mi, si, fi, vi) in ev:
            if mi == i:
                o.add_event(si, fi, vi)
        o.init_state(x[:, i], stall=int(st[i]), eng_state=int(en[i]))
        oracles.append(o)
    return simA, oracles, idx


SPREAD_FACTOR = 5.0     # GPU-vs-oracle error allowed as a multiple of the oracle-vs-oracle rounding spread measured in the same test


def test_cfg4_takeoff_and_landing_ensemble(fb, orc):
    """Discrete results (status, weight-on-wheels per strut, engine state, stall) are bit-exact at every sampled instant
    over 25 s. Continuous states cannot agree to 1e-9 on this path for ANY two implementations that round differently: the
    reference computes the strut compression from a difference of 6.4e6 m ECEF vectors (landinggear.jl:262-270) and a member
    at rest sits on the friction-direction switch at 1e-3 m/s (landinggear.jl:436-452). The bound is therefore MEASURED, not
    picked: the same members are run on the oracle under three rounding perturbations (FMA contraction on / off, a one-ulp
    change of the start, one-ulp noise after every step; tests/ground_util.py), and at every instant the GPU-vs-oracle error (median over members of the all-state error, maximum
    over members of the physical-state error) must stay within SPREAD_FACTOR times the largest oracle-vs-oracle spread seen
    up to that instant. While every strut is still unloaded (first instant) the airborne bound 1e-9 applies."""
    import ground_util as gu
    n = 64
    idx = list(range(0, n, 2))
    sim, oracles, idx = build_cfg4(fb, orc, n, idx)
    # the oracle-vs-oracle spread from the very same initial states, inputs and schedules
    x0 = np.array([o.x for o in oracles]).T
    disc = np.array([o.discrete for o in oracles])
    inputs = {f: sim.get_input(f)[idx] for f in gu.INPUT_FIELDS}
    events = [gu.cfg4_events(fb.U, i, n // 2, inputs["elevator"][k]) for k, i in enumerate(idx)]
    spread = gu.oracle_spread(orc, fb.U, fb.INPUT_NAMES, x0, disc[:, 0], disc[:, 1], inputs, events)
    seen_wow, seen_eng = set(), set(int(v) for v in sim.discrete()[1])
    worst_med, worst_max, run_med, run_max, worst_ratio = 0.0, 0.0, 0.0, 0.0, 0.0
    for chunk, sp in zip(gu.CHUNKS, spread):
        sim.step(chunk * 0.01)
        for o in oracles:
            o.step(chunk)
        xg = sim.x[:, idx]
        xo = np.array([o.x for o in oracles]).T
        st = sim.status()[idx]; so = np.array([o.status for o in oracles])
        assert np.array_equal(st, so)
        stall_g, eng_g = sim.discrete()
        disc_o = np.array([o.discrete for o in oracles])
        assert np.array_equal(stall_g[idx], disc_o[:, 0]) and np.array_equal(eng_g[idx], disc_o[:, 1])
        for k in ("wow_left", "wow_right", "wow_nose"):
            wg = sim.get_output(k)[idx]; wo = np.array([o.y(fb.OUTPUT_NAMES[k]) for o in oracles])
            assert np.array_equal(wg, wo), (k, np.nonzero(wg != wo))
            seen_wow |= set(int(v) for v in wg)
        seen_eng |= set(int(v) for v in eng_g)
        err, phys = gu.member_errors(xg, xo)
        run_med = max(run_med, *(float(np.median(sp[e][0])) for e in ("fma", "ulp", "noisy")))
        run_max = max(run_max, *(float(sp[e][1].max()) for e in ("fma", "ulp", "noisy")))
        med, mx = float(np.median(err)), float(phys.max())
        worst_med = max(worst_med, med); worst_max = max(worst_max, mx)
        worst_ratio = max(worst_ratio, med / max(run_med, 1e-9), mx / max(run_max, 1e-9))
        assert med <= max(SPREAD_FACTOR * run_med, 1e-9), (med, run_med)
        assert mx <= max(SPREAD_FACTOR * run_max, 1e-9), (mx, run_max, int(phys.argmax()))
    assert seen_wow == {0, 1} and seen_eng == {0, 1, 2}      # contact made and lost, engine off -> starting -> running
    # the takeoff half flew away, the approach half is rolling or stopped on the runway
    wow_any = np.maximum.reduce([sim.get_output(k) for k in ("wow_left", "wow_right", "wow_nose")])
    assert wow_any[: n // 2].mean() < 0.5 and wow_any[n // 2:].mean() > 0.9
    print(f"cfg4 GPU-vs-oracle median / max member error {worst_med:.3e} / {worst_max:.3e}; oracle-vs-oracle spread {run_med:.3e} / {run_max:.3e}; "
          f"worst ratio {worst_ratio:.2f} (allowed {SPREAD_FACTOR})")
    sim.close()


def test_crash_statuses_match_the_oracle(fb, orc):
    """GroundCrash terminations (landinggear.jl:331-347): a hard drop exceeds 10 m/s of compression rate; a member rolled
    on its side exceeds the 60 deg strut-terrain angle. The member freezes with the same status at the same step."""
    import bench
    n = 6
    g = float(bench.geoid_height(0.0, 0.0))
    h = g + np.array([1.9, 3.0, 9.0, 14.0, 2.0, 2.0])
    vD = np.array([0.0, 2.0, 6.0, 9.0, 0.0, 0.0])
    phi = np.array([0.0, 0.0, 0.0, 0.0, 1.3, -1.4])
    q0, q1 = np.cos(phi / 2), np.sin(phi / 2)
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=0.01, t_end=100.0, steps_per_launch=25)
    sim.init(fb.C172.Init(q0=q0, q1=q1, h_e=h, vD=vD))
    oracles = []
    for i in range(n):
        o = orc.Sim(variant=0, kin=0, dt=0.01)
        o.init(q0=q0[i], q1=q1[i], h_e=h[i], vD=vD[i])
        oracles.append(o)
    for _ in range(8):
        sim.step(0.25)
        for o in oracles:
            o.step(25)
        st = sim.status(); so = np.array([o.status for o in oracles])
        assert np.array_equal(st, so), (st, so)
        xg = sim.x
        for i, o in enumerate(oracles):
            if so[i] != 0:   # frozen members stay where they stopped
                assert np.abs(xg[:, i] - o.x).max() < 1e-5
    assert set(st) >= {0, fb.FB_ST_CRASH_RATE, fb.FB_ST_CRASH_ANGLE}, st
    sim.close()
