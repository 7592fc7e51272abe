"""Full-horizon parity of the two bench workloads: 64 sampled members of the cfg-2 ensemble over its whole
60 s (6,000 RK4 steps: doublets at 5-9 s and 20-24 s) and of the cfg-3 ensemble over its whole 600 s (60,000 steps, control
laws every step: BASELINE.json configs[1], configs[2]) against the CPU oracle stepped on all host cores. Continuous state
<= 1e-9 scaled error, control modes / altitude-tracking state / stall / engine state bit-exact, at three instants each."""
import os

import numpy as np
import pytest

from parity_util import scaled_err, state_floors

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1


def _compare(fb, sim, oracles, idx, kin, tol, modes):
    xg = sim.x
    stg, eng = sim.discrete()
    fl = state_floors(sim.nx, 0 if sim.nx < 30 else 1, kin)
    outs = {k: sim.get_output(k) for k in modes}
    status = sim.status()
    worst = 0.0
    for o, i in zip(oracles, idx):
        assert status[i] == o.status, (i, status[i], o.status)
        err = scaled_err(xg[:, i], o.x, fl)
        worst = max(worst, float(err.max()))
        assert err.max() < tol, (i, int(err.argmax()), float(err.max()))
        assert o.discrete == (int(stg[i]), int(eng[i])), i
        for k in modes:
            assert outs[k][i] == o.y(fb.OUTPUT_NAMES[k]), (k, i, outs[k][i], o.y(fb.OUTPUT_NAMES[k]))
    return worst


def test_cfg2_full_60_seconds(fb, orc):
    import bench
    n, dt = 2048, 0.01
    idx = list(range(5, n, 32))
    p = bench.member_params(0, n)
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=dt, t_end=100.0, steps_per_launch=100)
    ok, _, _ = sim.init(fb.C172.TrimParameters(**p))
    ele_f, ail_f = fb.U["FB_U_ELEVATOR"], fb.U["FB_U_AILERON"]
    ev = bench.doublet_events(n, sim.get_input(ele_f), sim.get_input(ail_f), ele_f, ail_f)
    sim.schedule_inputs(*ev)
    idx = [i for i in idx if ok[i]]
    oracles = []
    for i in idx:
        o = orc.Sim(variant=0, kin=0, dt=dt)
        ok_o, _, _ = o.init_trim(**{k: float(v[i]) for k, v in p.items()})
        assert ok_o
        for m, s, f, v in zip(*ev):
            if m == i:
                o.add_event(s, f, v)
        oracles.append(o)
    worst = 0.0
    for t_adv in (10.0, 20.0, 30.0):
        sim.step(t_adv)
        orc.batch_step(oracles, int(round(t_adv / dt)), THREADS)
        worst = max(worst, _compare(fb, sim, oracles, idx, 0, 1e-9, ()))
    assert sim.n_steps == 6000
    print(f"cfg2, {len(idx)} members x 6000 steps: worst scaled state error {worst:.3e}")
    sim.close()


def test_cfg3_full_600_seconds(fb, orc):
    import bench
    n, dt = 2048, 0.01
    idx = list(range(7, n, 32))
    p = bench.cfg3_member_params(0, n)
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Xv2(fb.WA())), n=n, dt=dt, Dt=dt, t_end=700.0, steps_per_launch=100)
    sim.set_input("wind_n", p["wind_n"]); sim.set_input("wind_e", p["wind_e"])
    ok, _, _ = sim.init(fb.C172.TrimParameters(**{k: p[k] for k in bench.TRIM_KEYS}))
    sim.set_input("lon_mode_req", bench.LON_EAS_CLM); sim.set_input("lat_mode_req", bench.LAT_PHI_BETA)
    sim.set_input("lon_clm_ref", p["clm_ref"]); sim.set_input("lat_phi_ref", p["phi_ref"])
    idx = [i for i in idx if ok[i]]
    U = fb.U
    oracles = []
    for i in idx:
        o = orc.Sim(variant=3, kin=0, dt=dt, Dt=dt)
        o.set_input(U["FB_U_WIND_N"], p["wind_n"][i]); o.set_input(U["FB_U_WIND_E"], p["wind_e"][i])
        ok_o, _, _ = o.init_trim(**{k: float(p[k][i]) for k in bench.TRIM_KEYS})
        assert ok_o
        o.set_input(U["FB_U_LON_MODE_REQ"], bench.LON_EAS_CLM); o.set_input(U["FB_U_LAT_MODE_REQ"], bench.LAT_PHI_BETA)
        o.set_input(U["FB_U_LON_CLM_REF"], p["clm_ref"][i]); o.set_input(U["FB_U_LAT_PHI_REF"], p["phi_ref"][i])
        oracles.append(o)
    modes = ("lon_mode", "lat_mode", "alt_state")
    worst = 0.0
    for t_adv in (60.0, 240.0, 300.0):
        sim.step(t_adv)
        orc.batch_step(oracles, int(round(t_adv / dt)), THREADS)
        worst = max(worst, _compare(fb, sim, oracles, idx, 0, 1e-9, modes))
    assert sim.n_steps == 60000
    lon = sim.get_output("lon_mode")[idx]
    assert (lon == bench.LON_EAS_CLM).all()
    print(f"cfg3, {len(idx)} members x 60000 steps: worst scaled state error {worst:.3e}")
    sim.close()
