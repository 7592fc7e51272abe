"""GPU: the decimated device log behind TimeSeries(sim) (lib/FlightCore/src/sim.jl:345-347,644-704) against the oracle's
saved outputs, including the reference-style lazy path drilling of the host mirror."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SEED = 20261018


def test_log_matches_oracle_and_paths_drill(fb, orc):
    n, dt, every = 16, 0.01, 10
    rng = np.random.default_rng(SEED + 60)
    EAS, h = rng.uniform(35, 55, n), rng.uniform(500, 2500, n)
    fields = ("theta", "h_e", "eas", "alpha", "q", "v_d", "eng_omega", "stall")
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=dt, t_end=6.0, saveat=every * dt, log_fields=fields,
                             steps_per_launch=64)
    ok, _, _ = sim.init(fb.C172.TrimParameters(EAS=EAS, h_e=h))
    assert ok.all()
    ele = sim.get_input("elevator")
    ev = [(i, 100, fb.U["FB_U_ELEVATOR"], ele[i] + 0.1) for i in range(n)]      # README example: elevator step at t = 1 s
    sim.schedule_inputs(*zip(*ev))
    sim.run()
    ts = fb.TimeSeries(sim)
    assert len(ts) == 61 and np.allclose(ts.t, np.arange(61) * every * dt)
    # reference-style paths resolve to the same columns
    assert ts.aircraft.vehicle.kinematics.e_nb.θ is ts.theta
    assert ts.aircraft.vehicle.kinematics.v_eb_n[2] is ts.v_d
    assert ts.aircraft.vehicle.kinematics.ω_wb_b[1] is ts.q
    assert ts.aircraft.vehicle.systems.pwp.engine.ω is ts["eng_omega"]
    assert "kinematics" in dir(ts.aircraft.vehicle)
    with pytest.raises(AttributeError):
        ts.aircraft.vehicle.nothing
    with pytest.raises(fb.FlightBatchError):
        ts.aircraft.vehicle.kinematics.e_nb.ψ          # a valid path that was not logged
    sub = ts[10:21]
    assert len(sub) == 11 and np.array_equal(sub.theta, ts.theta[10:21]) and sub.t[0] == ts.t[10]
    # every saved sample against the oracle's y at the same instants
    for i in range(0, n, 3):
        o = orc.Sim(variant=0, kin=0, dt=dt)
        ok_o, _, _ = o.init_trim(EAS=float(EAS[i]), h_e=float(h[i]))
        assert ok_o
        o.add_event(100, fb.U["FB_U_ELEVATOR"], float(ele[i] + 0.1))
        for k in range(61):
            if k:
                assert o.step(every) == 0
            for f in fields:
                a, b = getattr(ts, f)[k, i], o.y(fb.OUTPUT_NAMES[f])
                assert abs(a - b) <= 1e-9 * max(abs(b), 1e-2), (f, k, i, a, b)
    assert ts.theta[60].mean() > ts.theta[0].mean() + 0.05      # the step pitched the ensemble up
    sim.close()
