"""GPU property tests at BASELINE.json's full cfg-2 size (65,536 members), where the oracle is too slow to follow:
size-independent properties of the integrator."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(fb, n, i0, seconds, mixed=False):
    import bench
    p = bench.member_params(i0, n)
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=0.01, t_end=1e4, steps_per_launch=100)
    if mixed:   # put every 16th member low over raised terrain so that it lands / crashes during the run
        elev = np.zeros(n); low = np.arange(n) % 16 == 0
        elev[low] = p["h_e"][low] - bench.geoid_height(p["lat"][low], p["lon"][low]) - np.where(np.arange(n)[low] % 32 == 0, 3.0, 5.0)
        sim.set_input("trn_elevation", elev)
    ok, _, _ = sim.init(fb.C172.TrimParameters(**p))
    ele_f, ail_f = fb.U["FB_U_ELEVATOR"], fb.U["FB_U_AILERON"]
    sim.schedule_inputs(*bench.doublet_events(n, sim.get_input(ele_f), sim.get_input(ail_f), ele_f, ail_f))
    sim.step(seconds)
    out = dict(x=sim.x, status=sim.status(), disc=sim.discrete(), wow=np.maximum.reduce([sim.get_output(k) for k in ("wow_left", "wow_right", "wow_nose")]))
    sim.close()
    return out


def test_sharding_and_launch_shape_do_not_change_trajectories(fb):
    """65,536 members in one handle (time-sliced airborne launches: 14 member-warps per SM on 12 warps) against the same
    members in two handles of 32,768 (plain 384-thread shape): the per-member result must not depend on the ensemble it is computed in. Different template
    instantiations may round differently (FMA contraction), so the bound is the parity tolerance, not bit equality."""
    n = 65536
    a = _run(fb, n, 0, 10.0)
    b0, b1 = _run(fb, n // 2, 0, 10.0), _run(fb, n // 2, n // 2, 10.0)
    xb = np.concatenate([b0["x"], b1["x"]], axis=1)
    assert np.array_equal(a["status"], np.concatenate([b0["status"], b1["status"]]))
    run = a["status"] == 0
    err = np.abs(a["x"] - xb)[:, run] / np.maximum(np.abs(xb[:, run]), 1e-2)
    print("sharding: max scaled difference", err.max(), "bit-identical members", int((err.max(axis=0) == 0).sum()), "of", int(run.sum()))
    assert err.max() < 1e-9
    # quaternion blocks stay normalised to the reference's renormalisation threshold (kinematics.jl:114-118)
    q_wb, q_ew = a["x"][12:16, run], a["x"][16:20, run]
    assert np.abs(np.sqrt((q_wb ** 2).sum(axis=0)) - 1).max() < 2e-8 and np.abs(np.sqrt((q_ew ** 2).sum(axis=0)) - 1).max() < 2e-8


def test_handover_path_equals_complete_kernel(fb):
    """A mixed ensemble (1/16 of the members 3 or 5 m above raised terrain: inside or just outside the 4 m height gate, they
    touch down or crash within the run)
    through the airborne kernel + hand-over, against the same ensemble through the complete kernel alone
    (FLIGHTBATCH_KERNEL=complete in a child process): statuses, stall / engine states identical, states to parity tolerance."""
    n = 8192
    a = _run(fb, n, 0, 12.0, mixed=True)
    code = ("import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r); import flightbatch as fb; "
            "from test_gpu_properties import _run; r = _run(fb, %d, 0, 12.0, mixed=True); "
            "np.savez('/tmp/fb_complete.npz', x=r['x'], status=r['status'], stall=r['disc'][0], eng=r['disc'][1])") % (ROOT, os.path.join(ROOT, "tests"), n)
    subprocess.check_call([sys.executable, "-c", code], env=dict(os.environ, FLIGHTBATCH_KERNEL="complete"))
    b = np.load("/tmp/fb_complete.npz")
    low = np.arange(n) % 16 == 0
    touched = (a["wow"][low] > 0).sum() + np.isin(a["status"][low], (2, 3)).sum()
    assert touched > 20, touched                      # the hand-over path really ran: members on their wheels or crashed
    assert np.array_equal(a["status"], b["status"])
    assert np.array_equal(a["disc"][0], b["stall"]) and np.array_equal(a["disc"][1], b["eng"])
    run = a["status"] == 0
    err = np.abs(a["x"] - b["x"]) / np.maximum(np.abs(b["x"]), 1.0)
    air = run & ~low
    print("hand-over: max diff airborne", err[:, air].max(), "near-ground", err[:, run & low].max() if (run & low).any() else None,
          "statuses", np.bincount(a["status"], minlength=7).tolist())
    assert err[:, air].max() < 1e-9
    assert err[:, run & low].max() < 1e-5      # ground-contact conditioning


def test_pipelined_handover_pass_equals_serial(fb):
    """The hand-over pass of launch k runs on a stream of its own beside the airborne launch k + 1 (ownership words keep the two
    kernels on disjoint members; the host joins every eighth launch and before every read). The same mixed ensemble — 1/16 of
    the members landing or crashing — stepped through 12 back-to-back launches with the pass pipelined (default) and with it
    on the launch stream behind each airborne launch (FLIGHTBATCH_PIPELINE=0, child process): every member runs through the
    same kernels over the same steps either way, so states, statuses and discrete states are bit-identical. Reading the state
    between launches (a join after every launch: members that have climbed away return to the airborne kernel earlier) moves
    nothing beyond the rounding the two kernel instantiations differ by."""
    n = 8192
    a = _run(fb, n, 0, 12.0, mixed=True)
    code = ("import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r); import flightbatch as fb; "
            "from test_gpu_properties import _run; r = _run(fb, %d, 0, 12.0, mixed=True); "
            "np.savez('/tmp/fb_serial.npz', x=r['x'], status=r['status'], stall=r['disc'][0], eng=r['disc'][1])") % (ROOT, os.path.join(ROOT, "tests"), n)
    subprocess.check_call([sys.executable, "-c", code], env=dict(os.environ, FLIGHTBATCH_PIPELINE="0"), timeout=300)
    b = np.load("/tmp/fb_serial.npz")
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["disc"][0], b["stall"]) and np.array_equal(a["disc"][1], b["eng"])
    assert np.array_equal(a["x"], b["x"])
    low = np.arange(n) % 16 == 0
    assert (a["wow"][low] > 0).sum() + np.isin(a["status"][low], (2, 3)).sum() > 20
    # the same run with a read (= a join) after every launch
    import bench
    p = bench.member_params(0, n)
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=0.01, t_end=1e4, steps_per_launch=100)
    elev = np.zeros(n)
    elev[low] = p["h_e"][low] - bench.geoid_height(p["lat"][low], p["lon"][low]) - np.where(np.arange(n)[low] % 32 == 0, 3.0, 5.0)
    sim.set_input("trn_elevation", elev)
    sim.init(fb.C172.TrimParameters(**p))
    ele_f, ail_f = fb.U["FB_U_ELEVATOR"], fb.U["FB_U_AILERON"]
    sim.schedule_inputs(*bench.doublet_events(n, sim.get_input(ele_f), sim.get_input(ail_f), ele_f, ail_f))
    for _ in range(12):
        sim.step(1.0)
        st = sim.status()
    x = sim.x
    assert np.array_equal(st, a["status"])
    run = st == 0
    err = np.abs(x - a["x"]) / np.maximum(np.abs(a["x"]), 1.0)
    print("pipelined vs read-every-launch: max diff airborne", err[:, run & ~low].max(), "near-ground", err[:, run & low].max())
    assert err[:, run & ~low].max() < 1e-9 and err[:, run & low].max() < 1e-5
    sim.close()


def test_pipelined_state_read_matches_blocking_read(fb):
    """fb_get_state_async snapshots the state as of the work submitted so far: a read issued after step k and collected
    after step k+1 has been launched must equal the blocking read taken after step k (bit for bit), not the later state."""
    import torch
    import bench
    n = 4096
    p = bench.member_params(0, n)

    def make():
        sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=0.01, t_end=1e4, steps_per_launch=50)
        sim.init(fb.C172.TrimParameters(**p))
        sim.set_input("elevator", sim.get_input("elevator") + 0.05)
        return sim
    a, b = make(), make()
    bufs = [torch.empty((a.nx, n), dtype=torch.float64).pin_memory().numpy() for _ in range(2)]
    ref = []
    for k in range(4):
        a.step(0.5); ref.append(a.x.copy())
    got = []
    for k in range(4):
        b.step(0.5)
        b.get_state_async(bufs[k % 2])      # waits for read k-1 first, so bufs[(k-1) % 2] is complete here
        if k > 0:
            got.append(bufs[(k - 1) % 2].copy())
    b.wait_state(); got.append(bufs[3 % 2].copy())
    for k in range(4):
        assert np.array_equal(ref[k], got[k]), k
    a.close(); b.close()


@pytest.mark.parametrize("model", ["Cessna172Sv0", "Cessna172Xv2"])
def test_time_sliced_launch_is_bit_identical_to_plain_launch(fb, model):
    """A ragged ensemble of 60,001 members (13 member-warps per SM) advanced by the time-sliced launch shape (chunks of
    steps, member-warps rotating over the 12 hardware warps) and by the plain one-thread-per-member launches
    (FLIGHTBATCH_SLICED=0): same states, discrete states, statuses and log, bit for bit — scheduled inputs fire inside
    chunks, the autopilot runs every step, the log is decimated across chunk boundaries."""
    import bench
    n = 60001
    p = bench.member_params(0, n)

    def run(sliced):
        os.environ["FLIGHTBATCH_SLICED"] = "1" if sliced else "0"
        try:
            ac = getattr(fb, model)(fb.WA())
            sim = fb.BatchSimulation(fb.SimpleWorld(ac), n=n, dt=0.01, Dt=0.01, t_end=4.0, steps_per_launch=100, saveat=0.07,
                                     log_fields=("theta", "eas"))
        finally:
            del os.environ["FLIGHTBATCH_SLICED"]
        if model == "Cessna172Xv2":
            rng = np.random.default_rng(5)
            sim.set_input("wind_n", rng.uniform(-10, 10, n))
        sim.init(fb.C172.TrimParameters(**p))
        idx = np.arange(n)
        if model == "Cessna172Xv2":
            sim.set_input("lon_mode_req", 7.0); sim.set_input("lat_mode_req", 3.0)
            sim.set_input("lon_clm_ref", np.linspace(0.5, 3.0, n)); sim.set_input("lat_phi_ref", np.linspace(-0.5, 0.5, n))
            sim.schedule_inputs(idx, np.full(n, 123), np.full(n, fb.U["FB_U_LAT_PHI_REF"]), np.linspace(0.5, -0.5, n))
        else:
            ele = sim.get_input("elevator")
            sim.schedule_inputs(np.concatenate([idx, idx]), np.concatenate([np.full(n, 57), np.full(n, 130)]),
                                np.full(2 * n, fb.U["FB_U_ELEVATOR"]), np.concatenate([ele + 0.05, ele - 0.05]))
        sim.step(3.0)
        sim.step(0.01); sim.step(0.07)      # launches too short to be cut up: 13 member-warps on 12 warps in one chunk
        ts = fb.TimeSeries(sim)
        out = dict(x=sim.x, status=sim.status(), disc=sim.discrete(), counters=sim.counters(), log=[ts.theta.copy(), ts.eas.copy()])
        sim.close()
        return out
    a, b = run(True), run(False)
    assert a["counters"]["rhs_evals"] > b["counters"]["rhs_evals"]     # the sliced shape really ran (one extra RHS per chunk start)
    assert np.array_equal(a["status"], b["status"])
    assert np.array_equal(a["disc"][0], b["disc"][0]) and np.array_equal(a["disc"][1], b["disc"][1])
    assert np.array_equal(a["x"], b["x"])
    for la, lb in zip(a["log"], b["log"]):
        assert np.array_equal(la, lb, equal_nan=True)     # a frozen member leaves NaN in both


@pytest.mark.parametrize("model", ["Cessna172Sv0", "Cessna172Xv2"])
def test_dynamic_task_queue_is_bit_identical_to_plain_launch(fb, model):
    """The airborne launches that draw (member-warp, chunk) tasks from a global queue (one persistent CTA per SM, a member's
    state handed from SM to SM through L2 between chunks: FLIGHTBATCH_DYNAMIC=1) against the plain one-thread-per-member
    launches: a ragged 60,001-member ensemble, scheduled inputs inside chunks, autopilot every step, a decimated log, some members
    flown into the terrain gate so that the hand-over to the complete kernel happens in the middle of a chunk. Bit for bit."""
    import bench
    n = 60001
    p = bench.member_params(0, n)

    def run(dyn):
        os.environ["FLIGHTBATCH_DYNAMIC"] = "1" if dyn else "0"
        if not dyn:
            os.environ["FLIGHTBATCH_SLICED"] = "0"
        try:
            ac = getattr(fb, model)(fb.WA())
            sim = fb.BatchSimulation(fb.SimpleWorld(ac), n=n, dt=0.01, Dt=0.01, t_end=4.0, steps_per_launch=100, saveat=0.07,
                                     log_fields=("theta", "eas"))
        finally:
            del os.environ["FLIGHTBATCH_DYNAMIC"]
            os.environ.pop("FLIGHTBATCH_SLICED", None)
        rng = np.random.default_rng(5)
        if model == "Cessna172Xv2":
            sim.set_input("wind_n", rng.uniform(-10, 10, n))
        # every 97th member starts 4.5..6.5 m above its terrain and descends: it crosses the 4 m gate inside a chunk
        low = np.arange(0, n, 97)
        elev = np.zeros(n); elev[low] = p["h_e"][low] - bench.geoid_height(p["lat"][low], p["lon"][low]) - rng.uniform(4.5, 6.5, low.size)
        sim.set_input("trn_elevation", elev)
        sim.init(fb.C172.TrimParameters(**p))
        idx = np.arange(n)
        if model == "Cessna172Xv2":
            sim.set_input("lon_mode_req", 7.0); sim.set_input("lat_mode_req", 3.0)
            clm = np.linspace(0.5, 3.0, n); clm[low] = -3.0
            sim.set_input("lon_clm_ref", clm); sim.set_input("lat_phi_ref", np.linspace(-0.5, 0.5, n))
            sim.schedule_inputs(idx, np.full(n, 123), np.full(n, fb.U["FB_U_LAT_PHI_REF"]), np.linspace(0.5, -0.5, n))
        else:
            ele = sim.get_input("elevator")
            d = np.full(n, 0.05); d[low] = -0.08
            sim.schedule_inputs(np.concatenate([idx, idx]), np.concatenate([np.full(n, 57), np.full(n, 130)]),
                                np.full(2 * n, fb.U["FB_U_ELEVATOR"]), np.concatenate([ele + d, ele - 0.05]))
        sim.step(3.0)
        sim.step(0.01); sim.step(0.07)
        ts = fb.TimeSeries(sim)
        out = dict(x=sim.x, status=sim.status(), disc=sim.discrete(), counters=sim.counters(), log=[ts.theta.copy(), ts.eas.copy()],
                   low_h=sim.get_output("h_o")[low] - elev[low])
        sim.close()
        return out
    a, b = run(True), run(False)
    assert a["counters"]["rhs_evals"] > b["counters"]["rhs_evals"]     # the queue really cut the launches into chunks
    assert (b["low_h"] < 4.0).sum() > 50                               # members did go through the hand-over
    assert np.array_equal(a["status"], b["status"])
    assert np.array_equal(a["disc"][0], b["disc"][0]) and np.array_equal(a["disc"][1], b["disc"][1])
    assert np.array_equal(a["x"], b["x"])
    for la, lb in zip(a["log"], b["log"]):
        assert np.array_equal(la, lb, equal_nan=True)
