"""Ensemble statistics on the device (north star: "warp-level shuffles handle ensemble statistics", "NCCL allreduce used only
for ensemble statistics"; no reference counterpart, lib/FlightCore/src/sim.jl:173-182: one Simulation = one aircraft).
fb_stats (on-demand moments of one output) and the moments the integrator accumulates in its log tap are checked against
numpy over the per-member values; the in-library NCCL all-reduce against the same moments on one rank and, with two GPUs,
over a sharded group."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 20261018


def _host(values, keep):
    v = np.asarray(values, dtype=np.float64)[keep]
    return dict(count=int(v.size), sum=float(v.sum()), sum_sq=float((v * v).sum()), min=float(v.min()), max=float(v.max()))


def _close(a, b, tol=1e-12):
    assert a["count"] == b["count"], (a, b)
    for k in ("sum", "sum_sq"):
        assert abs(a[k] - b[k]) <= tol * max(1.0, abs(b[k])) * max(1, b["count"]) ** 0.5, (k, a[k], b[k])
    assert a["min"] == b["min"] and a["max"] == b["max"], (a, b)


def _ensemble(fb, n, rng, **kw):
    EAS = rng.uniform(30, 55, n)
    EAS[::97] = 12.0            # far below stall speed: these trims fail and the members are excluded from every moment
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=0.01, t_end=20.0, **kw)
    ok, _, _ = sim.init(fb.C172.TrimParameters(EAS=EAS, h_e=rng.uniform(500, 3000, n), psi_nb=rng.uniform(-3, 3, n)))
    assert 0 < (~ok).sum() < n // 50
    return sim, ok


@pytest.mark.parametrize("n", [1000, 4096 + 17])
def test_fb_stats_matches_host_moments(fb, n):
    """fb_stats over an ensemble whose size is not a multiple of the warp, with excluded (trim-failed) members."""
    rng = np.random.default_rng(SEED + 60)
    sim, ok = _ensemble(fb, n, rng)
    sim.set_input("elevator", sim.get_input("elevator") + rng.uniform(-0.05, 0.05, n))
    sim.step(1.0)
    keep = sim.status() == 0
    assert (keep == ok).all()
    for f in ("h_e", "theta", "eas", "alpha"):
        _close(sim.stats(f), _host(sim.get_output(f), keep))
    sim.close()


def test_log_moments_match_the_member_samples(fb):
    """The moments the integrator folds into its log tap equal numpy's over the per-member samples, sample by sample; a member
    frozen before a sample contributes neither a value (NaN in the log) nor a count."""
    n = 3000
    rng = np.random.default_rng(SEED + 61)
    fields = ("h_e", "eas", "theta", "phi")
    sim, ok = _ensemble(fb, n, rng, saveat=0.25, log_fields=fields, steps_per_launch=60)
    sim.set_input("aileron", rng.uniform(-0.05, 0.05, n))
    sim.step(3.0)                       # launches of 60 steps: log samples fall inside and across launches
    ts = fb.TimeSeries(sim)
    assert ts.n_samples == 13
    st = sim.log_stats()
    for f in fields:
        col = ts[f]
        assert np.isnan(col[1:, ~ok]).all() and not np.isnan(col[:, ok]).any()     # frozen members never wrote a sample
        for k in range(ts.n_samples):
            got = {q: st[f][q][k] for q in ("count", "sum", "sum_sq", "min", "max")}
            _close(got, _host(col[k], ok))
    # the asynchronous form returns the same numbers for a sub-range
    sim.log_stats_begin(4, 3)
    sim.step(0.5)                       # more work goes in behind the fetch
    part = sim.log_stats_end()
    assert (part["eas"]["sum"] == st["eas"]["sum"][4:7]).all() and (part["eas"]["count"] == st["eas"]["count"][4:7]).all()
    sim.close()


def test_moments_only_log_and_reinit(fb):
    """log_members=False keeps the moments only (no [samples][fields][N] array); a second init! starts a clean record."""
    n = 2048
    rng = np.random.default_rng(SEED + 62)
    fields = ("h_e", "eas")
    a, ok = _ensemble(fb, n, rng, saveat=0.5, log_fields=fields)
    rng = np.random.default_rng(SEED + 62)
    b, _ = _ensemble(fb, n, rng, saveat=0.5, log_fields=fields, log_members=False)
    for s in (a, b):
        s.step(2.0)
    sa, sb = a.log_stats(), b.log_stats()
    for f in fields:
        assert (sa[f]["count"] == sb[f]["count"]).all() and (sa[f]["min"] == sb[f]["min"]).all() and (sa[f]["max"] == sb[f]["max"]).all()
        assert np.allclose(sa[f]["sum"], sb[f]["sum"], rtol=1e-13, atol=0)      # the atomics commute, the additions do not associate
    with pytest.raises(fb.FlightBatchError, match="not kept"):
        fb.TimeSeries(b).h_e
    b.init(fb.C172.TrimParameters(EAS=45.0, h_e=1000.0))                           # every member the same
    b.step(1.0)
    m = b.log_stats()
    assert (m["eas"]["count"] == n).all() and m["eas"]["count"].size == 3
    assert np.allclose(m["eas"]["min"], m["eas"]["max"], rtol=1e-12)
    a.close(); b.close()


def test_allreduce_on_one_rank_is_the_identity(fb):
    """The in-library NCCL path (communicator from a unique id, side stream, fetch ordered behind it) on a world of one."""
    n = 1500
    rng = np.random.default_rng(SEED + 63)
    sim, ok = _ensemble(fb, n, rng, saveat=0.5, log_fields=("theta", "h_e"))
    sim.comm_init(1, 0, fb.comm_unique_id())
    sim.step(1.0)
    before = sim.log_stats()
    sim.log_stats_allreduce()
    sim.step(1.0)
    sim.log_stats_allreduce()            # only the samples completed since the previous call
    after = sim.log_stats()
    for q in ("count", "sum", "sum_sq", "min", "max"):
        assert (after["theta"][q][:3] == before["theta"][q]).all()
    assert after["theta"]["count"].size == 5 and (after["theta"]["count"] == int(ok.sum())).all()
    assert sim.comm_time_ms() > 0.0
    sim.close()


def test_group_shards_and_reduces_across_two_gpus(fb):
    """fb_group_*: one process, two GPUs, global host arrays; states equal a single handle's, moments are the whole ensemble's."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n = 5000
    rng = np.random.default_rng(SEED + 64)
    p = dict(EAS=rng.uniform(35, 55, n), h_e=rng.uniform(500, 3000, n), psi_nb=rng.uniform(-3, 3, n))
    ele = rng.uniform(-0.05, 0.05, n)
    kw = dict(n=n, dt=0.01, t_end=20.0, saveat=0.5, log_fields=("h_e", "theta"))
    one = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), **kw)
    grp = fb.BatchGroup(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), devices=(0, 1), **kw)
    assert grp.offsets == [0, 2500, 5000]
    ok1, ts1, _ = one.init(fb.C172.TrimParameters(**p))
    ok2, ts2, _ = grp.init(fb.C172.TrimParameters(**p))
    assert (ok1 == ok2).all() and np.array_equal(ts1, ts2)
    one.set_input("elevator", one.get_input("elevator") + ele)
    grp.set_input("elevator", np.concatenate([s.get_input("elevator") for s in grp.shards]) + ele)
    one.step(2.0); grp.step(2.0)
    one.sync(); grp.sync()
    assert np.array_equal(one.x, grp.x)
    ref = one.log_stats()
    for f in ("h_e", "theta"):
        got = grp.log_stats(f, 0, 5)
        for k in range(5):
            assert got[k]["count"] == ref[f]["count"][k] and got[k]["min"] == ref[f]["min"][k] and got[k]["max"] == ref[f]["max"][k]
            assert abs(got[k]["sum"] - ref[f]["sum"][k]) <= 1e-12 * abs(ref[f]["sum"][k]) * 100
    one.close(); grp.close()
