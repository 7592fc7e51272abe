"""The rounding spread of the reference's ground-contact formulation, measured on the CPU oracle alone (no GPU).

North star: states agree to <= 1e-9 relative. On the landing-gear path that cannot hold between ANY two implementations
whose rounding differs: this test runs the cfg-4 takeoff / landing members (BASELINE.json configs[3] at test size) on two
builds of the same restatement — the checker (-O2 -ffp-contract=off) and a second build of the same sources (-O2 -ffp-contract=fast -mfma, FMA contraction
allowed) — and on the checker started one ulp away, and records how far they drift apart. tests/test_gpu_ground.py bounds the
GPU-vs-oracle error by this spread (x5), measured there on the members it compares. Reference: landinggear.jl:262-275 (strut
compression from the difference of two 6.4e6 m ECEF vectors), :436-452 (friction direction switch at 1e-3 m/s)."""
import numpy as np

import ground_util as gu


def test_reference_formulation_is_ill_conditioned_on_the_ground(fb, orc):
    n = 24
    half = n // 2
    d = gu.cfg4_draws(n)
    ms = [gu.oracle_member(orc, fb.U, fb.INPUT_NAMES, d, i, half) for i in range(n)]
    x0 = np.array([o.x for o in ms]).T
    disc = np.array([o.discrete for o in ms])
    inputs = {f: np.array([o.get_input(fb.INPUT_NAMES[f]) for o in ms]) for f in gu.INPUT_FIELDS}
    events = [gu.cfg4_events(fb.U, i, half, ms[i].get_input(fb.U["FB_U_ELEVATOR"])) for i in range(n)]
    chunks = [1, 49, 150, 300, 500, 500, 1000]                       # 25 s
    sp = gu.oracle_spread(orc, fb.U, fb.INPUT_NAMES, x0, disc[:, 0], disc[:, 1], inputs, events, chunks=chunks)
    fma_max = [float(r["fma"][1].max()) for r in sp]
    ulp_max = [float(r["ulp"][1].max()) for r in sp]
    fma_med = [float(np.median(r["fma"][0])) for r in sp]
    noisy_max = [float(r["noisy"][1].max()) for r in sp]
    noisy_med = [float(np.median(r["noisy"][0])) for r in sp]
    print("t [steps]      ", np.cumsum(chunks).tolist())
    print("FMA build  max ", ["%.1e" % v for v in fma_max])
    print("FMA build  med ", ["%.1e" % v for v in fma_med])
    print("1-ulp start max", ["%.1e" % v for v in ulp_max])
    print("1-ulp/step  max", ["%.1e" % v for v in noisy_max])
    print("1-ulp/step  med", ["%.1e" % v for v in noisy_med])
    # one step in, before any strut carries load for long, the builds still agree to rounding
    assert fma_max[0] < 1e-13 and ulp_max[0] < 1e-13
    # on the gear they do not: two roundings of the SAME algorithm are more than a thousand times further apart than 1e-9 ...
    assert max(fma_max) > 1e-6
    # ... and one ulp of initial position / attitude is amplified by more than six orders of magnitude
    assert max(ulp_max) > 1e-10
    # the drift is rounding noise, not divergence: it stays bounded, and no discrete event (WoW, engine state, stall, status)
    # moves between the builds on these members
    assert max(fma_max) < 2e-2 and max(ulp_max) < 2e-2 and max(noisy_max) < 2e-2
    assert max(noisy_max) > 1e-6
    assert all(r["fma_same_discrete"].all() and r["ulp_same_discrete"].all() and r["noisy_same_discrete"].all() for r in sp)
