"""cfg-4 members (BASELINE.json configs[3] at test size) and the rounding spread of the reference's own ground-contact
formulation, measured oracle against oracle.

The reference computes a strut's compression from the difference of two ECEF position vectors of magnitude 6.4e6 m
(lib/FlightPhysics/src/landinggear.jl:262-270), so one ulp of position (9.3e-10 m) is a relative 1e-7..1e-6 of a
millimetre-to-centimetre compression, and a member at rest sits on the friction-direction switch at 1e-3 m/s
(landinggear.jl:436-452). Two builds of the SAME restatement that differ only in rounding therefore drift apart on the
gear-driven states by far more than 1e-9. `oracle_spread` measures that: the checker build (-O2 -ffp-contract=off)
against (a) the same sources compiled with FMA contraction (-O2 -ffp-contract=fast -mfma: a portable build, the same on every host), (b) the checker started from a state
perturbed by one ulp, and (c) the checker with every state component nudged by one ulp in a random direction after every
step, which is the step-to-step difference any other correct implementation (such as a GPU kernel with its own operation
order, fused multiply-adds and reciprocal sequences) shows against it. The GPU-vs-oracle bound of tests/test_gpu_ground.py is
this measured spread times a stated factor."""
import numpy as np

SEED = 20261018
CHUNKS = [1, 49, 50, 100, 100] + [200] * 11          # 25 s, compared after every chunk
INPUT_FIELDS = ("aileron", "elevator", "rudder", "flaps", "brake_left", "brake_right", "throttle", "mixture", "eng_start", "eng_stop",
                "mixture_auto", "steering_engaged", "m_pilot", "m_copilot", "m_lpass", "m_rpass", "m_baggage", "trn_elevation", "trn_surface")


def cfg4_draws(n):
    """Per-member random draws of the takeoff / landing ensemble (first half takeoff, second half approach)."""
    import bench
    rng = np.random.default_rng(SEED + 40)
    lat = np.deg2rad(rng.uniform(-60, 60, n)); lon = rng.uniform(-np.pi, np.pi, n)
    d = dict(lat=lat, lon=lon, geoid=bench.geoid_height(lat, lon), elev=rng.uniform(0, 1500, n), surface=rng.integers(0, 3, n).astype(float),
             psi=rng.uniform(-np.pi, np.pi, n))
    d["h_gnd"] = d["elev"] + d["geoid"] + 1.9 + rng.uniform(0, 1, n)
    d["brake"] = rng.uniform(0, 0.3, n) * (rng.uniform(0, 1, n) < 0.3)
    d["rudder"] = rng.uniform(-0.2, 0.2, n)
    d["fuel"] = rng.uniform(0.2, 1, n)
    d["EAS"] = rng.uniform(30, 35, n)
    d["h_app"] = d["elev"] + d["geoid"] + rng.uniform(3, 12, n)
    return d


def cfg4_events(U, i, half, elevator_trim):
    """Input schedule of member i: takeoff = starter on, throttle ramp, rotate; approach = flare + idle + brakes."""
    if i < half:
        return [(0, U["FB_U_ENG_START"], 1.0), (300, U["FB_U_ENG_START"], 0.0), (300, U["FB_U_THROTTLE"], 0.3), (500, U["FB_U_THROTTLE"], 1.0),
                (500, U["FB_U_BRAKE_LEFT"], 0.0), (500, U["FB_U_BRAKE_RIGHT"], 0.0), (1500, U["FB_U_ELEVATOR"], 0.3)]
    return [(150, U["FB_U_THROTTLE"], 0.0), (150, U["FB_U_ELEVATOR"], float(elevator_trim) + 0.15),
            (900, U["FB_U_BRAKE_LEFT"], 0.5), (900, U["FB_U_BRAKE_RIGHT"], 0.5)]


def oracle_member(orc, U, IN, d, i, half, dt=0.01, fast=False):
    """Member i of the ensemble built by the oracle alone (C172.Init on the gear / trimmed on the approach path)."""
    o = orc.Sim(variant=0, kin=0, dt=dt, fast=fast)
    o.set_input(U["FB_U_TRN_ELEVATION"], d["elev"][i]); o.set_input(U["FB_U_TRN_SURFACE"], d["surface"][i])
    if i < half:
        o.init(q0=np.cos(d["psi"][i] / 2), q3=np.sin(d["psi"][i] / 2), lat=d["lat"][i], lon=d["lon"][i], h_e=d["h_gnd"][i], mixture=0.5,
               brake_left=d["brake"][i], brake_right=d["brake"][i], rudder=d["rudder"][i], fuel_load=d["fuel"][i])
    else:
        ok, _, _ = o.init_trim(EAS=d["EAS"][i], h_e=d["h_app"][i], flaps=1.0, gamma_wb_n=np.deg2rad(-3.0), psi_nb=d["psi"][i], lat=d["lat"][i], lon=d["lon"][i])
        assert ok
    for (s, f, v) in cfg4_events(U, i, half, o.get_input(U["FB_U_ELEVATOR"])):
        o.add_event(s, f, v)
    return o


def clone(orc, U, IN, src_inputs, events, x0, stall, eng, dt=0.01, fast=False):
    """An oracle Simulation started from a given state, inputs {name: value} and event list [(step, field, value)]."""
    o = orc.Sim(variant=0, kin=0, dt=dt, fast=fast)
    for f, v in src_inputs.items():
        o.set_input(IN[f], v)
    for (s, f, v) in events:
        o.add_event(s, f, v)
    o.init_state(x0, stall=int(stall), eng_state=int(eng))
    return o


def member_errors(xa, xb):
    """The metric of the ground tests: |a - b| / max(|b|, 1) per state; returns (all-state max, physical-state max) per member
    (physical = without the six friction-regulator integrators, rows 2..7)."""
    e = np.abs(xa - xb) / np.maximum(np.abs(xb), 1.0)
    return e.max(axis=0), np.delete(e, np.arange(2, 8), axis=0).max(axis=0)


def oracle_spread(orc, U, IN, x0, stall, eng, inputs, events, chunks=CHUNKS, dt=0.01):
    """Rounding spread of the reference formulation over the given members (columns of x0): checker build vs the FMA-contracted
    build, checker vs checker started one ulp away (h_e and the first attitude-quaternion component nudged by one ulp), and
    checker vs checker with one-ulp noise on every state component after every step.
    Returns per comparison instant the per-member (all-state, physical-state) errors of both experiments, and the statuses / WoW
    histories needed to tell which members the perturbation pushed across a discrete event."""
    n = x0.shape[1]
    base = [clone(orc, U, IN, {f: v[i] for f, v in inputs.items()}, events[i], x0[:, i], stall[i], eng[i], dt) for i in range(n)]
    fma = [clone(orc, U, IN, {f: v[i] for f, v in inputs.items()}, events[i], x0[:, i], stall[i], eng[i], dt, fast="fma") for i in range(n)]
    x1 = x0.copy()
    kin0 = 12                                              # C172S/WA layout: kinematics block starts at row 12, h_e is its 9th entry
    x1[kin0 + 8] = np.nextafter(x1[kin0 + 8], np.inf)
    x1[kin0] = np.nextafter(x1[kin0], np.inf)
    ulp = [clone(orc, U, IN, {f: v[i] for f, v in inputs.items()}, events[i], x1[:, i], stall[i], eng[i], dt) for i in range(n)]
    noisy = [clone(orc, U, IN, {f: v[i] for f, v in inputs.items()}, events[i], x0[:, i], stall[i], eng[i], dt) for i in range(n)]
    out = []
    for k, c in enumerate(chunks):
        for group in (base, fma, ulp):
            for o in group:
                o.step(c)
        for i, o in enumerate(noisy):
            o.step_noisy(c, 1000003 * k + i)
        xb = np.array([o.x for o in base]).T
        rec = {}
        for name, group in (("fma", fma), ("ulp", ulp), ("noisy", noisy)):
            xg = np.array([o.x for o in group]).T
            rec[name] = member_errors(xg, xb)
            rec[name + "_same_discrete"] = np.array([(a.status == b.status) and (a.discrete == b.discrete) for a, b in zip(group, base)])
        out.append(rec)
    return out
