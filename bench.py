#!/usr/bin/env python
"""bench.py — aircraft-RK4-steps/s of the batched C172 hot path (BASELINE.json metric).

Workload (BASELINE.json configs[2], SURVEY.md §8d "cfg 3", the config the metric's "1/2/4/8 B200" is quoted on):
a 2^20-member Cessna172Xv2/WA fly-by-wire ensemble, randomised trim (EAS, altitude, mass, heading, position; seed
20261018, counter-based per GLOBAL member index), random constant wind, gain-scheduled autopilot in EAS_clm / phi_beta
(turning climbs) with the control laws run every step (dt = Dt = 0.01 s). The ensemble is SHARDED by member across
the N GPUs (strong scaling: the same 2^20 members for every N; no data-path collective). Ensemble moments of four
outputs are accumulated inside the integrator once per simulated second and all-reduced over NCCL (N > 1) on a side
stream — the only exchange of the path.
One bench "step" = one pass of the hot path over the ensemble = ONE fb_step(1 s) = one fused launch advancing every
member by RK_PER_STEP = 100 RK4 steps; K = 600 steps is the full 600 s job (default K = 20).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg3|cfg2] [--members M]

Under torchrun (N > 1) every rank drives one GPU; rank 0 prints ONE JSON line.
`--workload cfg2` times BASELINE.json configs[1] instead (65,536 Cessna172Sv0 members per GPU, weak scaling: round 1's
line); a short cfg-2 timing is also attached to the cfg-3 line under "aux" at N = 1 so that the history survives.
`--impl reference` times the reference's own CPU path: the reference is Julia and Julia is not installed, so this is
the oracle port (oracle/, a C++ restatement) on all host cores over a bounded sample of the same workload; with
`julia` and baseline/_ref present it runs baseline/flightjl_cpu.jl (Flight.jl itself) instead.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TOTAL_CFG3 = 1 << 20
N_PER_GPU = 65536            # cfg 2
DT = 0.01
RK_PER_STEP = 100
SEED = 20261018
METRIC = "aircraft-RK4-steps/sec"
UNIT = "aircraft-steps/s"
# Algorithmic work per aircraft-step, measured with the op-counting oracle (oracle/liborc_count.so) under
# SURVEY.md §8(d)'s convention (add/sub/mul/cmp = 1, FMA = 2, div/sqrt/transcendental = 1), pinned by
# tests/test_oracle_golden.py::test_algorithmic_flop_count_of_record:
A_STEP_FLOP = 15695.0        # airborne C172S/WA: 4 RHS x 3835 + RK combine + f_step!
A_STEP_FLOP_CFG3 = 17390.0   # airborne C172X/WA with the control laws every step
A_STEP_FLOP_GROUND = 22989.0  # C172S/WA standing on all three wheels, engine running (cfg 4): 4 RHS x 5657 + RK combine + f_step!
STATS_FIELDS = ("h_e", "eas", "theta", "phi")


_GEOID = None


def geoid_height(lat, lon):
    """EGM96 undulation N(lat, lon) [m], bilinear on the 15' grid of the table blob (workload generation only):
    the workload altitudes are heights above mean sea level, TrimParameters wants ellipsoidal altitude h_e = h_msl + N."""
    global _GEOID
    if _GEOID is None:
        import struct
        blob = open(os.path.join(ROOT, "flight.jl_b200", "data", "c172_tables.bin"), "rb").read()
        ns = struct.unpack_from("<I", blob, 8)[0]
        for k in range(ns):
            nm, dt, nd, d0, d1, d2, d3, off, nb = struct.unpack_from("<24sII4IQQ", blob, 16 + 64 * k)
            if nm.rstrip(b"\0") == b"geoid":
                _GEOID = np.frombuffer(blob, dtype="<f4", count=nb // 4, offset=off).reshape(1441, 721).astype(np.float64)
    fi = (np.asarray(lat) + np.pi / 2) / (np.pi / 720)
    fj = np.mod(np.asarray(lon) + 2 * np.pi, 2 * np.pi) / (2 * np.pi / 1440)
    i = np.clip(np.floor(fi).astype(int), 0, 719); j = np.clip(np.floor(fj).astype(int), 0, 1439)
    u, w = fi - i, fj - j
    g = _GEOID
    return (1 - w) * ((1 - u) * g[j, i] + u * g[j, i + 1]) + w * ((1 - u) * g[j + 1, i] + u * g[j + 1, i + 1])


def _blocked(i0, n, keys, draw):
    """Counter-based generation: member i's values come from the stream of its 4096-member block, so any shard
    [i0, i0 + n) of the global ensemble sees the same values whatever the GPU count."""
    out = {k: np.empty(n) for k in keys}
    blk = 4096
    for b0 in range(i0 // blk * blk, i0 + n, blk):
        cols = draw(b0 // blk, blk)
        lo, hi = max(b0, i0), min(b0 + blk, i0 + n)
        for k in keys:
            out[k][lo - i0:hi - i0] = cols[k][lo - b0:hi - b0]
    return out


TRIM_KEYS = ("EAS", "h_e", "flaps", "psi_nb", "lat", "lon", "fuel_load", "m_pilot", "m_copilot", "m_lpass", "m_rpass", "m_baggage")


def member_params(i0, n):
    """cfg-2 trim parameters for global members i0 .. i0+n-1."""
    def draw(b, m):
        rng = np.random.default_rng([SEED, b])
        EAS = rng.uniform(30, 55, m)
        flaps = rng.choice([0.0, 1 / 3, 2 / 3, 1.0], m)
        flaps[EAS < 35] = 1.0
        flaps[EAS > 42] = 0.0     # level trim with flaps out is power-limited above ~45 m/s EAS
        h_msl = rng.uniform(50, 3050, m)
        lat = np.deg2rad(rng.uniform(-60, 60, m)); lon = rng.uniform(-np.pi, np.pi, m)
        return dict(EAS=EAS, h_e=h_msl + geoid_height(lat, lon), flaps=flaps, psi_nb=rng.uniform(-np.pi, np.pi, m),
                    lat=lat, lon=lon, fuel_load=rng.uniform(0.1, 1, m),
                    m_pilot=rng.uniform(0, 100, m), m_copilot=rng.uniform(0, 100, m), m_lpass=rng.uniform(0, 100, m),
                    m_rpass=rng.uniform(0, 100, m), m_baggage=rng.uniform(0, 100, m))
    return _blocked(i0, n, TRIM_KEYS, draw)


CFG3_KEYS = TRIM_KEYS + ("wind_n", "wind_e", "clm_ref", "phi_ref")


def cfg3_member_params(i0, n):
    """cfg-3 members i0 .. i0+n-1: trim as cfg 2 but flaps 0 and EAS ~ U[40,55]; wind N,E ~ U[-10,10]
    m/s; clm_ref ~ U[0.5,3] m/s, phi_ref ~ U[-30,30] deg for the EAS_clm / phi_beta modes (README.md:107-127)."""
    def draw(b, m):
        rng = np.random.default_rng([SEED, 3, b])
        lat = np.deg2rad(rng.uniform(-60, 60, m)); lon = rng.uniform(-np.pi, np.pi, m)
        return dict(EAS=rng.uniform(40, 55, m), h_e=rng.uniform(50, 3050, m) + geoid_height(lat, lon), flaps=np.zeros(m),
                    psi_nb=rng.uniform(-np.pi, np.pi, m), lat=lat, lon=lon, fuel_load=rng.uniform(0.1, 1, m),
                    m_pilot=rng.uniform(0, 100, m), m_copilot=rng.uniform(0, 100, m), m_lpass=rng.uniform(0, 100, m),
                    m_rpass=rng.uniform(0, 100, m), m_baggage=rng.uniform(0, 100, m),
                    wind_n=rng.uniform(-10, 10, m), wind_e=rng.uniform(-10, 10, m),
                    clm_ref=rng.uniform(0.5, 3.0, m), phi_ref=np.deg2rad(rng.uniform(-30, 30, m)))
    return _blocked(i0, n, CFG3_KEYS, draw)


LON_EAS_CLM, LAT_PHI_BETA = 7.0, 3.0     # c172x_ctl.jl:29-41, :733-741


def doublet_events(n, ele0, ail0, ele_f, ail_f):
    """(member, step, field, value) arrays of the elevator / aileron doublets (tutorial02.md:164-168 shape)."""
    t_ele = [(5.0, +0.1), (7.0, -0.1), (9.0, 0.0)]
    t_ail = [(20.0, +0.1), (22.0, -0.1), (24.0, 0.0)]
    member, step, field, value = [], [], [], []
    idx = np.arange(n)
    for (t, d) in t_ele:
        member.append(idx); step.append(np.full(n, int(round(t / DT)))); field.append(np.full(n, ele_f)); value.append(ele0 + d)
    for (t, d) in t_ail:
        member.append(idx); step.append(np.full(n, int(round(t / DT)))); field.append(np.full(n, ail_f)); value.append(ail0 + d)
    return np.concatenate(member), np.concatenate(step), np.concatenate(field), np.concatenate(value)


class Workload:
    """What differs between the two bench configurations; everything else (timing, e2e, roofline) is shared."""

    def __init__(self, name, rank, world, members):
        import flightbatch  # noqa: F401  (loads the package under its importable name)
        from flightbatch.dist import shard_range as dist_shard_range
        self.name = name
        if name == "cfg3":
            self.n_total = members or N_TOTAL_CFG3
            self.lo, self.hi = dist_shard_range(self.n_total, rank, world)
            self.scaling, self.a_step = "strong", A_STEP_FLOP_CFG3
            self.kernel = "step_kernel<C172X1,WA,AIR>"
            self.desc = (f"cfg3: {self.n_total:,}-member Cessna172Xv2/WA fly-by-wire ensemble sharded by member over the GPUs, randomized trim + "
                         "random constant wind, gain-scheduled autopilot EAS_clm / phi_beta every step (dt = Dt = 0.01), 600 s turning climbs")
        else:
            per = members or N_PER_GPU
            self.n_total = per * world
            self.lo, self.hi = rank * per, (rank + 1) * per
            self.scaling, self.a_step = "weak", A_STEP_FLOP
            self.kernel = "step_kernel<C172S,WA,AIR>"
            self.desc = "cfg2: 65,536-member Cessna172Sv0/WA ensemble per GPU, randomized trim (EAS, altitude, flaps, mass) + elevator/aileron doublets, dt=0.01"
        self.n = self.hi - self.lo
        self.p = None

    def make(self, fb, device, t_end, stats=True):
        """A trimmed handle over this rank's shard; returns (sim, n_trim_ok)."""
        n, lo = self.n, self.lo
        kw = dict(n=n, dt=DT, Dt=DT, t_end=t_end, device=device, member_offset=lo, steps_per_launch=RK_PER_STEP)
        if stats and hasattr(fb.BatchSimulation, "log_stats"):
            kw.update(saveat=RK_PER_STEP * DT, log_fields=STATS_FIELDS, log_members=False)
        if self.name == "cfg3":
            if self.p is None:
                self.p = cfg3_member_params(lo, n)
            p = self.p
            sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Xv2(fb.WA())), **kw)
            sim.set_input("wind_n", p["wind_n"]); sim.set_input("wind_e", p["wind_e"])
            ok, _, _ = sim.init(fb.C172.TrimParameters(**{k: p[k] for k in TRIM_KEYS}))
            sim.set_input("lon_mode_req", LON_EAS_CLM); sim.set_input("lat_mode_req", LAT_PHI_BETA)
            sim.set_input("lon_clm_ref", p["clm_ref"]); sim.set_input("lat_phi_ref", p["phi_ref"])
            self.step_inputs = [("lon_clm_ref", p["clm_ref"]), ("lat_phi_ref", p["phi_ref"])]
        else:
            if self.p is None:
                self.p = member_params(lo, n)
            sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), **kw)
            ok, _, _ = sim.init(fb.C172.TrimParameters(**self.p))
            ele_f, ail_f = fb.U["FB_U_ELEVATOR"], fb.U["FB_U_AILERON"]
            # No doublet schedule on the end-to-end handle: there the host uploads each second's control inputs itself
            # (trim value + doublet offset), which is what the schedule would have applied
            self.ele0, self.ail0 = sim.get_input(ele_f), sim.get_input(ail_f)
            self.ev = doublet_events(n, self.ele0, self.ail0, ele_f, ail_f)
            self.step_inputs = [("elevator", self.ele0), ("aileron", self.ail0)]
        return sim, int(ok.sum())

    def host_inputs(self, k_sim_second):
        """The per-member inputs in force during simulated second k (what the e2e leg uploads before that step)."""
        if self.name == "cfg3":
            return self.step_inputs
        t = float(k_sim_second)
        de = 0.1 if 5 <= t < 7 else (-0.1 if 7 <= t < 9 else 0.0)
        da = 0.1 if 20 <= t < 22 else (-0.1 if 22 <= t < 24 else 0.0)
        return [("elevator", self.ele0 + de), ("aileron", self.ail0 + da)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [t.strip() for t in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); smax.append(float(p[1])); pw.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def _oracle_members(workload, n_s):
    """n_s oracle Simulations of the workload's first members (the CPU arm's bounded sample)."""
    from oracle import orc
    import flightbatch as fb
    sims = []
    if workload == "cfg3":
        p = cfg3_member_params(0, n_s)
        for i in range(n_s):
            o = orc.Sim(variant=3, kin=0, dt=DT, Dt=DT, fast=True)
            o.set_input(fb.U["FB_U_WIND_N"], p["wind_n"][i]); o.set_input(fb.U["FB_U_WIND_E"], p["wind_e"][i])
            ok, _, _ = o.init_trim(**{k: float(p[k][i]) for k in TRIM_KEYS})
            o.set_input(fb.U["FB_U_LON_MODE_REQ"], LON_EAS_CLM); o.set_input(fb.U["FB_U_LAT_MODE_REQ"], LAT_PHI_BETA)
            o.set_input(fb.U["FB_U_LON_CLM_REF"], p["clm_ref"][i]); o.set_input(fb.U["FB_U_LAT_PHI_REF"], p["phi_ref"][i])
            if ok:
                sims.append(o)
    else:
        p = member_params(0, n_s)
        for i in range(n_s):
            o = orc.Sim(variant=0, kin=0, dt=DT, fast=True)
            ok, _, _ = o.init_trim(**{k: float(v[i]) for k, v in p.items()})
            if ok:
                sims.append(o)
    return sims


def cpu_baseline(threads, workload="cfg3", seconds_target=12.0, sims=None):
    """Oracle port on the host cores over a bounded sample of the same workload (the timing build: -O3 -march=native,
    FMA contraction allowed; the checker build used by the tests stays -O2 -ffp-contract=off)."""
    from oracle import orc
    orc.build()
    if sims is None:
        sims = _oracle_members(workload, 64 * max(1, threads))
    # calibrate, then run ~seconds_target of CPU work
    t0 = time.perf_counter(); orc.batch_step(sims, 20, threads); t1 = time.perf_counter()
    rate = len(sims) * 20 / (t1 - t0)
    n_steps = int(max(50, min(6000, seconds_target * rate / len(sims))))
    t0 = time.perf_counter(); orc.batch_step(sims, n_steps, threads); t1 = time.perf_counter()
    value = len(sims) * n_steps / (t1 - t0)
    model = "Cessna172Xv2/WA with the control laws every step" if workload == "cfg3" else "Cessna172Sv0/WA"
    return {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{len(sims)} members x {n_steps} RK4 steps of the {workload} ensemble ({model}; 5 RHS/step as the reference), "
                      f"C++ oracle port built {orc.FAST_FLAGS} on {threads} threads, {t1 - t0:.1f} s"}, (t1 - t0), n_steps * len(sims), sims


def julia_reference(args, wl_name):
    """Flight.jl itself (BASELINE.md §4 Tier A) when a Julia toolchain and the instantiated reference are present."""
    julia = shutil.which("julia")
    proj = os.path.join(ROOT, "baseline", "_ref")
    script = os.path.join(ROOT, "baseline", "flightjl_cpu.jl")
    if not julia or not os.path.isdir(proj) or not os.path.exists(script):
        return None
    threads = os.cpu_count() or 1
    # the members the GPU arm integrates (same counter-based draws), as a text table the script reads
    n_s = 64 * threads
    p = cfg3_member_params(0, n_s) if wl_name == "cfg3" else member_params(0, n_s)
    members = os.path.join(ROOT, "baseline", f"_members_{wl_name}.txt")
    keys = list(p.keys())
    with open(members, "w") as f:
        f.write(" ".join(keys) + "\n")
        for i in range(n_s):
            f.write(" ".join(repr(float(p[k][i])) for k in keys) + "\n")
    try:
        out = subprocess.run([julia, f"--project={proj}", f"--threads={threads}", script, wl_name, members, str(args.steps), str(args.warmup)],
                             capture_output=True, text=True, timeout=1500)
        for line in reversed(out.stdout.splitlines()):
            if line.startswith("{"):
                return json.loads(line)
    except Exception:
        return None
    return None


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = args.workload
    threads = os.cpu_count() or 1
    jl = julia_reference(args, wl)
    if jl is not None:
        cb = {"value": jl["value"], "unit": UNIT, "cores": jl.get("threads", threads), "kind": "reference", "sample": jl.get("sample", "")}
        value, t_total = jl["value"], jl.get("seconds", 0.0)
        note = "Flight.jl itself (baseline/flightjl_cpu.jl): one Simulation per member under Threads.@threads"
    else:
        sims = None
        for _ in range(max(1, args.warmup)):
            _, _, _, sims = cpu_baseline(threads, wl, seconds_target=0.5, sims=sims)
        t_total, work = 0.0, 0
        budget = max(1.0, min(8.0, 150.0 / max(1, args.steps)))
        cb = None
        for _ in range(args.steps):
            cb, dt_s, w, sims = cpu_baseline(threads, wl, seconds_target=budget, sims=sims)
            t_total += dt_s; work += w
        value = work / t_total
        cb["value"] = value
        note = "reference is Julia (not installed, cannot be run here): this arm is the C++ oracle port on all host cores"
    workload_desc = ("cfg3: Cessna172Xv2/WA fly-by-wire ensemble, randomized trim + wind, EAS_clm / phi_beta autopilot every step, dt = Dt = 0.01 (bounded CPU sample)"
                     if wl == "cfg3" else
                     "cfg2: Cessna172Sv0/WA ensemble, randomized trim + elevator/aileron doublets, dt=0.01 (bounded CPU sample)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True, "scaling": "strong" if wl == "cfg3" else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_desc, "members_total": N_TOTAL_CFG3 if wl == "cfg3" else N_PER_GPU * max(1, args.gpus), "rk4_steps_per_bench_step": RK_PER_STEP},
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": note}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def timed_leg(torch, dist, sim, stream, K, W, world, flush, comm_stats):
    """W warm-up steps, then K timed steps (CUDA events on the library's launch stream, L2 flushed between steps).
    Returns (per-step ms list, wall seconds, counters before/after, stats all-reduce wall ms list)."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(W):
        sim.step(RK_PER_STEP * DT)
        if comm_stats:
            sim.log_stats_allreduce()
    barrier()
    c0 = sim.counters()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    barrier()
    wall0 = time.perf_counter()
    for k in range(K):
        flush.fill_(float(k))                       # L2 flush between timed steps (256 MiB write), outside the event pair
        stream.wait_stream(torch.cuda.current_stream())
        starts[k].record(stream)
        sim.step(RK_PER_STEP * DT)                  # ONE fused launch: 100 RK4 steps x n members (+ the hand-over pass)
        ends[k].record(stream)
        if comm_stats:
            sim.log_stats_allreduce()               # NCCL all-reduce of this second's moments, on the library's side stream
    sim.sync()
    barrier()
    wall1 = time.perf_counter()
    c1 = sim.counters()
    ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    return ms, wall1 - wall0, c0, c1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="flightbatch")
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg2"])
    ap.add_argument("--members", type=int, default=0, help="cfg3: total members (sharded); cfg2: members per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--save-peak", default="", help="write the FP64 probe result and its clock record to this JSON file")
    ap.add_argument("--rk-per-step", type=int, default=0, help="diagnostic: RK4 steps fused per bench step / launch (default 100 = 1 s)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.rk_per_step > 0:
        global RK_PER_STEP
        RK_PER_STEP = args.rk_per_step
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import flightbatch as fb
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the engine has no CPU path"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, W = args.steps, max(args.warmup, 0)
    wl = Workload(args.workload, rank, world, args.members)
    n = wl.n
    have_stats = hasattr(fb.BatchSimulation, "log_stats")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- build + trim the shard (untimed)
    t_end = (K + W) * RK_PER_STEP * DT + 1.0
    sim, n_ok = wl.make(fb, local, t_end)
    if wl.name == "cfg2":
        sim.schedule_inputs(*wl.ev)
    def comm_setup(sim_):
        """The library owns the NCCL communicator of the statistics all-reduce; torch.distributed only carries its id."""
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.frombuffer(bytearray(fb.comm_unique_id()), dtype=torch.uint8).clone()
        uid = uid.cuda(); dist.broadcast(uid, 0)
        sim_.comm_init(world, rank, uid.cpu().numpy().tobytes())

    comm_stats = have_stats and world > 1
    if comm_stats:
        comm_setup(sim)
    stream = torch.cuda.ExternalStream(sim.stream_handle(), device=torch.device("cuda", local))
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")

    # ---------------------------------------------------------------- device-resident timing
    sampler = ClockSampler(local) if rank == 0 else None
    ms, wall_s, c0, c1 = timed_leg(torch, dist, sim, stream, K, W, world, flush, comm_stats)
    clocks = sampler.stop() if sampler else None
    dev_s = sum(ms) * 1e-3
    t = torch.tensor([dev_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s = float(t.item())
    status = sim.status()
    alive = int((status == 0).sum())
    tot = torch.tensor([alive, n_ok], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_members = float(tot[0].item())
    value = total_members * RK_PER_STEP * K / dev_s
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    rhs_per_member_step = (c1["rhs_evals"] - c0["rhs_evals"]) / max(1.0, alive * RK_PER_STEP * K)
    stats_line = None
    if have_stats:
        # the moments of the last timed second, after the all-reduce (global over all ranks when N > 1), fetched asynchronously
        t0 = time.perf_counter()
        m = sim.log_stats(first=W + K, count=1)
        stats_line = {"fields": list(STATS_FIELDS), "every_steps": RK_PER_STEP, "accumulated": "inside the integrator's log tap (warp reduce + one atomic set per warp)",
                      "allreduce": "NCCL (in-library communicator, side stream), once per bench step" if comm_stats else "none (one rank)",
                      "last_sample": {f: {"count": int(m[f]["count"][0]), "mean": float(m[f]["mean"][0]), "std": float(m[f]["std"][0]),
                                          "min": float(m[f]["min"][0]), "max": float(m[f]["max"][0])} for f in STATS_FIELDS},
                      "fetch_ms": (time.perf_counter() - t0) * 1e3}
        if comm_stats:
            stats_line["allreduce_latency_ms"] = sim.comm_time_ms() / max(1, K + W)
            stats_line["allreduce_latency_what"] = ("CUDA events around each all-reduce on the library's high-priority side stream: includes the wait for an SM "
                                                    "slot behind the next step's launch and for the slowest rank; it overlaps the stepping, whose rate is `value`")
    nx_state = sim.nx
    sim.close()

    # ---------------------------------------------------------------- end to end through the public API, host buffers
    # A fresh handle with the same members, so that this leg covers the same simulated interval (W warm-up seconds,
    # then K timed seconds). Per step: the step's per-member inputs from pinned host memory (H2D), fb_step(1 s), and
    # the whole ensemble state back into one of two pinned host arrays (D2H on its own stream, overlapping the next
    # step; every read has landed before the clock stops).
    e2e = None
    if not args.no_e2e:
        sim, _ = wl.make(fb, local, t_end)
        if comm_stats:
            comm_setup(sim)
        x_pin = [torch.empty((nx_state, n), dtype=torch.float64).pin_memory() for _ in range(2)]
        x_h = [t_.numpy() for t_ in x_pin]
        in_pin = {}

        def e2e_step(k_second, slot):
            for name, vals in wl.host_inputs(k_second):
                if name not in in_pin:
                    in_pin[name] = [torch.empty(n, dtype=torch.float64).pin_memory(), None]
                if in_pin[name][1] is not vals:                     # this second's inputs, written by the host when they change
                    in_pin[name][0].numpy()[:] = vals
                    in_pin[name][1] = vals
                sim.set_input(name, in_pin[name][0].numpy())        # H2D from pinned memory, every step
            sim.step(RK_PER_STEP * DT)
            if comm_stats:
                sim.log_stats_allreduce()
            sim.get_state_async(x_h[slot])                          # D2H of the step's result, pipelined

        for k in range(W):                                          # warm-up runs the same body (first-use allocations: pinned rows, device snapshot)
            e2e_step(k, k % 2)
        sim.wait_state()
        sim.sync()
        barrier()
        e2e_ms = []
        e0 = time.perf_counter()
        for k in range(K):
            s0 = time.perf_counter()
            e2e_step(W + k, k % 2)
            e2e_ms.append((time.perf_counter() - s0) * 1e3)
        sim.wait_state()
        sim.sync()
        barrier()
        e1 = time.perf_counter()
        te = torch.tensor([e1 - e0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        te = float(te.item())
        alive_e = torch.tensor([float((sim.status() == 0).sum())], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(alive_e, op=dist.ReduceOp.SUM)
        e2e = {"value": float(alive_e.item()) * RK_PER_STEP * K / te, "unit": UNIT, "h2d_bytes_per_step": len(wl.step_inputs) * n * 8,
               "d2h_bytes_per_step": nx_state * n * 8, "steps": K, "seconds": te,
               "ms_per_step_min_med_max": [min(e2e_ms), float(np.median(e2e_ms)), max(e2e_ms)],
               "what": "per step: fb_set_input of the step's two per-member input rows from pinned host arrays + fb_step(1 s) + fb_get_state_async of the "
                       "whole state into one of two pinned host arrays (device snapshot, D2H on the copy stream overlapping the next step; all reads landed "
                       "before the clock stops); wall clock of ONE batch of K steps on a fresh handle over the same simulated interval, max over ranks; bytes are per rank"}
        sim.close()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        ps = ClockSampler(local)
        fp64_peak = fb.measure_fp64_peak(local)
        peak_clocks = ps.stop()
        if args.save_peak:
            json.dump({"fp64_fma_tflops": fp64_peak, "how": "fb_measure_fp64_peak: 8 independent FMA chains per thread, 256 threads x 8 CTAs per SM, best of 5 launches, FMA = 2 flop",
                       "clocks_during_probe": peak_clocks, "gpu": torch.cuda.get_device_name(local)}, open(args.save_peak, "w"), indent=1)
        per_gpu_rate = value / world
        achieved_tf = per_gpu_rate * wl.a_step * 1e-12
        prof = {}
        ppath = os.path.join(ROOT, "profiles", f"r02_{wl.name}_kernel.json")
        if os.path.exists(ppath):
            try:
                prof = json.load(open(ppath))
            except Exception:
                prof = {}
        # Algorithmic HBM bytes per member and launch of 100 fused steps: state in + out; the input rows the variant reads (27 for
        # C172S; + 7 actuator commands, 19 control-law and 9 guidance inputs for the fly-by-wire variants) and, there, the 7
        # commands written back and the 48 rows of controller memory in + out; ~40 B of discrete state, status, event cursor and
        # ownership word. (The measured DRAM traffic of the cfg-3 launch is 1.5x this: the 14 tap rows and partially used sectors.)
        ctl_variant = wl.name == "cfg3"
        state_bytes = (2 * nx_state * 8 + (62 if ctl_variant else 27) * 8 + (7 * 8 + 2 * 48 * 8 if ctl_variant else 0) + 40) * n
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * dev_s / K, "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.desc, "members_total": wl.n_total, "members_per_gpu": n, "members_alive": total_members,
                       "trim_ok": float(tot[1].item()), "rk4_steps_per_bench_step": RK_PER_STEP, "rhs_evals_per_member_step": rhs_per_member_step,
                       "l2": "flushed between timed steps (256 MiB write outside the event pair); the kernel is FP64-issue bound, not memory bound",
                       "timing": "CUDA events on the library's launch stream, sum over steps, max over ranks", "wall_s": wall_s,
                       "ms_per_step_min_max": [min(ms), max(ms)],
                       "launches_per_step": "2: airborne-layout kernel + the hand-over pass of the complete kernel, pipelined on the library's side stream (148 CTAs that return at once while no member is near the terrain)"},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak if fp64_peak else None,
                         "traffic": prof.get("dram_bytes_per_launch"), "kernel": wl.kernel,
                         "executed_fp64_tflops": (prof["fp64_thread_flop_per_member_step"] * per_gpu_rate * 1e-12) if "fp64_thread_flop_per_member_step" in prof else None,
                         "pipe_fp64_pct_ncu": prof.get("pipe_fp64_pct"),
                         "note": f"path is FP64-FMA-pipe bound, not HBM/tensor (SURVEY.md §8d); peak = FP64 FMA probe measured in this run (MEASURED_PEAKS.json carries "
                                 f"no FP64 figure; kept in profiles/fp64_peak.json); achieved = {wl.a_step:,.0f} unit-FLOP/aircraft-step (the reference's arithmetic, op-counting "
                                 "oracle) x steps/s per GPU; executed_fp64_tflops / pipe_fp64_pct_ncu / traffic are from the committed ncu capture of this kernel "
                                 f"(profiles/r02_{wl.name}_kernel.json), not from this run",
                         "hbm": {"achieved_gbs": state_bytes / (dev_s / K) * 1e-9, "peak_gbs": peaks.get("hbm_gbs"), "algorithmic_bytes_per_launch": state_bytes}},
        }
        if stats_line:
            line["stats"] = stats_line
        if world == 1 and wl.name == "cfg3" and not args.no_aux:
            # round-1 line (cfg 2, 65,536 Cessna172Sv0 members) kept as an auxiliary figure: device-resident only
            w2 = Workload("cfg2", 0, 1, 0)
            s2, _ = w2.make(fb, local, 100.0, stats=False)
            s2.schedule_inputs(*w2.ev)
            st2 = torch.cuda.ExternalStream(s2.stream_handle(), device=torch.device("cuda", local))
            ms2, _, _, _ = timed_leg(torch, dist, s2, st2, 10, 3, 1, flush, False)
            v2 = float((s2.status() == 0).sum()) * RK_PER_STEP * 10 / (sum(ms2) * 1e-3)
            s2.close()
            line["aux"] = {"cfg2": {"value": v2, "unit": UNIT, "ms_per_step": sum(ms2) / 10, "steps": 10, "members": N_PER_GPU,
                                    "frac": v2 * A_STEP_FLOP * 1e-12 / fp64_peak if fp64_peak else None,
                                    "what": "BASELINE.json configs[1], round 1's bench line: 65,536 Cessna172Sv0/WA members, doublets, device-resident"}}
        if world == 1 and not args.no_cpu_baseline:
            cb, _, _, _ = cpu_baseline(os.cpu_count() or 1, wl.name, seconds_target=12.0)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
