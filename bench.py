#!/usr/bin/env python
"""bench.py — aircraft-RK4-steps/s of the batched C172 hot path (BASELINE.json metric).

Workload (BASELINE.json configs[1], SURVEY.md §8d "cfg 2"): a 65,536-member Cessna172Sv0/WA ensemble per GPU,
randomised trim (EAS, altitude, flaps, mass, heading, position; seed 20261018, counter-based per global member
index), elevator doublet at t in [5,7),[7,9) s and aileron doublet at [20,22),[22,24) s, dt = Δt = 0.01 s.
One bench "step" = one pass of the hot path over the ensemble = ONE fused kernel launch advancing every member by
RK_PER_STEP = 100 RK4 steps (1 s of simulated time); the default K = 60 steps is the full 60 s job.
Multi-GPU: the ensemble shards by member (no data-path collective) -> weak scaling, 65,536 members per rank.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Under torchrun (N > 1) every rank drives one GPU; rank 0 prints ONE JSON line.
`--impl reference` times the reference's own CPU path: the reference is Julia and Julia is not installed, so this
is the oracle port (oracle/, a C++ restatement) on all host cores, bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_GPU = 65536
DT = 0.01
RK_PER_STEP = 100
SEED = 20261018
METRIC = "aircraft-RK4-steps/sec"
UNIT = "aircraft-steps/s"
# algorithmic work per aircraft-step, measured with the op-counting oracle (oracle/liborc_count.so) under
# SURVEY.md §8(d)'s convention (add/sub/mul/cmp = 1, FMA = 2, div/sqrt/transcendental = 1):
# 4 RHS x 3835 + RK combine + f_step! = 15,695 unit-FLOP for an airborne C172S/WA member.
A_STEP_FLOP = 15695.0


_GEOID = None


def geoid_height(lat, lon):
    """EGM96 undulation N(lat, lon) [m], bilinear on the 15' grid of the table blob (workload generation only):
    the cfg-2 altitudes are heights above mean sea level, TrimParameters wants ellipsoidal altitude h_e = h_msl + N."""
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


def member_params(i0, n):
    """cfg-2 trim parameters for global members i0 .. i0+n-1; counter-based so sharding does not change values."""
    out = {k: np.empty(n) for k in ("EAS", "h_e", "flaps", "psi_nb", "lat", "lon", "fuel_load", "m_pilot", "m_copilot",
                                    "m_lpass", "m_rpass", "m_baggage")}
    blk = 4096
    for b0 in range(i0 // blk * blk, i0 + n, blk):
        rng = np.random.default_rng([SEED, b0 // blk])
        m = blk
        EAS = rng.uniform(30, 55, m)
        flaps = rng.choice([0.0, 1 / 3, 2 / 3, 1.0], m)
        flaps[EAS < 35] = 1.0
        flaps[EAS > 42] = 0.0     # level trim with flaps out is power-limited above ~45 m/s EAS
        h_msl = rng.uniform(50, 3050, m)
        lat = np.deg2rad(rng.uniform(-60, 60, m)); lon = rng.uniform(-np.pi, np.pi, m)
        cols = dict(EAS=EAS, h_e=h_msl + geoid_height(lat, lon), flaps=flaps, psi_nb=rng.uniform(-np.pi, np.pi, m),
                    lat=lat, lon=lon, fuel_load=rng.uniform(0.1, 1, m),
                    m_pilot=rng.uniform(0, 100, m), m_copilot=rng.uniform(0, 100, m), m_lpass=rng.uniform(0, 100, m),
                    m_rpass=rng.uniform(0, 100, m), m_baggage=rng.uniform(0, 100, m))
        lo, hi = max(b0, i0), min(b0 + blk, i0 + n)
        for k in out:
            out[k][lo - i0:hi - i0] = cols[k][lo - b0:hi - b0]
    return out


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
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [t.strip() for t in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); smax.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(threads, seconds_target=12.0, impl_line=False):
    """Oracle port on the host cores over a bounded sample of the same workload."""
    from oracle import orc
    orc.build()
    n_s = 64 * max(1, threads)
    p = member_params(0, n_s)
    sims = []
    for i in range(n_s):
        o = orc.Sim(variant=0, kin=0, dt=DT)
        ok, _, _ = o.init_trim(**{k: float(v[i]) for k, v in p.items()})
        if ok:
            sims.append(o)
    # calibrate, then run ~seconds_target of CPU work
    t0 = time.perf_counter(); orc.batch_step(sims, 20, threads); t1 = time.perf_counter()
    rate = len(sims) * 20 / (t1 - t0)
    n_steps = int(max(50, min(6000, seconds_target * rate / len(sims))))
    t0 = time.perf_counter(); orc.batch_step(sims, n_steps, threads); t1 = time.perf_counter()
    value = len(sims) * n_steps / (t1 - t0)
    return {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{len(sims)} members x {n_steps} RK4 steps of the cfg-2 ensemble (5 RHS/step as the reference), "
                      f"C++ oracle port on {threads} threads, {t1 - t0:.1f} s"}, (t1 - t0), n_steps * len(sims)


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per = []
    cb = None
    for _ in range(args.warmup):
        cpu_baseline(threads, seconds_target=1.0)
    t_total, work = 0.0, 0
    budget = max(1.0, min(8.0, 150.0 / max(1, args.steps)))
    for _ in range(args.steps):
        cb, dt_s, w = cpu_baseline(threads, seconds_target=budget)
        t_total += dt_s; work += w
    value = work / t_total
    cb["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg2: Cessna172Sv0/WA ensemble, randomized trim + elevator/aileron doublets, dt=0.01 (bounded CPU sample)",
                       "members_per_gpu": N_PER_GPU, "rk4_steps_per_bench_step": RK_PER_STEP},
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference is Julia (not installed, cannot be run here): this arm is the C++ oracle port on all host cores"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="flightbatch")
    ap.add_argument("--members", type=int, default=N_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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
    n = args.members
    K, W = args.steps, max(args.warmup, 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- build + trim the ensemble (untimed)
    p = member_params(rank * n, n)
    t_end = (K + W) * RK_PER_STEP * DT * 2 + 10.0
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=DT, t_end=t_end, device=local,
                             member_offset=rank * n, steps_per_launch=RK_PER_STEP)
    ok, ts, cost = sim.init(fb.C172.TrimParameters(**p))
    n_ok = int(ok.sum())
    ele_f, ail_f = fb.U["FB_U_ELEVATOR"], fb.U["FB_U_AILERON"]
    ev = doublet_events(n, sim.get_input(ele_f), sim.get_input(ail_f), ele_f, ail_f)
    sim.schedule_inputs(*ev)
    stream = torch.cuda.ExternalStream(sim.stream_handle(), device=torch.device("cuda", local))
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")

    # ---------------------------------------------------------------- device-resident timing
    for _ in range(W):
        sim.step(RK_PER_STEP * DT)
    barrier()
    c0 = sim.counters()
    sampler = ClockSampler(local) if rank == 0 else None
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    barrier()
    wall0 = time.perf_counter()
    for k in range(K):
        flush.fill_(float(k))                       # L2 flush between timed steps (256 MiB write), outside the event pair
        stream.wait_stream(torch.cuda.current_stream())
        starts[k].record(stream)
        sim.step(RK_PER_STEP * DT)                  # ONE fused launch: 100 RK4 steps x n members
        ends[k].record(stream)
    sim.sync()
    barrier()
    wall1 = time.perf_counter()
    clocks = sampler.stop() if sampler else None
    c1 = sim.counters()
    ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    dev_s = sum(ms) * 1e-3
    t = torch.tensor([dev_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s = float(t.item())
    status = sim.status()
    alive = int((status == 0).sum())
    tot_alive = torch.tensor([alive, n_ok], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot_alive, op=dist.ReduceOp.SUM)
    total_members = float(tot_alive[0].item())
    value = total_members * RK_PER_STEP * K / dev_s
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    rhs_per_member_step = (c1["rhs_evals"] - c0["rhs_evals"]) / max(1.0, alive * RK_PER_STEP * K)

    # ---------------------------------------------------------------- end to end through the public API, host buffers
    # A second handle with the same members, trims and doublet schedule, so that this leg covers the same simulated
    # interval as the device-resident leg (W warm-up seconds, then Ke timed seconds) instead of continuing after it.
    Ke = min(K, 20)
    nx_state = sim.nx
    sim.close()
    ele_h = ail_h = None
    # the caller-owned pinned host buffers the state lands in, alternating: step k's read-back (fb_get_state_async: device
    # snapshot + D2H on the copy stream) overlaps step k+1's launch; every read has landed before the clock stops
    x_pin = [torch.empty((nx_state, n), dtype=torch.float64).pin_memory() for _ in range(2)]
    x_h = [t.numpy() for t in x_pin]
    batches = []
    for rep in range(2):   # two batches of Ke steps, each on a fresh handle; the faster one is reported (wall-clock noise guard)
        sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=DT, t_end=t_end, device=local,
                                 member_offset=rank * n, steps_per_launch=RK_PER_STEP)
        sim.init(fb.C172.TrimParameters(**p))
        sim.schedule_inputs(*ev)
        if ele_h is None:
            ele_h = torch.from_numpy(sim.get_input(ele_f).copy()).pin_memory()
            ail_h = torch.from_numpy(sim.get_input(ail_f).copy()).pin_memory()
        for _ in range(W):
            sim.step(RK_PER_STEP * DT)
        sim.sync()
        barrier()
        e2e_ms = []
        e0 = time.perf_counter()
        for k in range(Ke):
            s0 = time.perf_counter()
            sim.set_input(ele_f, ele_h.numpy())         # H2D: this second's control inputs, one value per member
            sim.set_input(ail_f, ail_h.numpy())
            sim.step(RK_PER_STEP * DT)
            sim.get_state_async(x_h[k % 2])             # D2H: the full ensemble state (the step's result), pipelined
            e2e_ms.append((time.perf_counter() - s0) * 1e3)
        sim.wait_state()                                # the last read-back
        sim.sync()
        barrier()
        e1 = time.perf_counter()
        te = torch.tensor([e1 - e0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        batches.append((float(te.item()), e2e_ms))
        if rep == 0:
            sim.close()
    te_best, e2e_ms = min(batches, key=lambda b: b[0])
    e2e_value = total_members * RK_PER_STEP * Ke / te_best
    h2d = 2 * n * 8
    d2h = sim.nx * n * 8

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        fp64_peak = fb.measure_fp64_peak(local)
        per_gpu_rate = value / world
        achieved_tf = per_gpu_rate * A_STEP_FLOP * 1e-12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        state_bytes = (2 * sim.nx * 8 + 40 * 8) * n   # state in+out + inputs read, per launch
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * dev_s / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg2: 65,536-member Cessna172Sv0/WA ensemble per GPU, randomized trim (EAS, altitude, flaps, mass) + elevator/aileron doublets, dt=0.01",
                       "members_per_gpu": n, "members_alive": total_members, "trim_ok": float(tot_alive[1].item()), "rk4_steps_per_bench_step": RK_PER_STEP,
                       "rhs_evals_per_member_step": rhs_per_member_step,
                       "l2": "flushed between timed steps (256 MiB write outside the event pair); state+inputs (56 MB) < L2, kernel is FP64-bound",
                       "timing": "CUDA events on the library's launch stream, sum over steps, max over ranks", "wall_s": wall1 - wall0,
                       "ms_per_step_min_max": [min(ms), max(ms)],
                       "launches_per_step": "2: airborne-layout kernel + hand-over pass of the complete kernel (returns at once when no member is near the terrain)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                    "ms_per_step_min_med_max": [min(e2e_ms), float(np.median(e2e_ms)), max(e2e_ms)],
                    "batches_s": [b[0] for b in batches],
                    "what": "per step: fb_set_input(elevator, aileron) from pinned host arrays + fb_step(1 s) + fb_get_state_async of the whole state into one of two pinned host arrays (device snapshot, D2H on the copy stream overlapping the next step; all reads landed before the clock stops), wall clock; faster of two batches, each on a fresh handle over the same simulated interval"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak if fp64_peak else None,
                         "traffic": traffic, "kernel": "step_kernel<C172S,WA,AIR>",
                         "note": "path is FP64-FMA-pipe bound, not HBM/tensor (SURVEY.md §8d); peak = FP64 FMA probe measured in this run "
                                 "(MEASURED_PEAKS.json carries no FP64 figure); achieved = 15,695 unit-FLOP/aircraft-step x steps/s per GPU",
                         "hbm": {"achieved_gbs": state_bytes / (dev_s / K) * 1e-9, "peak_gbs": peaks.get("hbm_gbs"), "algorithmic_bytes_per_launch": state_bytes}},
        }
        if world == 1 and not args.no_cpu_baseline:
            cb, _, _ = cpu_baseline(os.cpu_count() or 1, seconds_target=12.0)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    sim.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
