#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations (bench.py is cfg 2, the contract line; these are the
parity-test configurations at or near full size, timed the same way: CUDA events on the library's launch stream,
warm-up first, one JSON line per configuration).

    python tools/bench_configs.py cfg3 [--members N] [--seconds S]     Cessna172Xv2 autopilot ensemble (EAS_clm / phi_beta, wind)
    python tools/bench_configs.py cfg4 [--members N] [--seconds S]     takeoff / landing ensemble (gear contact, engine start)
    python tools/bench_configs.py cfg5 [--members N]                   envelope sweep: batched trim + linearize
    python tools/bench_configs.py lss  [--members N] [--seconds S]     LinearizedSS ensemble (warp-per-member linear RK4)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload generators, geoid)

DT = 0.01
SEED = 20261018


def timed_steps(fb, sim, seconds, chunk_s=1.0, warm=3):
    import torch
    stream = torch.cuda.ExternalStream(sim.stream_handle())
    for _ in range(warm):
        sim.step(chunk_s)
    sim.sync()
    c0 = sim.counters()
    k = int(round(seconds / chunk_s))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(k):
        sim.step(chunk_s)
    e1.record(stream)
    sim.sync()
    c1 = sim.counters()
    ms = e0.elapsed_time(e1)
    steps = k * int(round(chunk_s / DT))
    return ms, steps, c0, c1


def cfg3(fb, n, seconds):
    rng = np.random.default_rng(SEED + 3)
    lat = np.deg2rad(rng.uniform(-60, 60, n)); lon = rng.uniform(-np.pi, np.pi, n)
    p = dict(EAS=rng.uniform(40, 55, n), h_e=rng.uniform(50, 3050, n) + bench.geoid_height(lat, lon), psi_nb=rng.uniform(-np.pi, np.pi, n),
             lat=lat, lon=lon, fuel_load=rng.uniform(0.1, 1, n), m_pilot=rng.uniform(0, 100, n), m_copilot=rng.uniform(0, 100, n),
             m_lpass=rng.uniform(0, 100, n), m_rpass=rng.uniform(0, 100, n), m_baggage=rng.uniform(0, 100, n))
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Xv2(fb.WA())), n=n, dt=DT, Dt=DT, t_end=1e6, steps_per_launch=100)
    sim.set_input("wind_n", rng.uniform(-10, 10, n)); sim.set_input("wind_e", rng.uniform(-10, 10, n))
    ok, _, _ = sim.init(fb.C172.TrimParameters(**p))
    sim.set_input("lon_mode_req", 7.0); sim.set_input("lat_mode_req", 3.0)          # EAS_clm, phi_beta
    sim.set_input("lon_clm_ref", rng.uniform(0.5, 3.0, n)); sim.set_input("lat_phi_ref", np.deg2rad(rng.uniform(-30, 30, n)))
    ms, steps, c0, c1 = timed_steps(fb, sim, seconds)
    alive = int((sim.status() == 0).sum())
    modes = (np.bincount(sim.get_output("lon_mode").astype(int), minlength=9).tolist(), np.bincount(sim.get_output("lat_mode").astype(int), minlength=5).tolist())
    sim.close()
    return {"config": "cfg3: Cessna172Xv2/WA, EAS_clm + phi_beta autopilot, random constant wind, dt = Dt = 0.01", "members": n, "trim_ok": int(ok.sum()),
            "alive": alive, "rk4_steps": steps, "ms": ms, "aircraft_steps_per_s": alive * steps / (ms * 1e-3),
            "rhs_per_member_step": (c1["rhs_evals"] - c0["rhs_evals"]) / max(1, alive * steps), "kernel_launches": c1["kernel_launches"] - c0["kernel_launches"],
            "lon_mode_hist": modes[0], "lat_mode_hist": modes[1]}


def cfg4(fb, n, seconds):
    # load both integrator kernels before the clock starts (CUDA loads a kernel's image at its first launch): a throw-away
    # 64-member ensemble dropped on the gear, so that the timed run itself can start at t = 0 of the takeoff / landing sequence
    w = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=64, dt=DT, t_end=10.0)
    w.init(fb.C172.Init(h_e=2.0 + float(bench.geoid_height(0.0, 0.0))))
    w.step(0.05); w.sync(); w.close()
    rng = np.random.default_rng(SEED + 40)
    half = n // 2
    lat = np.deg2rad(rng.uniform(-60, 60, n)); lon = rng.uniform(-np.pi, np.pi, n)
    geoid = bench.geoid_height(lat, lon)
    elev = rng.uniform(0, 1500, n); surface = rng.integers(0, 3, n).astype(float); psi = rng.uniform(-np.pi, np.pi, n)
    U = fb.U
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=DT, t_end=1e6, steps_per_launch=100)
    sim.set_input("trn_elevation", elev); sim.set_input("trn_surface", surface)
    h_gnd = elev + geoid + 1.9 + rng.uniform(0, 1, n)
    brake = rng.uniform(0, 0.3, n) * (rng.uniform(0, 1, n) < 0.3)
    sim.init(fb.C172.Init(q0=np.cos(psi / 2), q3=np.sin(psi / 2), lat=lat, lon=lon, h_e=h_gnd, mixture=0.5, brake_left=brake, brake_right=brake,
                          rudder=rng.uniform(-0.2, 0.2, n), fuel_load=rng.uniform(0.2, 1, n)))
    xA = sim.x; stA, enA = sim.discrete()
    names = ("throttle", "elevator", "aileron", "rudder", "flaps", "brake_left", "brake_right", "mixture", "m_pilot", "m_copilot", "m_lpass", "m_rpass", "m_baggage")
    inA = {f: sim.get_input(f) for f in names}
    ok, _, _ = sim.init(fb.C172.TrimParameters(EAS=rng.uniform(30, 35, n), h_e=elev + geoid + rng.uniform(3, 12, n), flaps=1.0,
                                               gamma_wb_n=np.deg2rad(-3.0), psi_nb=psi, lat=lat, lon=lon))
    xB = sim.x; stB, enB = sim.discrete()
    inB = {f: sim.get_input(f) for f in names}
    for f in names:
        sim.set_input(f, np.concatenate([inA[f][:half], inB[f][half:]]))
    sim.init_state(np.concatenate([xA[:, :half], xB[:, half:]], axis=1), stall=np.concatenate([stA[:half], stB[half:]]),
                   eng_state=np.concatenate([enA[:half], enB[half:]]))
    a, b = np.arange(half), np.arange(half, n)
    ev = [(a, 0, U["FB_U_ENG_START"], 1.0), (a, 300, U["FB_U_ENG_START"], 0.0), (a, 300, U["FB_U_THROTTLE"], 0.3), (a, 500, U["FB_U_THROTTLE"], 1.0),
          (a, 500, U["FB_U_BRAKE_LEFT"], 0.0), (a, 500, U["FB_U_BRAKE_RIGHT"], 0.0), (a, 1500, U["FB_U_ELEVATOR"], 0.3),
          (b, 150, U["FB_U_THROTTLE"], 0.0), (b, 150, U["FB_U_ELEVATOR"], inB["elevator"][half:] + 0.15),
          (b, 900, U["FB_U_BRAKE_LEFT"], 0.5), (b, 900, U["FB_U_BRAKE_RIGHT"], 0.5)]
    sim.schedule_inputs(np.concatenate([m for m, _, _, _ in ev]), np.concatenate([np.full(len(m), s) for m, s, _, _ in ev]),
                        np.concatenate([np.full(len(m), f) for m, _, f, _ in ev]),
                        np.concatenate([np.broadcast_to(np.asarray(v, float), (len(m),)) for m, _, _, v in ev]))
    ms, steps, c0, c1 = timed_steps(fb, sim, seconds, warm=0)
    st = sim.status()
    wow = np.maximum.reduce([sim.get_output(k) for k in ("wow_left", "wow_right", "wow_nose")])
    out = {"config": "cfg4: Cessna172Sv0/WA takeoff rolls (engine start, throttle ramp, rotation) + 3-degree approaches to touchdown and roll-out",
           "members": n, "rk4_steps": steps, "ms": ms, "aircraft_steps_per_s": float((st == 0).sum()) * steps / (ms * 1e-3), "status_hist": np.bincount(st, minlength=7).tolist(),
           "on_ground_at_end": float(wow.mean()), "rhs_per_member_step": (c1["rhs_evals"] - c0["rhs_evals"]) / max(1, n * steps),
           "kernel_launches": c1["kernel_launches"] - c0["kernel_launches"]}
    # roofline of the ground path: unit-FLOP of a step on all three wheels with the engine running (op-counting oracle,
    # tests/test_oracle_golden.py) against the FP64 FMA probe of this device
    peak = fb.measure_fp64_peak() if hasattr(fb, "measure_fp64_peak") else None
    if peak:
        ach = out["aircraft_steps_per_s"] * bench.A_STEP_FLOP_GROUND * 1e-12
        out["roofline"] = {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                           "note": "22,989 unit-FLOP per on-ground aircraft-step x aircraft-steps/s; peak = FP64 FMA probe of this run"}
    sim.close()
    return out


def cfg5(fb, n, seconds):
    """BASELINE.json configs[4]: n lattice points sharded over the ranks (torchrun: one rank per GPU, no exchange); per rank a
    batched trim solve + forward-difference linearize() per point. Two legs: Jacobians left on the device as a LinearizedSS
    ensemble (fb_lss_from_linearize: what a downstream batched analysis consumes) and Jacobians gathered to the host."""
    import torch
    from flightbatch.dist import shard_range
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # regular lattice EAS x h x flaps x payload, NED kinematics (linearize's layout)
    m = int(round(n ** 0.25)) + 1
    EAS, h, fl, mass = np.meshgrid(np.linspace(25, 55, m), np.linspace(50, 3050, m), np.linspace(0, 1, m), np.linspace(0, 100, m), indexing="ij")
    EAS, h, fl, mass = (a.ravel()[:n] for a in (EAS, h, fl, mass))
    fl = np.where(EAS > 42, 0.0, fl)
    lo, hi = shard_range(n, rank, world)
    sl = slice(lo, hi)

    tok = torch.zeros(1, device="cuda")

    def barrier():   # an NCCL all-reduce of one element: dist.barrier() itself costs ~20 ms per call here, ten times the linearize leg
        torch.cuda.synchronize()
        if world > 1:
            dist.all_reduce(tok)
            torch.cuda.synchronize()

    def leg(gather):
        sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.NED())), n=hi - lo, dt=DT, t_end=10.0, device=local, member_offset=lo)
        barrier()
        t0 = time.perf_counter()
        ok, _, _ = sim.init(fb.C172.TrimParameters(EAS=EAS[sl], h_e=h[sl], flaps=fl[sl], m_pilot=mass[sl], m_copilot=mass[sl]))
        sim.sync(); t1 = time.perf_counter()
        if gather:
            out = sim.linearize(); shape = list(np.shape(out["A"]))
        else:
            out = fb.LinearizedSS(sim); out.sync(); shape = [hi - lo, out.nx, out.nx]
        sim.sync(); barrier(); t2 = time.perf_counter()
        tt = torch.tensor([t1 - t0, t2 - t1, t2 - t0], dtype=torch.float64, device="cuda")
        okc = torch.tensor([float(ok.sum())], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX); dist.all_reduce(okc, op=dist.ReduceOp.SUM)
        if not gather:
            out.close()
        sim.close()
        return [float(v) for v in tt.tolist()], int(okc.item()), shape
    leg(False)                                   # warm-up (context, module load)
    td, okd, shape = leg(False)
    tg, okg, _ = leg(True)
    res = {"config": "cfg5: envelope sweep, Cessna172Sv0/NED, batched trim + forward-difference A,B,C,D per lattice point, points sharded over the GPUs (no exchange)",
           "points": n, "n_gpus": world, "points_per_gpu": hi - lo, "trim_ok": okd,
           "device_resident": {"trim_s": td[0], "linearize_s": td[1], "points_per_s": n / td[2], "what": "Jacobians stay on the device as a LinearizedSS ensemble (fb_lss_from_linearize)"},
           "gathered": {"trim_s": tg[0], "linearize_s": tg[1], "points_per_s": n / tg[2], "what": "A, B, C, D, x0, u0, y0 copied to pageable host arrays (fb_linearize)", "d2h_bytes_per_gpu": int((hi - lo) * 8 * (16 * 16 + 16 * 4 + 33 * 16 + 33 * 4 + 16 + 4 + 33 + 16))},
           "A_shape_per_gpu": shape, "timing": "host wall clock, barrier + synchronize on both sides, max over ranks"}
    if world > 1:
        dist.destroy_process_group()
    return res if rank == 0 else None


def lss(fb, n, seconds):
    """LinearizedSS ensemble (README example 3 per member): linearize a trimmed C172S/NED ensemble on the device, step the
    N linear models with the warp-per-member RK4 kernel."""
    import torch
    rng = np.random.default_rng(SEED + 5)
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.NED())), n=n, dt=DT, t_end=10.0)
    ok, _, _ = sim.init(fb.C172.TrimParameters(EAS=rng.uniform(30, 55, n), h_e=rng.uniform(50, 3050, n)))
    m = fb.LinearizedSS(sim)
    u = m.u; u[2] += 0.1; m.u = u
    stream = torch.cuda.ExternalStream(m.stream_handle())
    m.step(1.0); m.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = int(round(seconds))
    e0.record(stream)
    for _ in range(k):
        m.step(1.0)
    e1.record(stream)
    m.sync()
    ms = e0.elapsed_time(e1)
    steps = k * 100
    out = {"config": "LinearizedSS ensemble: N 16-state linear models from linearize() of a trimmed Cessna172Sv0/NED ensemble, elevator step, RK4 dt = 0.01",
           "members": n, "rk4_steps": steps, "ms": ms, "model_steps_per_s": n * steps / (ms * 1e-3),
           "fma_tflops": n * steps / (ms * 1e-3) * 4 * (16 * 16 + 16) * 2e-12}
    m.close(); sim.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["cfg3", "cfg4", "cfg5", "lss"])
    ap.add_argument("--members", type=int, default=0)
    ap.add_argument("--seconds", type=float, default=10.0)
    args = ap.parse_args()
    import flightbatch as fb
    n = args.members or {"cfg3": 1 << 20, "cfg4": 1 << 18, "cfg5": 100000, "lss": 65536}[args.which]
    out = {"cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5, "lss": lss}[args.which](fb, n, args.seconds)
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
