#!/usr/bin/env python3
"""Mixed ensemble: cfg 2's 65,536 Cessna172Sv0/WA members with every 64th member a few metres above raised terrain (they land
or crash during the run, i.e. they are handed over to the complete kernel), stepped through back-to-back 100-step launches.
Prints ms per launch with the hand-over pass pipelined on the side stream (default), with it on the launch stream behind each
airborne launch (FLIGHTBATCH_PIPELINE=0), and for the same ensemble without any low member. usage: bench_mixed.py [members]"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(n, every, launches=12):
    import torch
    import bench
    import flightbatch as fb
    p = bench.member_params(0, n)
    sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=0.01, t_end=1e4, steps_per_launch=100)
    if every:
        elev = np.zeros(n); low = np.arange(n) % every == 0
        elev[low] = p["h_e"][low] - bench.geoid_height(p["lat"][low], p["lon"][low]) - np.where(np.arange(n)[low] % (2 * every) == 0, 3.0, 5.0)
        sim.set_input("trn_elevation", elev)
    sim.init(fb.C172.TrimParameters(**p))
    sim.step(1.0); sim.sync()                     # warm-up launch (module load); the low members are handed over here already
    t0 = time.perf_counter()
    sim.step(float(launches))                     # `launches` back-to-back 100-step launches, no read in between
    sim.sync()
    dt = time.perf_counter() - t0
    st = sim.status()
    out = {"members": n, "low_members": int(n // every) if every else 0, "launches": launches, "ms_per_launch": 1e3 * dt / launches,
           "status_hist": np.bincount(st, minlength=7).tolist(), "pipeline": os.environ.get("FLIGHTBATCH_PIPELINE", "1")}
    sim.close()
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    if os.environ.get("FB_MIXED_CHILD"):
        print(json.dumps(run(n, int(os.environ["FB_MIXED_CHILD"]) if os.environ["FB_MIXED_CHILD"] != "0" else 0)))
    else:
        res = {}
        for tag, every, pipe in (("all_airborne", "0", "1"), ("mixed_pipelined", "64", "1"), ("mixed_serial", "64", "0")):
            env = dict(os.environ, FB_MIXED_CHILD=every, FLIGHTBATCH_PIPELINE=pipe)
            o = subprocess.run([sys.executable, __file__, str(n)], env=env, capture_output=True, text=True, timeout=240)
            res[tag] = json.loads(o.stdout.strip().splitlines()[-1]) if o.returncode == 0 else {"error": o.stderr[-500:]}
        print(json.dumps(res))
