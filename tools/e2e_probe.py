"""development probe: per-launch time of the cfg-2 ensemble over 90 s, and how many members come near the terrain"""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench, flightbatch as fb
n=65536
p = bench.member_params(0, n)
sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=0.01, t_end=1e5, steps_per_launch=100)
sim.init(fb.C172.TrimParameters(**p))
ele_f, ail_f = fb.U["FB_U_ELEVATOR"], fb.U["FB_U_AILERON"]
ev = bench.doublet_events(n, sim.get_input(ele_f), sim.get_input(ail_f), ele_f, ail_f)
sim.schedule_inputs(*ev)
ms = []
for k in range(90):
    sim.sync(); t0=time.perf_counter(); sim.step(1.0); sim.sync(); ms.append((time.perf_counter()-t0)*1e3)
    if k % 10 == 9:
        h = sim.get_output("h_o"); st = sim.status()
        print(k+1, "s: ms", [round(m,2) for m in ms[-10:]], "min h_o", float(h[st==0].min()), "below 50 m:", int((h[st==0] < 50).sum()), "status hist", np.bincount(st, minlength=7).tolist())
