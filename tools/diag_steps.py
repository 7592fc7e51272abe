"""Per-bench-step device time over the 60 s cfg-2 job, with ensemble health indicators."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import flightbatch as fb
import bench
n = 65536
p = bench.member_params(0, n)
sim = fb.BatchSimulation(fb.SimpleWorld(fb.Cessna172Sv0(fb.WA())), n=n, dt=0.01, t_end=200.0, steps_per_launch=100)
ok, ts, cost = sim.init(fb.C172.TrimParameters(**p))
ele_f, ail_f = fb.U["FB_U_ELEVATOR"], fb.U["FB_U_AILERON"]
sim.schedule_inputs(*bench.doublet_events(n, sim.get_input(ele_f), sim.get_input(ail_f), ele_f, ail_f))
stream = torch.cuda.ExternalStream(sim.stream_handle())
c_prev = sim.counters()
for k in range(63):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); sim.step(1.0); e1.record(stream); sim.sync()
    c = sim.counters()
    if k % 3 == 0 or k < 12:
        st, en = sim.discrete()
        al = sim.get_output("alpha"); h = sim.get_output("h_o")
        print(f"t={k+1:3d}s ms={e0.elapsed_time(e1):7.3f} rhs/step={(c['rhs_evals']-c_prev['rhs_evals'])/n/100:.3f} alive={(sim.status()==0).sum()} stall={st.sum()} "
              f"alpha[min,max]=({al.min():.3f},{al.max():.3f}) h_o min={h.min():.1f} eng_off={(en!=2).sum()}")
    c_prev = c
