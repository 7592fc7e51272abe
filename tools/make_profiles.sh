#!/bin/bash
# usage: tools/make_profiles.sh — turns the outputs of tools/gpu_r02final.sh (gpurun_out/r02f_*) into the tracked files under profiles/
cd /root/repo
tools/disasm_lines.sh > /dev/null
D=build/dis_f; K=step_kernelILi2ELi0ELb1ELi384; G=gpurun_out
{
echo "# Fly-by-wire integrator step_kernel<C172X1,WA,AIR> (8 warps x 255 registers), cfg 3 workload, final build of round 2; ncu --set full --clock-control none"
echo "# (a) 1,048,576 members: the N = 1 bench launch, plain one-thread-per-member shape (4096 CTAs x 256 threads = 27.7 waves), one 100-step launch"
python tools/ncu_summary.py $G/r02f_ncu_1048576_raw.csv
echo; echo "# phases (tools/ncu_phase.py)"
python tools/ncu_phase.py $G/r02f_ncu_1048576_sass.csv $D/dis_flightbatch.txt $K 1048576 100 12
echo; echo "# (b) 131,072 members: the shard of the 8-GPU split, global task queue (148 persistent CTAs, 8,192 (member-warp, chunk) tasks), one 100-step launch"
python tools/ncu_summary.py $G/r02f_ncu_131072_raw.csv
echo; echo "# phases (tools/ncu_phase.py)"
python tools/ncu_phase.py $G/r02f_ncu_131072_sass.csv $D/dis_flightbatch.txt $K 131072 100 12
echo; echo "# local-memory traffic by source line, capture (a) (tools/ncu_local_lines.py)"
python tools/ncu_local_lines.py $G/r02f_ncu_1048576_sass.csv $D/dis_flightbatch.txt $K 1048576 100 14
echo; echo "# (c) the complete (ground-capable) kernel step_kernel<C172S,WA,complete> inside the cfg-4 run (2^18 members, half of them on the runway): one hand-over pass (static grid, on the side stream beside the next airborne launch), 8 warps x 255 registers"
python tools/ncu_summary.py $G/r02f_c4_raw.csv
echo; python tools/ncu_opmix.py $G/r02f_c4_sass.csv 14
} > profiles/r02_step_kernel_ncu_full.txt 2>&1
python tools/ncu_by_line.py $G/r02f_ncu_1048576_sass.csv $D/dis_flightbatch.txt $K > profiles/r02_step_kernel_by_source_line.txt 2>&1
python tools/ncu_opmix.py $G/r02f_ncu_1048576_sass.csv > profiles/r02_step_kernel_opcode_mix.txt 2>&1
for f in bench_n1 bench_128k bench_ref cfg4 cfg5 lss; do cp $G/r02f_$f.json profiles/r02_$f.json; done
cp $G/r02f_launches.csv profiles/r02_launches.csv; cp $G/r02f_fp64_peak.json profiles/fp64_peak.json
grep -c "" profiles/r02_step_kernel_ncu_full.txt; grep -n "Traceback" profiles/r02_step_kernel_ncu_full.txt | head -3
