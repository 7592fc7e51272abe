#!/bin/bash
# round 2, GPU call F: final evidence of the round-2 build — tests, smoke, bench line, reference arm, launch list, full captures (2^20-member
# plain launch = the N=1 bench workload; 131,072-member task-queue launch = the 8-GPU shard), cfg 4 / cfg 5 / lss lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --save-peak $O/r02f_fp64_peak.json > $O/r02f_bench_n1.json 2> $O/r02f_bench_n1.err; echo "bench rc=$?"
python bench.py --steps 10 --warmup 3 --members 131072 --no-aux --no-cpu-baseline > $O/r02f_bench_128k.json 2> $O/r02f_bench_128k.err; echo "bench128k rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02f_bench_ref.json 2> $O/r02f_bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02f_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r02f_ncu_l.log 2>&1
B="python bench.py --no-cpu-baseline --no-aux --no-e2e --warmup 1 --steps 1"
for m in 1048576 131072; do
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 2 -c 1 -f -o $O/r02f_ncu_$m $B --members $m > $O/r02f_ncu_$m.log 2>&1
  ncu -i $O/r02f_ncu_$m.ncu-rep --page raw --csv > $O/r02f_ncu_${m}_raw.csv 2>/dev/null
  ncu -i $O/r02f_ncu_$m.ncu-rep --page source --csv --print-source sass > $O/r02f_ncu_${m}_sass.csv 2>/dev/null
  rm -f $O/r02f_ncu_$m.ncu-rep
done
python tools/bench_configs.py cfg4 > $O/r02f_cfg4.json 2> $O/r02f_cfg4.err; tail -c 700 $O/r02f_cfg4.json
python tools/bench_configs.py cfg5 > $O/r02f_cfg5.json 2> $O/r02f_cfg5.err; tail -c 500 $O/r02f_cfg5.json
python tools/bench_configs.py lss > $O/r02f_lss.json 2> $O/r02f_lss.err; tail -c 400 $O/r02f_lss.json
head -c 1800 $O/r02f_bench_n1.json; echo; head -c 600 $O/r02f_bench_128k.json; echo; cat $O/r02f_bench_ref.json | head -c 900
# the complete (ground) kernel inside the cfg-4 run: one full capture of a hand-over pass with half the ensemble on the runway
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:step_kernelILi0ELi0ELb0 --launch-skip 4 -c 1 -f -o $O/r02f_c4 python tools/bench_configs.py cfg4 > $O/r02f_c4_ncu.log 2>&1
ncu -i $O/r02f_c4.ncu-rep --page raw --csv > $O/r02f_c4_raw.csv 2>/dev/null
ncu -i $O/r02f_c4.ncu-rep --page source --csv --print-source sass > $O/r02f_c4_sass.csv 2>/dev/null
rm -f $O/r02f_c4.ncu-rep
