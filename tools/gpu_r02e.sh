#!/bin/bash
# round 2, GPU call E: smoke, the C driver, the 384-thread fly-by-wire shape for comparison, final ncu captures + launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests/test_gpu_abi_c.py tests/test_gpu_linearize.py tests/test_gpu_properties.py -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|cfg5|run:" | tail
B="python bench.py --no-cpu-baseline --no-aux --no-e2e --warmup 3"
run() { local name=$1; shift
  "$@" > $O/r02e_$name.json 2> $O/r02e_$name.err
  python - "$O/r02e_$name.json" "$name" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "steps/s %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "rhs/step %.3f" % d["config"]["rhs_evals_per_member_step"])
except Exception as e:
    print(sys.argv[2], "FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
}
run x384_1m env FLIGHTBATCH_LIB=$PWD/devlibs/x1_384.so $B --steps 4
run x384_128k env FLIGHTBATCH_LIB=$PWD/devlibs/x1_384.so $B --steps 10 --members 131072
run main_1m $B --steps 4
# the bench line and its launch list (same command, second time under ncu: the first exited 0 without it)
python bench.py --steps 5 --warmup 3 --save-peak $O/r02e_fp64_peak.json > $O/r02e_bench_n1.json 2> $O/r02e_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02e_bench_ref.json 2> $O/r02e_bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02e_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r02e_ncu_l.log 2>&1
# full capture of the fly-by-wire integrator: 4 waves of the 8-warp shape (151,552 members), one timed launch
timeout 900 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 2 -c 1 -f -o $O/r02e_ncu \
    $B --steps 1 --warmup 1 --members 151552 > $O/r02e_ncu.log 2>&1
ncu -i $O/r02e_ncu.ncu-rep --page raw --csv > $O/r02e_ncu_raw.csv 2>/dev/null
ncu -i $O/r02e_ncu.ncu-rep --page source --csv --print-source sass > $O/r02e_ncu_sass.csv 2>/dev/null
rm -f $O/r02e_ncu.ncu-rep
# cfg 4 and cfg 5 on one GPU with the new build
python tools/bench_configs.py cfg4 > $O/r02e_cfg4.json 2> $O/r02e_cfg4.err; tail -c 600 $O/r02e_cfg4.json
python tools/bench_configs.py cfg5 > $O/r02e_cfg5.json 2> $O/r02e_cfg5.err; tail -c 900 $O/r02e_cfg5.json
