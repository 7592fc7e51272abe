#!/bin/bash
# round 2, GPU call B: all GPU tests on the new build (8-warp fly-by-wire shape, shared-memory tap, in-kernel moments), cfg-3 bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/r02b_tests.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02b_tests.log
tail -15 $O/r02b_tests.log
B="python bench.py --no-cpu-baseline --no-aux --no-e2e --warmup 3"
run() { local name=$1; shift
  $B "$@" > $O/r02b_$name.json 2> $O/r02b_$name.err
  python - "$O/r02b_$name.json" "$name" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "steps/s %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "rhs/step %.3f" % d["config"]["rhs_evals_per_member_step"], d["clocks"]["sm_mhz"] if d["clocks"] else None, d.get("stats", {}).get("last_sample", {}).get("eas"))
except Exception as e:
    print(sys.argv[2], "FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
}
run c3_1m --steps 5
run c3_128k --steps 10 --members 131072
run c3_256k --steps 10 --members 262144
run c3_512k --steps 10 --members 524288
run c2_64k --workload cfg2 --steps 10
python bench.py --steps 10 --warmup 3 > $O/r02b_bench_n1.json 2> $O/r02b_bench_n1.err; echo "bench rc=$?"; tail -c 1500 $O/r02b_bench_n1.json; tail -5 $O/r02b_bench_n1.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 2 -c 1 -f -o $O/r02b_ncu \
    $B --steps 1 --warmup 1 --members 227328 > $O/r02b_ncu.log 2>&1
ncu -i $O/r02b_ncu.ncu-rep --page raw --csv > $O/r02b_ncu_raw.csv 2>/dev/null
ncu -i $O/r02b_ncu.ncu-rep --page source --csv --print-source sass > $O/r02b_ncu_sass.csv 2>/dev/null
rm -f $O/r02b_ncu.ncu-rep
