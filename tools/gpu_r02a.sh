#!/bin/bash
# round 2, GPU call A: sanity of the existing GPU tests, the new cfg-3 bench line, launch-shape variants, ncu captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r02a_tests.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02a_tests.log
tail -3 $O/r02a_tests.log
B="python bench.py --no-cpu-baseline --no-aux --no-e2e --warmup 3"
run() { # name lib args...
  local name=$1 lib=$2; shift 2
  FLIGHTBATCH_LIB=$PWD/$lib $B "$@" > $O/r02a_$name.json 2> $O/r02a_$name.err
  python - "$O/r02a_$name.json" "$name" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "steps/s %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "rhs/step %.3f" % d["config"]["rhs_evals_per_member_step"], d["clocks"]["sm_mhz"] if d["clocks"] else None)
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
MAIN=flight.jl_b200/libflightbatch.so
run c3_main_1m   $MAIN --steps 5
run c3_main_128k $MAIN --steps 10 --members 131072
run c3_main_256k $MAIN --steps 10 --members 262144
run c3_t256_1m   devlibs/x1_t256.so --steps 5
run c3_t256_128k devlibs/x1_t256.so --steps 10 --members 131072
run c3_t256_256k devlibs/x1_t256.so --steps 10 --members 262144
run c3_w4_1m     devlibs/x1_w4.so --steps 5
run c2_main_64k  $MAIN --workload cfg2 --steps 10
run c2_main_1m   $MAIN --workload cfg2 --steps 5 --members 1048576
run c2_t256_64k  devlibs/s_t256.so --workload cfg2 --steps 10
run c2_t256_1m   devlibs/s_t256.so --workload cfg2 --steps 5 --members 1048576
# the full default line once (e2e, aux, cpu baseline)
python bench.py --steps 10 --warmup 3 --save-peak $O/r02a_fp64_peak.json > $O/r02a_bench_n1.json 2> $O/r02a_bench_n1.err; echo "bench rc=$?"
tail -c 600 $O/r02a_bench_n1.json
# ncu: full capture of the fly-by-wire integrator, both shapes (one timed launch each, after one warm-up launch pair)
for v in main:$MAIN t256:devlibs/x1_t256.so; do
  name=${v%%:*}; lib=${v#*:}
  FLIGHTBATCH_LIB=$PWD/$lib timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 2 -c 1 -f -o $O/r02a_ncu_$name \
    $B --steps 1 --warmup 1 --members 227328 > $O/r02a_ncu_$name.log 2>&1
  ncu -i $O/r02a_ncu_$name.ncu-rep --page raw --csv > $O/r02a_ncu_${name}_raw.csv 2>/dev/null
  ncu -i $O/r02a_ncu_$name.ncu-rep --page source --csv --print-source sass > $O/r02a_ncu_${name}_sass.csv 2>/dev/null
  rm -f $O/r02a_ncu_$name.ncu-rep
done
ls -la $O | tail -30
