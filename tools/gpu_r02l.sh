#!/bin/bash
# round 2, GPU call L: payload moments + full input cache + anchors (incremental lat/lon, ISA pressure) + split tap
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -s 2>&1 | grep -E "AssertionError|passed|failed|FAILED|worst|Error" | cut -c1-300 | tail -34
B="python bench.py --no-cpu-baseline --no-aux --no-e2e --warmup 3"
one() { # lib members extra...
  local lib=$1 m=$2; shift 2
  FLIGHTBATCH_LIB=$PWD/$lib $B --steps 6 --members $m "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', '$m', '$*', 'steps/s %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'rhs/step', d['config'].get('rhs_evals_per_member_step'))"
}
one devlibs/b0.so 1048576
one devlibs/b1.so 1048576
one devlibs/b0.so 131072
one devlibs/b1.so 131072
one flight.jl_b200/libflightbatch.so 1048576
one flight.jl_b200/libflightbatch.so 131072
one flight.jl_b200/libflightbatch.so 65536 --workload cfg2
one flight.jl_b200/libflightbatch.so 1048576 --workload cfg2
python tools/bench_configs.py cfg4 2>/dev/null | tail -c 400
