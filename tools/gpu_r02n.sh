#!/bin/bash
# round 2, GPU call N: airborne launches on a global (member-warp, chunk) task queue, one persistent CTA per SM
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "AssertionError|passed|failed|FAILED|Error" | cut -c1-300 | tail -12
B="python bench.py --no-cpu-baseline --no-aux --no-e2e --warmup 3"
one() { # env members extra...
  local ev=$1 m=$2; shift 2
  env $ev timeout 300 $B --steps 6 --members $m "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$ev', '$m', '$*', 'steps/s %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'rhs/step', d['config'].get('rhs_evals_per_member_step'))"
}
one A=1 131072
one FLIGHTBATCH_DYNAMIC=0 131072
one A=1 262144
one A=1 524288
one A=1 1048576
one FLIGHTBATCH_DYNAMIC=0 1048576
one A=1 65536 --workload cfg2
one FLIGHTBATCH_DYNAMIC=0 65536 --workload cfg2
one A=1 1048576 --workload cfg2
