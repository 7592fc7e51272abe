#!/bin/bash
# round 2, GPU call Q: first-order strut geodesy (ground kernel), yf kept in registers on the common path
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "AssertionError|passed|failed|FAILED|Error|cfg4|ground" | cut -c1-300 | tail -14
B="python bench.py --no-cpu-baseline --no-aux --no-e2e --warmup 3"
one() { # lib dyn members extra...
  local lib=$1 dyn=$2 m=$3; shift 3
  FLIGHTBATCH_LIB=$PWD/$lib timeout 300 $B --steps 8 --members $m "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', '$m', '$*', 'steps/s %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'alive', d['config'].get('members_alive'))"
}
for rep in 1 2; do
one devlibs/x_cg.so 0 1048576
one devlibs/e1.so 0 1048576
one devlibs/x_cg.so 0 131072
one devlibs/e1.so 0 131072
done
python tools/bench_configs.py cfg4 2>/dev/null | tail -c 420
