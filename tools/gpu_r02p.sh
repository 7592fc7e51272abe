#!/bin/bash
# round 2, GPU call P: tests incl. the dynamic-queue bit-identity test; 10- and 12-warp shapes of the leaner fly-by-wire kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "AssertionError|passed|failed|FAILED|Error" | cut -c1-300 | tail -12
B="python bench.py --no-cpu-baseline --no-aux --no-e2e --warmup 3"
one() { # lib dyn members extra...
  local lib=$1 dyn=$2 m=$3; shift 3
  FLIGHTBATCH_LIB=$PWD/$lib FLIGHTBATCH_DYNAMIC=$dyn timeout 300 $B --steps 6 --members $m "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib dyn=$dyn', '$m', '$*', 'steps/s %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'alive', d['config'].get('members_alive'))"
}
one devlibs/x_cg.so 0 1048576
one devlibs/x_t320.so 0 1048576
one devlibs/x_t384.so 0 1048576
one devlibs/x_t320.so 1 131072
one devlibs/x_t384.so 1 131072
python bench.py --steps 10 --warmup 3 --members 131072 2>/dev/null | tail -1 | cut -c1-1500
