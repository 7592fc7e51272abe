#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
B="python bench.py --no-cpu-baseline --no-aux --no-e2e --warmup 3"
one() { local lib=$1 m=$2; shift 2
  FLIGHTBATCH_LIB=$PWD/$lib timeout 300 $B --steps 8 --members $m "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', '$m', '$*', 'steps/s %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'])"
}
for rep in 1 2; do
one devlibs/x_cg.so 1048576
one devlibs/p3.so 1048576
one devlibs/x_cg.so 131072
one devlibs/p3.so 131072
done
for lib in g_exact g_lin g_exact g_lin; do
echo "== $lib"; FLIGHTBATCH_LIB=$PWD/devlibs/$lib.so python tools/bench_configs.py cfg4 2>/dev/null | tail -c 230
done
