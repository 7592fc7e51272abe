#!/bin/bash
# round 2, GPU call O: A/B of the L2-only load policy (-dlcm=cg) and the dynamic task queue, same box, interleaved repeats
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
B="python bench.py --no-cpu-baseline --no-aux --no-e2e --warmup 3"
one() { # lib dyn members extra...
  local lib=$1 dyn=$2 m=$3; shift 3
  FLIGHTBATCH_LIB=$PWD/devlibs/$lib.so FLIGHTBATCH_DYNAMIC=$dyn timeout 300 $B --steps 8 --members $m "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib dyn=$dyn', '$m', '$*', 'steps/s %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'alive', d['config'].get('members_alive'))"
}
for rep in 1 2; do
for lib in s_ca s_cg; do for dyn in 0 1; do
one $lib $dyn 65536 --workload cfg2
one $lib $dyn 1048576 --workload cfg2
done; done
for lib in x_ca x_cg; do for dyn in 0 1; do
one $lib $dyn 131072
one $lib $dyn 1048576
done; done
done
