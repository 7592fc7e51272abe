#!/bin/bash
# usage: tools/gpu_ab.sh "<lib a> <lib b> ..."  — interleaved cfg-3 A/B of development builds on one box (2^20 and 131,072 members)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
B="python bench.py --no-cpu-baseline --no-aux --no-e2e --warmup 3"
one() { local lib=$1 m=$2; shift 2
  FLIGHTBATCH_LIB=$PWD/$lib timeout 300 $B --steps 8 --members $m "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', '$m', '$*', 'steps/s %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'])"
}
for rep in 1 2; do for lib in $1; do one $lib 1048576; one $lib 131072; done; done
