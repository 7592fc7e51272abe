#!/bin/bash
# cfg-3 A/B of development builds (X1/WA only)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
B="python bench.py --no-cpu-baseline --no-aux --no-e2e --warmup 3 --steps 8"
one() { local lib=$1 m=$2
  FLIGHTBATCH_LIB=$PWD/devlibs/$lib.so timeout 200 $B --members $m 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', '$m', 'steps/s %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'])"
}
for rep in 1 2; do for l in ta tb; do one $l 1048576; done; done
for l in ta tb; do one $l 131072; done
