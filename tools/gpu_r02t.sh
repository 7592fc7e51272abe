#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "AssertionError|passed|failed|FAILED|Error|cfg4 GPU" | cut -c1-300 | tail -8
for lib in g_exact g_mid g_exact g_mid; do
echo "== $lib"; FLIGHTBATCH_LIB=$PWD/devlibs/$lib.so python tools/bench_configs.py cfg4 2>/dev/null | tail -c 230
done
B="python bench.py --no-cpu-baseline --no-aux --no-e2e --warmup 3"
for m in 131072 1048576; do $B --steps 8 --members $m 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$m', 'steps/s %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'])"; done
$B --steps 8 --workload cfg2 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg2', 'steps/s %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'])"
