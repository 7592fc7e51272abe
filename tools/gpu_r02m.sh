#!/bin/bash
# round 2, GPU call M: tests after the clat refinement; ncu full capture of the time-sliced 131,072-member cfg-3 launch
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -s 2>&1 | grep -E "AssertionError|passed|failed|FAILED|Error" | cut -c1-300 | tail -12
B="python bench.py --no-cpu-baseline --no-aux --no-e2e --warmup 3"
$B --steps 6 --members 131072 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('128k', 'steps/s %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'])"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 2 -c 1 -f -o $O/r02m_ncu \
    $B --steps 1 --warmup 1 --members 131072 > $O/r02m_ncu.log 2>&1
ncu -i $O/r02m_ncu.ncu-rep --page raw --csv > $O/r02m_ncu_raw.csv 2>/dev/null
ncu -i $O/r02m_ncu.ncu-rep --page source --csv --print-source sass > $O/r02m_ncu_sass.csv 2>/dev/null
rm -f $O/r02m_ncu.ncu-rep
ls -la $O/r02m*
