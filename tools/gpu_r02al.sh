#!/bin/bash
# round 2, GPU call AL: final check of the build (smoke, whole GPU suite) and the complete-only mode on cfg 4 (what the 1/8 rule would pick)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $O/r02al_tests.log
for mode in pair complete; do
  if [ $mode = complete ]; then export FLIGHTBATCH_KERNEL=complete; fi
  timeout 200 python tools/bench_configs.py cfg4 > $O/r02al_cfg4_$mode.json 2> $O/r02al_cfg4_$mode.err
  python -c "
import json; d=json.loads(open('$O/r02al_cfg4_$mode.json').read().strip().splitlines()[-1]); print('cfg4 $mode %.4g' % d['aircraft_steps_per_s'], d['kernel_launches'])"
done
