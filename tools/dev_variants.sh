#!/bin/bash
# usage: tools/dev_variants.sh [members] [steps]  — benches every development build in devlibs/*.so (FLIGHTBATCH_LIB override)
for so in devlibs/*.so; do
  FLIGHTBATCH_LIB=$PWD/$so python bench.py --steps "${2:-8}" --warmup 3 --no-cpu-baseline --members "${1:-65536}" 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', 'steps/s %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'rhs/step', d['config'].get('rhs_evals_per_member_step'), d['clocks']['sm_mhz'])"
done
