#!/bin/bash
# round 2, GPU call AD: after the initialisation staging arena — GPU suite, cfg 5 line (three repetitions: its timing is host wall clock)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $O/r02ad_tests.log
for r in 1 2 3; do python tools/bench_configs.py cfg5 > $O/r02ad_cfg5_$r.json 2> $O/r02ad_cfg5_$r.err; python -c "
import json; d=json.loads(open('$O/r02ad_cfg5_$r.json').read().strip().splitlines()[-1]); print('cfg5', d['device_resident'], d['gathered']['points_per_s'])"; done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02ad_bench_n1.json 2> $O/r02ad_bench_n1.err; python -c "
import json; d=json.loads(open('$O/r02ad_bench_n1.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['e2e']['value'])"
