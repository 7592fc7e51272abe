#!/bin/bash
# round 2, GPU call AA: complete kernels at 256 threads / 255 registers — full GPU suite, cfg 4 line, full capture of the cfg-4 kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $O/r02aa_tests.log
python tools/bench_configs.py cfg4 > $O/r02aa_cfg4.json 2> $O/r02aa_cfg4.err; tail -c 600 $O/r02aa_cfg4.json
timeout 900 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 6 -c 1 -f -o $O/r02aa_c4 python tools/bench_configs.py cfg4 > $O/r02aa_ncu.log 2>&1
ncu -i $O/r02aa_c4.ncu-rep --page raw --csv > $O/r02aa_c4_raw.csv 2>/dev/null
ncu -i $O/r02aa_c4.ncu-rep --page source --csv --print-source sass > $O/r02aa_c4_sass.csv 2>/dev/null
rm -f $O/r02aa_c4.ncu-rep
ls -la $O/r02aa_c4_raw.csv $O/r02aa_c4_sass.csv
