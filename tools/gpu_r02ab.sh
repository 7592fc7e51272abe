#!/bin/bash
# round 2, GPU call AB: hand-over pass as persistent CTAs over a member-warp queue — cfg-4 A/B, ground / property tests, full capture
# of the complete kernel inside the cfg-4 run
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
one() { local tag=$1 dyn=$2
  FLIGHTBATCH_RESUME_DYN=$dyn timeout 300 python tools/bench_configs.py cfg4 > $O/r02ab_$tag.json 2> $O/r02ab_$tag.err
  python - "$O/r02ab_$tag.json" "$tag" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], "cfg4 %.4g" % d["aircraft_steps_per_s"], "ms %.2f" % d["ms"], d["status_hist"], d["rhs_per_member_step"])
except Exception as e:
    print(sys.argv[2], "FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1200:])
PY
}
for rep in 1 2; do one static 0; one queue 1; done
timeout 900 python -m pytest tests/test_gpu_ground.py tests/test_gpu_properties.py tests/test_gpu_branches.py tests/test_gpu_c172x.py -m gpu -q -x 2>&1 | tail -4 | tee $O/r02ab_tests.log
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:step_kernelILi0ELi0ELb0 --launch-skip 4 -c 1 -f -o $O/r02ab_c4 python tools/bench_configs.py cfg4 > $O/r02ab_ncu.log 2>&1
ncu -i $O/r02ab_c4.ncu-rep --page raw --csv > $O/r02ab_c4_raw.csv 2>/dev/null
ncu -i $O/r02ab_c4.ncu-rep --page source --csv --print-source sass > $O/r02ab_c4_sass.csv 2>/dev/null
rm -f $O/r02ab_c4.ncu-rep
ls -la $O/r02ab_c4_raw.csv $O/r02ab_c4_sass.csv
