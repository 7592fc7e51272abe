#!/bin/bash
# round 2, GPU call D: GPU suite again (ground spread with per-step noise, C driver), time-slicing diagnostics
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -s > $O/r02d_tests.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02d_tests.log
grep -E "passed|failed|FAILED|Error|cfg4|run: " $O/r02d_tests.log | tail -20
B="python bench.py --no-cpu-baseline --no-aux --no-e2e --warmup 3"
run() { local name=$1; shift
  "$@" > $O/r02d_$name.json 2> $O/r02d_$name.err
  python - "$O/r02d_$name.json" "$name" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "steps/s %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "rhs/step %.3f" % d["config"]["rhs_evals_per_member_step"], d["config"]["ms_per_step_min_max"])
except Exception as e:
    print(sys.argv[2], "FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
}
run one_wave_100 $B --steps 10 --members 37888
run one_wave_400 $B --steps 5 --members 37888 --rk-per-step 400
run two_wave_100 $B --steps 10 --members 75776
run s128k_100 $B --steps 10 --members 131072
run s128k_200 $B --steps 5 --members 131072 --rk-per-step 200
run s128k_400 $B --steps 5 --members 131072 --rk-per-step 400
run p1m_100 $B --steps 4
run p1m_50 $B --steps 6 --rk-per-step 50
