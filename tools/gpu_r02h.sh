#!/bin/bash
# round 2, GPU call H: first-order CoM geodesy (no second ECEF -> geodetic conversion per right-hand side)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
B="python bench.py --no-cpu-baseline --no-aux --no-e2e --warmup 3"
run() { local name=$1; shift
  "$@" > $O/r02h_$name.json 2> $O/r02h_$name.err
  python - "$O/r02h_$name.json" "$name" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "steps/s %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "rhs/step %.3f" % d["config"]["rhs_evals_per_member_step"], "e2e", d["e2e"]["value"] if d.get("e2e") else None)
except Exception as e:
    print(sys.argv[2], "FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
}
run s128k $B --steps 10 --members 131072
run full $B --steps 5
run c2 $B --workload cfg2 --steps 10
