#!/bin/bash
# round 2, GPU call C: the whole GPU suite (new stats / long-horizon / branch / ground-spread tests), strong-scaling shard diagnostics
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -s > $O/r02c_tests.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02c_tests.log
grep -E "passed|failed|FAILED|Error|worst|cfg4|cfg2,|cfg3," $O/r02c_tests.log | tail -40
B="python bench.py --no-cpu-baseline --no-aux --no-e2e --warmup 3"
run() { local name=$1; shift
  "$@" > $O/r02c_$name.json 2> $O/r02c_$name.err
  python - "$O/r02c_$name.json" "$name" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "steps/s %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "rhs/step %.3f" % d["config"]["rhs_evals_per_member_step"], {k: v["count"] for k, v in d.get("stats", {}).get("last_sample", {}).items()})
except Exception as e:
    print(sys.argv[2], "FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
}
run c3_1m $B --steps 5
run c3_128k $B --steps 10 --members 131072
run c3_128k_plain env FLIGHTBATCH_SLICED=0 $B --steps 10 --members 131072
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02c_launches_128k.csv $B --steps 3 --warmup 3 --members 131072 > $O/r02c_ncu_l.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/r02c_launches_128k.csv") if not l.startswith("==")))
h = rows[0]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
for r in rows[1:]:
    if len(r) > iv: print(r[ik][:60], r[iv])
PY
