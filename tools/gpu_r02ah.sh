#!/bin/bash
# round 2, GPU call AH: pipelined hand-over pass — the whole GPU suite (short timeout), then the bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee $O/r02ah_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02ah_bench_n1.json 2> $O/r02ah_bench_n1.err; echo "bench rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --members 131072 --no-aux --no-cpu-baseline > $O/r02ah_bench_128k.json 2> $O/r02ah_bench_128k.err; echo "bench128k rc=$?"
for r in 1 2; do timeout 200 python tools/bench_configs.py cfg4 > $O/r02ah_cfg4.json 2> $O/r02ah_cfg4.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02ah_cfg4.json").read().strip().splitlines()[-1]); print("cfg4 %.4g" % d["aircraft_steps_per_s"], d["status_hist"], d["kernel_launches"])
except Exception as e:
    print("cfg4 FAILED", e); print(open("gpurun_out/r02ah_cfg4.err").read()[-1500:])
PY
done
python - <<'PY'
import json
for f in ("r02ah_bench_n1", "r02ah_bench_128k"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"], "aux", d.get("aux", {}).get("cfg2", {}).get("value"), "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
