#!/bin/bash
# round 2, GPU call Y: the FB_C172XA2 variant (Actuator2 + VariablePitch) parity tests, then an A/B of the tuned kernels
# (cfg 3 / cfg 2 aux / cfg 4) after the host restructuring, full GPU test suite last
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_components.py -m gpu -x -q -s 2>&1 | tail -25 | tee $O/r02y_components.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02y_bench_n1.json 2> $O/r02y_bench_n1.err; echo "bench rc=$?"
python bench.py --steps 10 --warmup 3 --members 131072 --no-aux --no-cpu-baseline > $O/r02y_bench_128k.json 2> $O/r02y_bench_128k.err; echo "bench128k rc=$?"
python tools/bench_configs.py cfg4 > $O/r02y_cfg4.json 2> $O/r02y_cfg4.err
python - <<'PY'
import json
for f in ("r02y_bench_n1", "r02y_bench_128k"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"], "aux", d.get("aux", {}).get("cfg2", {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
try:
    d = json.loads(open("gpurun_out/r02y_cfg4.json").read().strip().splitlines()[-1]); print("cfg4 %.4g" % d["aircraft_steps_per_s"])
except Exception as e:
    print("cfg4 FAILED", e); print(open("gpurun_out/r02y_cfg4.err").read()[-1500:])
PY
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/r02y_tests.log
