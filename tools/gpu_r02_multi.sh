#!/bin/bash
# round 2, multi-GPU call: usage tools/gpu_r02_multi.sh "<N list>"   (run under gpurun --gpus <max N>)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
NS=${1:-"2"}
nvidia-smi -L | head -8
python -m pytest tests/test_gpu_stats.py -m gpu -q -k "group" -s 2>&1 | grep -E "passed|failed|skipped|FAILED|Error" | tail -5
for N in $NS; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N))"
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/r02_bench_n$N.json 2> $O/r02_bench_n$N.err; echo "bench N=$N rc=$?"
  grep -h "^{" $O/r02_bench_n$N.json > $O/r02_bench_n$N.line; mv $O/r02_bench_n$N.line $O/r02_bench_n$N.json
  grep -E "AllReduce|NVLS|comm 0x.* rank 0 nranks" $O/r02_bench_n$N.json $O/r02_bench_n$N.err | head -6 > $O/r02_comm_n$N.log
  python - "$O/r02_bench_n$N.json" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("N=%d" % d["n_gpus"], "steps/s %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e %.4g" % d["e2e"]["value"], "allreduce ms/step", d.get("stats", {}).get("allreduce_ms_per_step"), d["stats"]["last_sample"]["eas"])
except Exception as e:
    print("FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-2500:])
PY
  $TR tools/bench_configs.py cfg5 > $O/r02_cfg5_n$N.json 2> $O/r02_cfg5_n$N.err; echo "cfg5 N=$N rc=$?"; grep -h "^{" $O/r02_cfg5_n$N.json | cut -c1-700
done
cat $O/r02_comm_n*.log | head -12
