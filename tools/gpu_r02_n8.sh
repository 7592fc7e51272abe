#!/bin/bash
# round 2, 8-GPU call: the cfg-3 bench line under torchrun at N = 8 (and N = 4 if asked), NCCL log of the in-library all-reduce
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
NS=${1:-"8"}
nvidia-smi -L | head -8
for N in $NS; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N))"
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/r02_bench_n$N.out 2> $O/r02_bench_n$N.err; echo "bench N=$N rc=$?"
  grep -h "^{" $O/r02_bench_n$N.out > $O/r02_bench_n$N.json
  { grep -E "NCCL version|Init COMPLETE|NVLS|Connected all" $O/r02_bench_n$N.out | head -8; grep -E "AllReduce" $O/r02_bench_n$N.out | head -6; echo "AllReduce lines: $(grep -c AllReduce $O/r02_bench_n$N.out)"; } | cut -c1-260 > $O/r02_comm_n$N.log
  rm -f $O/r02_bench_n$N.out
  python - "$O/r02_bench_n$N.json" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("N=%d" % d["n_gpus"], "steps/s %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e %.4g" % d["e2e"]["value"], "allreduce latency ms", d.get("stats", {}).get("allreduce_latency_ms"), d["stats"]["last_sample"]["eas"])
except Exception as e:
    print("FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-2500:])
PY
done
cat $O/r02_comm_n*.log | head -12
