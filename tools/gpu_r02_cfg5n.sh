#!/bin/bash
# usage (under gpurun --gpus N): tools/gpu_r02_cfg5n.sh N — cfg 5 sharded over N GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}; O=gpurun_out; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) tools/bench_configs.py cfg5 > $O/r02_cfg5_n$N.json 2> $O/r02_cfg5_n$N.err; echo "cfg5 N=$N rc=$?"; grep -h "^{" $O/r02_cfg5_n$N.json | cut -c1-800
