#!/bin/bash
# round 2, GPU call Z: cfg-4 A/B — complete kernel at 384 threads / 168 registers vs 256 / 255, static grid vs member-warp queue
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
one() { local tag=$1 lib=$2 dyn=$3
  FLIGHTBATCH_LIB=$PWD/$lib FLIGHTBATCH_COMPLETE_DYN=$dyn timeout 300 python tools/bench_configs.py cfg4 > $O/r02z_$tag.json 2> $O/r02z_$tag.err
  python - "$O/r02z_$tag.json" "$tag" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], "cfg4 %.4g" % d["aircraft_steps_per_s"], "ms %.2f" % d["ms"], d["status_hist"], d["rhs_per_member_step"])
except Exception as e:
    print(sys.argv[2], "FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1200:])
PY
}
for rep in 1 2; do
  one s384 flight.jl_b200/libflightbatch.so 0
  one d384 flight.jl_b200/libflightbatch.so 1
  one s256 devlibs/c256.so 0
  one d256 devlibs/c256.so 1
done
