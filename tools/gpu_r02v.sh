#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
grep -m1 "model name" /proc/cpuinfo; nproc
for lib in devlibs/g_exact.so devlibs/g_mid.so; do echo "== $lib"; FLIGHTBATCH_LIB=$PWD/$lib python -m pytest tests/test_gpu_ground.py -m gpu -q -s -k cfg4 2>&1 | grep -E "AssertionError|passed|failed|cfg4 GPU" | cut -c1-300; done
