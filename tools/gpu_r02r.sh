#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for lib in g_exact g_lin g_exact g_lin; do
echo "== $lib"; FLIGHTBATCH_LIB=$PWD/devlibs/$lib.so python tools/bench_configs.py cfg4 2>/dev/null | tail -c 300
done
