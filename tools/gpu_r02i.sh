#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for lib in flight.jl_b200/libflightbatch.so devlibs/full_exact.so; do
  echo "== $lib"
  FLIGHTBATCH_LIB=$PWD/$lib python -m pytest tests -m gpu -q -s 2>&1 | grep -E "AssertionError|passed|failed|FAILED|worst|Error" | cut -c1-300 | tail -30
done
