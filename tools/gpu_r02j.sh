#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
FLIGHTBATCH_LIB=$PWD/flight.jl_b200/libflightbatch.so python tools/diag_libdiff.py run /tmp/a.npz
FLIGHTBATCH_LIB=$PWD/devlibs/full_exact.so python tools/diag_libdiff.py run /tmp/b.npz
python tools/diag_libdiff.py cmp /tmp/a.npz /tmp/b.npz
