#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for lib in g_exact g_mid; do echo "== $lib"; FLIGHTBATCH_LIB=$PWD/devlibs/$lib.so python tools/diag_ground2.py 2>&1 | tail -14; done
