#!/bin/bash
# usage: tools/bench_variant.sh <lib.so> <members> [steps]   (swaps the engine library, runs a short bench, prints key numbers)
[ "$1" != "flight.jl_b200/libflightbatch.so" ] && cp "$1" flight.jl_b200/libflightbatch.so
python bench.py --steps "${3:-8}" --warmup 3 --no-cpu-baseline --members "$2" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', '$2', 'steps/s %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], d['clocks'])"
