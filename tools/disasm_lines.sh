#!/bin/bash
# usage: tools/disasm_lines.sh — nvdisasm -g -c listings (address -> source line) of the integrator translation units of the
# current build, for tools/ncu_by_line.py / ncu_phase.py / ncu_local_lines.py: build/dis_f/dis_main.txt, dis_complete.txt
set -e
D=/root/repo/build/dis_f
for tu in flightbatch flightbatch_complete; do
  W=$D/$tu; mkdir -p $W
  (cd $W && cuobjdump -xelf all /root/repo/build/$tu.o > /dev/null && nvdisasm -g -c $(ls $W | grep cubin | head -1) > $D/dis_$tu.txt 2>/dev/null)
done
ls -la $D
