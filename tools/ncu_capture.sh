#!/bin/bash
# ncu --set full capture of one kernel of a batch-2731 bench step (one chunk of the default 8192-proof step); raw + source pages as CSV.
# usage: tools/ncu_capture.sh <KernelName> <launches> [extra bench args]
set -u
K=$1; C=$2; shift 2
OUT=gpurun_out/ncu_$K
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$K -c $C -o $OUT -f \
    python bench.py --batch 2731 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e "$@" > $OUT.log 2>&1
ncu -i $OUT.ncu-rep --page raw --csv > $OUT.raw.csv 2>>$OUT.log
ncu -i $OUT.ncu-rep --page source --csv > $OUT.source.csv 2>>$OUT.log
ls -la $OUT.ncu-rep >> $OUT.log
# keep the report only if it is small enough to travel (gpurun brings back <= 64 MiB in total)
if [ $(stat -c %s $OUT.ncu-rep) -gt 12000000 ]; then rm -f $OUT.ncu-rep; fi
tail -3 $OUT.log
