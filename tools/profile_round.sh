#!/bin/bash
# Round profile on one B200: clean build on the box, smoke, ncu launch list of one bench step, ncu --set full of the heavy kernels.
set -u
rm -rf bulletproofs_r1cs_gadgets_b200/_obj bulletproofs_r1cs_gadgets_b200/libbp_b200.so oracle/_build
( time python -c "import __graft_entry__ as g; g.build(); g.smoke()" ) > gpurun_out/clean_build.log 2>&1
cat bulletproofs_r1cs_gadgets_b200/_obj/build_stamp.json >> gpurun_out/clean_build.log
tail -5 gpurun_out/clean_build.log
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/launches.csv \
    python bench.py --batch 2731 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-extras --no-verify > gpurun_out/launches.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv "ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --batch 2731 --steps 1 --warmup 0 (serialised, cold-cache launch times: compare SHARES)" > gpurun_out/launch_summary.csv
head -14 gpurun_out/launch_summary.csv
for k in KBucketAccumulate:2 KFoldTable:1 KFoldGens:2 KMsmAccumulate:2 KBucketReduce:2 sort_coarse_kernel:2; do
  ./tools/ncu_capture.sh ${k%%:*} ${k##*:} --no-verify
  python tools/ncu_summary.py gpurun_out/ncu_${k%%:*}.raw.csv > gpurun_out/ncu_summary_${k%%:*}.csv
done
ls -la gpurun_out | tail -30
