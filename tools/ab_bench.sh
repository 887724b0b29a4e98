#!/bin/bash
# A/B of kernel-variant builds (tools/build_variant.py): one short bench per library, the per-kernel ms of a step side by side.
# usage: tools/ab_bench.sh <batch> <steps> name=path ...   (name "product" = the in-tree library)
B=$1; S=$2; shift 2
for nv in "$@"; do
  n=${nv%%=*}; p=${nv#*=}
  if [ "$p" = "product" ]; then unset BP_B200_LIB; else export BP_B200_LIB=$p; fi
  python bench.py --batch $B --steps $S --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ab_$n.json 2> gpurun_out/ab_$n.err
  python - "$n" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open("gpurun_out/ab_%s.json" % n))
    print(n, "proofs/s %.1f" % d["value"], "ms/step %.1f" % d["ms_per_step"], json.dumps(d["kernel_ms_per_step"]))
except Exception as e:
    print(n, "FAILED", e, open("gpurun_out/ab_%s.err" % n).read()[-800:])
PY
done
