#!/bin/bash
# A/B of whole-step time (CUDA-graph replay, L2 flushed between steps): legacy sweep vs warp-specialised sweep.
# usage: scripts/gpu_ab.sh <tag> [extra bench args]
TAG=${1:-ab}; shift || true
OUT=gpurun_out; mkdir -p $OUT
for v in ws legacy ws legacy; do
  if [ $v = legacy ]; then export RNNTB200_SWEEP_LEGACY=1; else unset RNNTB200_SWEEP_LEGACY; fi
  python bench.py --steps 200 --warmup 20 --no-cpu-baseline "$@" > $OUT/${TAG}_$v.json 2> $OUT/${TAG}_$v.err
  python - << PY
import json
d = json.load(open("$OUT/${TAG}_$v.json"))
print("$v", d["ms_per_step"], d["value"] / 1e9, {k: round(x["us"], 1) for k, x in d["kernels"].items()})
PY
done
