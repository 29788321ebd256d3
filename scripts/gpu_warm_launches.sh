#!/bin/bash
# Per-kernel durations inside the step with the caches as the previous kernel left them
# (ncu --cache-control none): what each kernel costs IN the chain, not alone with a cold L2.
# usage: scripts/gpu_warm_launches.sh <tag> [extra bench args]
TAG=${1:-warm}; shift || true
OUT=gpurun_out; mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --eager "$@" \
    > $OUT/${TAG}_ncu.log 2>&1
python - << PY
import csv
rows = [r for r in csv.reader(open("$OUT/${TAG}_launches.csv")) if len(r) > 10]
hdr = rows[0]
kn, val = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(r[kn][:60], float(r[val].replace(",", ""))) for r in rows[1:]]
# last full step: print the last 14 launches
for k, v in seq[-16:]:
    print(f"{v/1000:8.2f} us  {k}")
PY
