#!/bin/bash
# Launch list of one cfg-4 step (V = 1024): which kernels are ours, which are library code (the projections for V > 80).
TAG=${1:-r2c4}; OUT=gpurun_out; mkdir -p $OUT
python bench.py --cfg 4 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --cfg 4 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
echo "exit $?"; wc -l $OUT/${TAG}_launches.csv
