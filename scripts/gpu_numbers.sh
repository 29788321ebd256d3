#!/bin/bash
# The table of profiles/r2_results_table.md: every BASELINE config, both joints, full and ragged, the batch
# sweep, the AMP variant's step, the full training step.  usage: scripts/gpu_numbers.sh <tag>
TAG=${1:-num}; OUT=gpurun_out; mkdir -p $OUT
run() { n=$1; shift; timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-gpu-baseline "$@" > $OUT/${TAG}_$n.json 2> $OUT/${TAG}_$n.err; echo "$n exit $?"; }
run cg_cfg1 --cfg 1
run cg_cfg2 --cfg 2
run cg_cfg2_ragged --cfg 2 --ragged
run cg_cfg2_det --cfg 2 --deterministic
run cg_cfg3 --cfg 3
run cg_cfg4 --cfg 4
run cg_cfg2_b128 --cfg 2 --batch 128
run cg_cfg2_b512 --cfg 2 --batch 512
run at_cfg2 --cfg 2 --mode add_tanh --gemm bf16
run at_cfg3 --cfg 3 --mode add_tanh --gemm bf16
run at_cfg4 --cfg 4 --mode add_tanh --gemm bf16 --steps 5 --warmup 3
run cfg5 --cfg 5 --steps 8 --warmup 3
for c in 3 4; do timeout 900 python bench.py --steps 20 --warmup 5 --cfg $c --no-cpu-baseline > $OUT/${TAG}_gpubar_cfg$c.json 2> $OUT/${TAG}_gpubar_cfg$c.err; echo "gpubar cfg$c exit $?"; done
python bench.py --steps 100 --warmup 10 > $OUT/${TAG}_headline.json 2> $OUT/${TAG}_headline.err; echo "headline exit $?"
python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_reference.json 2> $OUT/${TAG}_reference.err; echo "reference exit $?"
