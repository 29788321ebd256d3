#!/bin/bash
# GPU visit for the tensor-core (add_tanh, bf16) kernels: parity tests, bench, ncu capture.
TAG=${1:-at}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_joint_at.py -m gpu -q --timeout 300 > $OUT/${TAG}_pytest_at.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_at.log
tail -5 $OUT/${TAG}_pytest_at.log
for m in "--cfg 2" "--cfg 3" "--cfg 4"; do
  n=$(echo $m | tr -d ' -'); timeout 600 python bench.py --mode add_tanh --gemm bf16 --steps 20 --warmup 3 --no-cpu-baseline $m > $OUT/${TAG}_bench_$n.json 2> $OUT/${TAG}_bench_$n.err; echo "bench $m exit $?"
done
B="python bench.py --mode add_tanh --gemm bf16 --cfg 2 --steps 2 --warmup 3 --no-cpu-baseline"
$B > $OUT/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"at_lse_tc|at_grad_tc" -s 4 -c 2 -f -o $OUT/${TAG}_prof $B > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -4
