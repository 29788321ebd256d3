#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list + one full capture of the named kernel.
# usage: scripts/gpu_check.sh <tag> [kernel-regex] [extra bench args]
TAG=${1:-r1}; KREGEX=${2:-lattice_sweep}; shift 2 || true
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.csv 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --ignore=tests/test_gpu_joint_at.py > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -12 $OUT/${TAG}_pytest.log
timeout 600 python -m pytest tests/test_gpu_joint_at.py -m gpu -q --timeout 300 > $OUT/${TAG}_pytest_at.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_at.log
tail -12 $OUT/${TAG}_pytest_at.log
python bench.py --steps 100 --warmup 10 "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
tail -c 2500 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
python bench.py --steps 50 --warmup 5 --eager --no-cpu-baseline "$@" > $OUT/${TAG}_bench_eager.json 2> $OUT/${TAG}_bench_eager.err
for m in "--mode add_tanh --gemm bf16 --cfg 2" "--mode add_tanh --gemm bf16 --cfg 4" "--mode add_tanh --gemm fp32 --cfg 2"; do
  n=$(echo $m | tr -d ' -'); timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $m > $OUT/${TAG}_bench_$n.json 2> $OUT/${TAG}_bench_$n.err; echo "bench $m exit $?"
done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > $OUT/${TAG}_ncu_launches.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 9 -c 6 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -5
