#!/bin/bash
# Weak-scaling run on one box: N = 1, 2, 4, 8 back to back (scripts/gpu_scale.sh <tag>)
TAG=${1:-scale}; OUT=gpurun_out; mkdir -p $OUT
python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_n1.json 2> $OUT/${TAG}_n1.err; echo "n1 exit $?"
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) \
      bench.py --gpus $N --steps 100 --warmup 10 > $OUT/${TAG}_n$N.json 2> $OUT/${TAG}_n$N.err; echo "n$N exit $?"
done

