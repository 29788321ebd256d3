#!/bin/bash
# Multi-GPU bench under torchrun (weak scaling by utterance): scripts/gpu_multi.sh <N> <tag>
N=${1:-2}; TAG=${2:-mg}; OUT=gpurun_out; mkdir -p $OUT
python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_n1.json 2> $OUT/${TAG}_n1.err; echo "n1 exit $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 100 --warmup 10 > $OUT/${TAG}_n$N.json 2> $OUT/${TAG}_n$N.err; echo "n$N exit $?"
tail -c 600 $OUT/${TAG}_n$N.json; tail -3 $OUT/${TAG}_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $OUT/${TAG}_ref_n$N.json 2> $OUT/${TAG}_ref_n$N.err; echo "ref exit $?"
