#!/bin/bash
# Two-GPU visit (gpurun --gpus 2): DDP correctness over NCCL, the peer-memory all-reduce, scaling lines
# at N=1,2 for the joint+loss step (peer vs NCCL collective) and for the full training step (cfg 5).
TAG=${1:-mg}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_comm.py tests/test_gpu_ddp.py -m gpu -q --timeout 600 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -n 8 $OUT/${TAG}_pytest.log
tr() { n=$1; g=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $g --steps 50 --warmup 10 --no-cpu-baseline --no-gpu-baseline "$@" > $OUT/${TAG}_$n.json 2> $OUT/${TAG}_$n.err; echo "$n exit $?"; tail -n 2 $OUT/${TAG}_$n.err; }
python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_n1.json 2> $OUT/${TAG}_n1.err; echo "n1 exit $?"
tr n2_peer 2
tr n2_nccl 2 --allreduce nccl
tr n2_peer_rep 2
python bench.py --cfg 5 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_cfg5_n1.json 2> $OUT/${TAG}_cfg5_n1.err; echo "cfg5 n1 exit $?"
tr cfg5_n2 2 --cfg 5 --steps 10 --warmup 3
python - << PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("/")[-1], "N", d["n_gpus"], "ms", round(d["ms_per_step"], 4), "Gc/s", round(d["value"] / 1e9, 3), "e2e", round(d["e2e"]["value"] / 1e9, 3), d.get("allreduce"))
PY
