#!/bin/bash
# Eight-GPU visit (gpurun --gpus 8): weak scaling of the joint+loss step at N = 1, 2, 4, 8 with the peer-memory
# all-reduce inside the step's CUDA graph, the NCCL variant beside it at N = 8, and the full training step
# (cfg 5, torch DDP) at N = 1 and 8.
TAG=${1:-mg8}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
tr() { n=$1; g=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $g --steps 50 --warmup 10 --no-cpu-baseline --no-gpu-baseline "$@" > $OUT/${TAG}_$n.json 2> $OUT/${TAG}_$n.err; echo "$n exit $?"; tail -n 2 $OUT/${TAG}_$n.err | cut -c1-300; }
timeout 300 python -m pytest tests/test_gpu_comm.py -m gpu -q --timeout 200 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest comm exit $?"; tail -n 3 $OUT/${TAG}_pytest.log
python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_n1.json 2> $OUT/${TAG}_n1.err; echo "n1 exit $?"
tr n8_peer 8
tr n8_nccl 8 --allreduce nccl
tr n8_peer_rep 8
tr n4_peer 4
tr n2_peer 2
python bench.py --cfg 5 --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_cfg5_n1.json 2> $OUT/${TAG}_cfg5_n1.err; echo "cfg5 n1 exit $?"
tr cfg5_n8 8 --cfg 5 --steps 8 --warmup 3
python - << PY
import json, glob
base = {}
for f in sorted(glob.glob("$OUT/${TAG}_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("/")[-1], "N", d["n_gpus"], "ms", round(d["ms_per_step"], 4), "Gc/s", round(d["value"] / 1e9, 4), "e2e", round(d["e2e"]["value"] / 1e9, 3), d.get("allreduce"))
PY
