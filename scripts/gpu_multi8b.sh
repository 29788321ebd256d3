#!/bin/bash
# Eight-GPU visit B: where does the N = 8 step go?  Per-step distribution, every rank's step without the
# collective, the collective alone -- peer-memory kernel and NCCL.
TAG=${1:-mg8b}; OUT=gpurun_out; mkdir -p $OUT
tr() { n=$1; g=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $g --steps 100 --warmup 10 --no-cpu-baseline --no-gpu-baseline "$@" > $OUT/${TAG}_$n.json 2> $OUT/${TAG}_$n.err; echo "$n exit $?"; tail -n 2 $OUT/${TAG}_$n.err | cut -c1-300; }
tr n8_peer 8
tr n8_nccl 8 --allreduce nccl
tr n4_peer 4
python - << PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("/")[-1], "N", d["n_gpus"], "ms", round(d["ms_per_step"], 4), "Gc/s", round(d["value"] / 1e9, 4), json.dumps(d.get("allreduce")))
PY
