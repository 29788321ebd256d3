#!/bin/bash
# Eight-GPU visit D (final code): the scaling lines after taking NVML init out of the barrier-to-first-step window:
# N = 1, 2, 4, 8 as the driver launches them (--steps 20 --warmup 5), N = 8 again with 100 steps, NCCL beside it.
TAG=${1:-mg8c}; OUT=gpurun_out; mkdir -p $OUT
tr() { n=$1; g=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $g --no-cpu-baseline --no-gpu-baseline "$@" > $OUT/${TAG}_$n.json 2> $OUT/${TAG}_$n.err; echo "$n exit $?"; tail -n 2 $OUT/${TAG}_$n.err | cut -c1-300; }
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_n1.json 2> $OUT/${TAG}_n1.err; echo "n1 exit $?"
tr n2 2 --steps 20 --warmup 5
tr n4 4 --steps 20 --warmup 5
tr n8 8 --steps 20 --warmup 5
tr n8_rep 8 --steps 20 --warmup 5
tr n8_100 8 --steps 100 --warmup 10
python - << PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    a = d.get("allreduce") or {}
    b = a.get("breakdown") or {}
    print(f.split("/")[-1], "N", d["n_gpus"], "steps", d["steps"], "ms", round(d["ms_per_step"], 4), "Gc/s", round(d["value"] / 1e9, 4),
          "p10/50/90/max", [round(x, 4) for x in (b.get("step_ms_p10_p50_p90_max") or d.get("step_ms_p10_p50_p90_max") or [])],
          "ar", a.get("us_alone_peer"), a.get("us_alone_nccl"))
PY
