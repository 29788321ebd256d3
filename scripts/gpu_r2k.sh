#!/bin/bash
# Round-2 visit K (2 GPUs): the whole GPU suite incl. the 2-GPU tests on the final code, N = 1 / 2 bench lines, headline + reference.
TAG=${1:-r2k}; OUT=gpurun_out; mkdir -p $OUT
rm -f $OUT/parity_r2.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=5 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -n 12 $OUT/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 $OUT/${TAG}_smoke.log
tr() { n=$1; g=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $g --steps 50 --warmup 10 --no-cpu-baseline --no-gpu-baseline "$@" > $OUT/${TAG}_$n.json 2> $OUT/${TAG}_$n.err; echo "$n exit $?"; tail -n 2 $OUT/${TAG}_$n.err | cut -c1-300; }
python bench.py --steps 100 --warmup 10 > $OUT/${TAG}_headline.json 2> $OUT/${TAG}_headline.err; echo "headline exit $?"
tr n2 2
tr n2_rep 2
tr n2_nccl 2 --allreduce nccl
python bench.py --steps 50 --warmup 5 --cfg 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_cfg3.json 2> $OUT/${TAG}_cfg3.err; echo "cfg3 exit $?"
python bench.py --steps 50 --warmup 5 --cfg 1 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_cfg1.json 2> $OUT/${TAG}_cfg1.err; echo "cfg1 exit $?"
python bench.py --steps 50 --warmup 5 --cfg 2 --act-dtype fp16 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_cfg2_fp16.json 2> $OUT/${TAG}_cfg2_fp16.err; echo "cfg2 fp16 exit $?"
python - << PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    a = d.get("allreduce") or {}
    print(f.split("/")[-1], "N", d["n_gpus"], "ms", round(d["ms_per_step"], 4), "Gc/s", round(d["value"] / 1e9, 3), "e2e", round(d["e2e"]["value"] / 1e9, 3),
          {k: round(v["us"], 1) for k, v in d.get("kernels", {}).items()}, a.get("us_alone_peer"), a.get("us_alone_nccl"))
PY
