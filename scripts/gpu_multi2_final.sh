#!/bin/bash
# Two-GPU check of the final code (gpurun --gpus 2): the two tests a one-GPU box skips, and the driver's own N=2 launch.
TAG=${1:-r2x2}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_ddp.py -m gpu -q --timeout 500 -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -n 4 $OUT/${TAG}_pytest.log
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_n1.json 2> $OUT/${TAG}_n1.err; echo "n1 exit $?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/${TAG}_n2.json 2> $OUT/${TAG}_n2.err; echo "n2 exit $?"
python - << PY
import json
for n in ("n1", "n2"):
    try:
        d = json.load(open("$OUT/${TAG}_%s.json" % n))
        print(n, "N", d["n_gpus"], "ms", round(d["ms_per_step"], 4), "Gc/s", round(d["value"] / 1e9, 3), "e2e", round(d["e2e"]["value"] / 1e9, 3))
    except Exception as e:
        print(n, "no line", e)
PY
