#!/bin/bash
# Round-2 visit G: the loss reduction folded into the kernels -- tests, cfg-2 step (fp32 and fp16 activations), launch list.
TAG=${1:-r2g}; OUT=gpurun_out; mkdir -p $OUT
T="tests/test_gpu_joint_cg.py tests/test_gpu_loss.py tests/test_gpu_amp.py tests/test_gpu_joint_at.py tests/test_gpu_ddp.py tests/test_gpu_decode.py"
timeout 900 python -m pytest $T -m gpu -q --timeout 600 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -n 8 $OUT/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 $OUT/${TAG}_smoke.log
run() { n=$1; shift; timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-gpu-baseline "$@" > $OUT/${TAG}_$n.json 2> $OUT/${TAG}_$n.err; echo "$n exit $?"; tail -n 2 $OUT/${TAG}_$n.err; }
run cfg2 --cfg 2
run cfg2_rep --cfg 2
run cfg2_fp16 --cfg 2 --act-dtype fp16
run cfg2_bf16 --cfg 2 --act-dtype bf16
run cfg3 --cfg 3
run cfg1 --cfg 1
python - << PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("/")[-1], "ms", round(d["ms_per_step"], 4), "Gc/s", round(d["value"] / 1e9, 3), "e2e", round(d["e2e"]["value"] / 1e9, 3), "h2d", d["e2e"]["h2d_bytes_per_step"],
          {k: round(v["us"], 1) for k, v in d.get("kernels", {}).items()}, d.get("step_ms_p10_p50_p90_max"), d["loss"])
PY
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
