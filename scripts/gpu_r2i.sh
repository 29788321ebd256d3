#!/bin/bash
# Round-2 visit I: programmatic dependent launch between the kernels of the step -- tests, A/B of the cfg-2 / cfg-3 / cfg-1 step.
TAG=${1:-r2i}; OUT=gpurun_out; mkdir -p $OUT
T="tests/test_gpu_joint_cg.py tests/test_gpu_loss.py tests/test_gpu_amp.py tests/test_gpu_joint_at.py tests/test_gpu_ddp.py tests/test_gpu_decode.py tests/test_gpu_fullsize.py"
timeout 1200 python -m pytest $T -m gpu -q --timeout 900 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -n 6 $OUT/${TAG}_pytest.log
run() { n=$1; shift; timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-gpu-baseline "$@" > $OUT/${TAG}_$n.json 2> $OUT/${TAG}_$n.err; echo "$n exit $?"; tail -n 2 $OUT/${TAG}_$n.err; }
run cfg2_pdl --cfg 2
RNNTB200_PDL=0 run cfg2_nopdl --cfg 2
run cfg2_pdl_rep --cfg 2
RNNTB200_PDL=0 run cfg2_nopdl_rep --cfg 2
run cfg3_pdl --cfg 3
RNNTB200_PDL=0 run cfg3_nopdl --cfg 3
run cfg1_pdl --cfg 1
RNNTB200_PDL=0 run cfg1_nopdl --cfg 1
run cfg2_ragged_pdl --cfg 2 --ragged
python - << PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("/")[-1], "ms", round(d["ms_per_step"], 4), "Gc/s", round(d["value"] / 1e9, 3), "e2e", round(d["e2e"]["value"] / 1e9, 3), d.get("step_ms_p10_p50_p90_max"), d["loss"])
PY
