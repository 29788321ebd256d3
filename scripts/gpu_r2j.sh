#!/bin/bash
# Round-2 visit J: PDL attribute WITHOUT the early trigger (the next kernel is pre-launched but only starts when its
# predecessor has exited) -- A/B on the cfg-2 / cfg-1 / cfg-3 step.
TAG=${1:-r2j}; OUT=gpurun_out; mkdir -p $OUT
run() { n=$1; shift; timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-gpu-baseline "$@" > $OUT/${TAG}_$n.json 2> $OUT/${TAG}_$n.err; echo "$n exit $?"; tail -n 2 $OUT/${TAG}_$n.err; }
RNNTB200_PDL=1 run cfg2_pdl --cfg 2
run cfg2_nopdl --cfg 2
RNNTB200_PDL=1 run cfg2_pdl_rep --cfg 2
run cfg2_nopdl_rep --cfg 2
RNNTB200_PDL=1 run cfg1_pdl --cfg 1
run cfg1_nopdl --cfg 1
RNNTB200_PDL=1 run cfg3_pdl --cfg 3
run cfg3_nopdl --cfg 3
RNNTB200_PDL=1 timeout 600 python -m pytest tests/test_gpu_joint_cg.py tests/test_gpu_loss.py -m gpu -q --timeout 600 > $OUT/${TAG}_pytest_pdl.log 2>&1; echo "pytest pdl exit $?"; tail -n 3 $OUT/${TAG}_pytest_pdl.log
python - << PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("/")[-1], "ms", round(d["ms_per_step"], 4), "Gc/s", round(d["value"] / 1e9, 3), d.get("step_ms_p10_p50_p90_max"))
PY
