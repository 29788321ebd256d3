#!/bin/bash
# Round-2 visit B: the self-contained ("tp") sweep and the decoupled chain in the warp-specialised sweep --
# parity tests under both, then sweep timings per config (A/B through RNNTB200_SWEEP / RNNTB200_SWEEP_BW).
TAG=${1:-r2b}; OUT=gpurun_out; mkdir -p $OUT
T="tests/test_gpu_loss.py tests/test_gpu_joint_cg.py tests/test_abi.py"
timeout 900 python -m pytest $T -m gpu -q --timeout 600 -x > $OUT/${TAG}_pytest_tp.log 2>&1; echo "pytest tp exit $?"; tail -n 6 $OUT/${TAG}_pytest_tp.log
RNNTB200_SWEEP=ws timeout 900 python -m pytest $T -m gpu -q --timeout 600 -x > $OUT/${TAG}_pytest_ws.log 2>&1; echo "pytest ws exit $?"; tail -n 6 $OUT/${TAG}_pytest_ws.log
for bw in 1 3 4; do
RNNTB200_SWEEP_BW=$bw timeout 600 python -m pytest tests/test_gpu_loss.py -m gpu -q --timeout 600 -x -k "boundaries or long_lattice" > $OUT/${TAG}_pytest_tp_bw$bw.log 2>&1; echo "pytest tp bw$bw exit $?"; tail -n 3 $OUT/${TAG}_pytest_tp_bw$bw.log
done
run() { n=$1; shift; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline "$@" > $OUT/${TAG}_$n.json 2> $OUT/${TAG}_$n.err; echo "$n exit $?"; }
run cfg2_tp --cfg 2
RNNTB200_SWEEP=ws run cfg2_ws --cfg 2
run cfg1_tp --cfg 1
run cfg4_tp --cfg 4
RNNTB200_SWEEP=ws run cfg4_ws --cfg 4
for bw in 1 2 3 4; do RNNTB200_SWEEP_BW=$bw run cfg3_tp_bw$bw --cfg 3; done
RNNTB200_SWEEP=ws run cfg3_ws --cfg 3
run b128_tp --cfg 2 --batch 128
run cfg2_ragged_tp --cfg 2 --ragged
python - << PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    sat = (d.get("roofline") or {}).get("saturating_batch") or {}
    print(f.split("/")[-1], "ms", round(d["ms_per_step"], 4), "Gc/s", round(d["value"] / 1e9, 3),
          {k: round(v["us"], 1) for k, v in d.get("kernels", {}).items()}, "sat B", sat.get("B"), "us", round(sat.get("us", 0), 1), "frac", round(sat.get("frac", 0), 3))
PY
