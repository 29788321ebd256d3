#!/bin/bash
# Round-2 visit D: tp sweep after the instruction diet (edge through the rotate shuffle, fast-path loader / stores),
# the chain-warps-last experiment on the warp-specialised sweep, one ncu capture of the tp cluster sweep at cfg 3.
TAG=${1:-r2d}; OUT=gpurun_out; mkdir -p $OUT
T="tests/test_gpu_loss.py tests/test_gpu_joint_cg.py tests/test_abi.py tests/test_gpu_comm.py tests/test_gpu_amp.py"
timeout 900 python -m pytest $T -m gpu -q --timeout 600 -x > $OUT/${TAG}_pytest_tp.log 2>&1; echo "pytest tp exit $?"; tail -n 4 $OUT/${TAG}_pytest_tp.log
RNNTB200_SWEEP=ws RNNTB200_WS_CHAIN_LAST=1 timeout 900 python -m pytest tests/test_gpu_loss.py -m gpu -q --timeout 600 -x > $OUT/${TAG}_pytest_wslast.log 2>&1; echo "pytest ws chain-last exit $?"; tail -n 3 $OUT/${TAG}_pytest_wslast.log
RNNTB200_SWEEP=tp timeout 600 python -m pytest tests/test_gpu_loss.py -m gpu -q --timeout 600 -x -k "boundaries or long_lattice" > $OUT/${TAG}_pytest_tpcl.log 2>&1; echo "pytest tp cluster exit $?"; tail -n 3 $OUT/${TAG}_pytest_tpcl.log
run() { n=$1; shift; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline "$@" > $OUT/${TAG}_$n.json 2> $OUT/${TAG}_$n.err; echo "$n exit $?"; tail -n 2 $OUT/${TAG}_$n.err; }
run cfg2_tp --cfg 2
RNNTB200_SWEEP=ws run cfg2_ws --cfg 2
RNNTB200_SWEEP=ws RNNTB200_WS_CHAIN_LAST=1 run cfg2_wslast --cfg 2
run cfg1_tp --cfg 1
run cfg4_tp --cfg 4
RNNTB200_SWEEP=tp RNNTB200_SWEEP_BW=2 run cfg3_tpcl --cfg 3
run cfg3 --cfg 3
python - << PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    sat = (d.get("roofline") or {}).get("saturating_batch") or {}
    print(f.split("/")[-1], "ms", round(d["ms_per_step"], 4), "Gc/s", round(d["value"] / 1e9, 3),
          {k: round(v["us"], 1) for k, v in d.get("kernels", {}).items()}, "sat", sat.get("B"), round(sat.get("us", 0), 1), round(sat.get("frac", 0), 3))
PY
RNNTB200_SWEEP=tp RNNTB200_SWEEP_BW=2 ncu --set full --clock-control none --import-source on -k regex:lattice_sweep -s 6 -c 1 -f -o $OUT/${TAG}_prof_tpcl \
    python bench.py --cfg 3 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_ncu_tpcl.log 2>&1
ls -la $OUT/${TAG}_prof*
