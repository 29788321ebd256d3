#!/bin/bash
# Round-2 visit E: AMP / decode / comm tests, tp cluster sweep (boundary fetch hoisted out of the step loop) at cfg 3.
TAG=${1:-r2e}; OUT=gpurun_out; mkdir -p $OUT
T="tests/test_gpu_loss.py tests/test_gpu_joint_cg.py tests/test_gpu_amp.py tests/test_gpu_decode.py tests/test_gpu_comm.py tests/test_gpu_ddp.py"
timeout 900 python -m pytest $T -m gpu -q --timeout 600 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -n 6 $OUT/${TAG}_pytest.log
for bw in 2 3 4; do
RNNTB200_SWEEP=tp RNNTB200_SWEEP_BW=$bw timeout 600 python -m pytest tests/test_gpu_loss.py -m gpu -q --timeout 600 -x -k "boundaries or long_lattice" > $OUT/${TAG}_pytest_tpcl$bw.log 2>&1; echo "pytest tp cluster bw$bw exit $?"; tail -n 2 $OUT/${TAG}_pytest_tpcl$bw.log
done
run() { n=$1; shift; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline "$@" > $OUT/${TAG}_$n.json 2> $OUT/${TAG}_$n.err; echo "$n exit $?"; tail -n 2 $OUT/${TAG}_$n.err; }
run cfg2 --cfg 2
for bw in 2 3 4; do RNNTB200_SWEEP=tp RNNTB200_SWEEP_BW=$bw run cfg3_tpcl$bw --cfg 3; done
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
