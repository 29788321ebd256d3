#!/bin/bash
# Round-2 visit C: whole GPU suite on the new default sweep, DDP / training-step tests, tp cluster sweep
# (data-is-flag boundary) timings at cfg 3, cfg 5 at one GPU, ncu (launch list + full capture of the sweep).
TAG=${1:-r2c}; OUT=gpurun_out; mkdir -p $OUT
rm -f $OUT/parity_r2.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=8 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -n 25 $OUT/${TAG}_pytest.log
run() { n=$1; shift; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline "$@" > $OUT/${TAG}_$n.json 2> $OUT/${TAG}_$n.err; echo "$n exit $?"; tail -n 3 $OUT/${TAG}_$n.err; }
run cfg2 --cfg 2
for bw in 2 3 4; do RNNTB200_SWEEP_BW=$bw run cfg3_tp_bw$bw --cfg 3; done
run cfg5 --cfg 5 --steps 10 --warmup 3
python - << PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    sat = (d.get("roofline") or {}).get("saturating_batch") or {}
    print(f.split("/")[-1], "ms", round(d["ms_per_step"], 4), "Gc/s", round(d["value"] / 1e9, 3), "e2e", round(d["e2e"]["value"] / 1e9, 3),
          {k: round(v["us"], 1) for k, v in d.get("kernels", {}).items()}, "sat", sat.get("B"), round(sat.get("us", 0), 1), round(sat.get("frac", 0), 3), d.get("joint_loss_share_of_step"))
PY
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lattice_sweep -s 6 -c 2 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/${TAG}_prof* $OUT/${TAG}_launches.csv
