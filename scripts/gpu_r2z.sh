#!/bin/bash
# Round-2 last visit: the driver's own commands on the final code (default bench, reference arm, smoke), the ncu launch
# list of the default bench command and one --set full capture of the sweep.
TAG=${1:-r2z}; OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 $OUT/${TAG}_smoke.log
python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; echo "bench default exit $?"
python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_bench_driver.json 2> $OUT/${TAG}_bench_driver.err; echo "bench driver-style exit $?"
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_reference.json 2> $OUT/${TAG}_reference.err; echo "reference exit $?"
python - << PY
import json
for n in ("bench_default", "bench_driver", "reference"):
    d = json.load(open("$OUT/${TAG}_%s.json" % n))
    print(n, "ms", round(d["ms_per_step"], 4), "value", round(d["value"] / 1e9, 6), "e2e", round(d["e2e"]["value"] / 1e9, 6), (d.get("roofline") or {}).get("frac"),
          ((d.get("roofline") or {}).get("saturating_batch") or {}).get("frac"), (d.get("gpu_baseline") or {}).get("ms_per_step"), (d.get("cpu_baseline") or {}).get("value"))
PY
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lattice_sweep -s 6 -c 1 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/${TAG}_prof* $OUT/${TAG}_launches.csv
