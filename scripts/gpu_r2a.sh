#!/bin/bash
# Round-2 visit A: full GPU test-suite (incl. the oracle-pinned full-size tests), default bench with the
# gpu_baseline / cpu_baseline legs, the sweep at a saturating batch (ws vs single-role A/B), cfg 3 / 4
# bench lines with the torchaudio-GPU bar, compute-sanitizer logs.
TAG=${1:-r2a}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.csv 2>&1
nproc > $OUT/${TAG}_host.txt; free -g >> $OUT/${TAG}_host.txt
rm -f $OUT/parity_r2.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=15 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -n 30 $OUT/${TAG}_pytest.log
python bench.py --steps 50 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
tail -c 3000 $OUT/${TAG}_bench.json; tail -n 5 $OUT/${TAG}_bench.err
for b in 128 512; do
  python bench.py --steps 20 --warmup 5 --batch $b --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_b${b}_ws.json 2> $OUT/${TAG}_b${b}_ws.err; echo "b$b ws exit $?"
  RNNTB200_SWEEP_LEGACY=1 python bench.py --steps 20 --warmup 5 --batch $b --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_b${b}_legacy.json 2> $OUT/${TAG}_b${b}_legacy.err; echo "b$b legacy exit $?"
done
for c in 3 4; do
  timeout 600 python bench.py --steps 20 --warmup 5 --cfg $c --no-cpu-baseline > $OUT/${TAG}_cfg${c}.json 2> $OUT/${TAG}_cfg${c}.err; echo "cfg$c exit $?"
done
python - << PY
import json, glob
for f in sorted(glob.glob("$OUT/${TAG}_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    gb = d.get("gpu_baseline", {})
    print(f, round(d["ms_per_step"], 4), round(d["value"] / 1e9, 3), "e2e", round(d["e2e"]["value"] / 1e9, 3),
          {k: round(v["us"], 1) for k, v in d.get("kernels", {}).items()},
          "sat", (d.get("roofline") or {}).get("saturating_batch"), "gpu_baseline", gb.get("ms_per_step"), gb.get("fp16_autocast"), gb.get("unavailable"))
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_reference.json 2> $OUT/${TAG}_reference.err; echo "reference exit $?"; tail -c 800 $OUT/${TAG}_reference.json
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > $OUT/${TAG}_sanitizer_$tool.log 2>&1; echo "$tool exit $?"; tail -n 4 $OUT/${TAG}_sanitizer_$tool.log
done
