#!/bin/bash
# Round-2 visit "w": the wide (128 < V <= 4096) factorised concat-GELU kernels -- parity tests, cfg 4 timings, and a
# cfg 2 line to confirm the V <= 128 path is where it was.
TAG=${1:-r2w}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_joint_cg_wide.py tests/test_gpu_joint_cg.py "tests/test_gpu_fullsize.py::test_cfg4_large_vocab_vs_oracle" \
    -q -m gpu --tb=short -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -n 40 $OUT/${TAG}_pytest.log
for cfg in 4 2; do
  timeout 300 python bench.py --cfg $cfg --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_bench_cfg$cfg.json 2> $OUT/${TAG}_bench_cfg$cfg.err; echo "bench cfg $cfg exit $?"
done
python - << PY
import json
for n in ("cfg4", "cfg2"):
    try:
        d = json.load(open("$OUT/${TAG}_bench_%s.json" % n))
        print(n, "ms", round(d["ms_per_step"], 4), "Gcells/s", round(d["value"] / 1e9, 4), {k: round(v["us"], 1) for k, v in (d.get("kernels") or {}).items()})
    except Exception as e:
        print(n, "no line:", e)
PY
