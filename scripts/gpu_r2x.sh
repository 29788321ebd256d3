#!/bin/bash
# Round-2 visit "x": the whole GPU suite on the code with the wide concat-GELU kernels, smoke, the cfg 4 lines of the
# results table (tag r2hn), the driver-style default bench, and one ncu --set full capture of the two wide kernels.
TAG=${1:-r2x}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -n 6 $OUT/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 $OUT/${TAG}_smoke.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-gpu-baseline --cfg 4 > $OUT/r2hn_cg_cfg4.json 2> $OUT/r2hn_cg_cfg4.err; echo "cg_cfg4 exit $?"
timeout 600 python bench.py --steps 20 --warmup 5 --cfg 4 --no-cpu-baseline > $OUT/r2hn_gpubar_cfg4.json 2> $OUT/r2hn_gpubar_cfg4.err; echo "gpubar cfg4 exit $?"
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-gpu-baseline --cfg 4 --deterministic > $OUT/${TAG}_cg_cfg4_det.json 2> $OUT/${TAG}_cg_cfg4_det.err; echo "cg_cfg4 det exit $?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_bench_driver.json 2> $OUT/${TAG}_bench_driver.err; echo "bench driver-style exit $?"
python - << PY
import json
for n in ("r2hn_cg_cfg4", "r2hn_gpubar_cfg4", "${TAG}_cg_cfg4_det", "${TAG}_bench_driver"):
    try:
        d = json.load(open("$OUT/%s.json" % n))
        print(n, "ms", round(d["ms_per_step"], 4), "Gcells/s", round(d["value"] / 1e9, 4), "e2e", round(d["e2e"]["value"] / 1e9, 4),
              {k: round(v["us"], 1) for k, v in (d.get("kernels") or {}).items()}, (d.get("gpu_baseline") or {}).get("ms_per_step"))
    except Exception as e:
        print(n, "no line:", e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"cg_lse_mmw|cg_grad_mm|cg_factor_rows_wide" -s 9 -c 3 -f -o $OUT/${TAG}_prof_wide \
    python bench.py --cfg 4 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/${TAG}_prof_wide*
