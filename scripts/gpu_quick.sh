#!/bin/bash
# Short GPU-box visit: the concat-GELU parity tests, the default bench, optionally one ncu capture.
# usage: scripts/gpu_quick.sh <tag> [kernel-regex or "none"] [pytest target] [extra bench args]
TAG=${1:-q}; KREGEX=${2:-none}; TESTS=${3:-tests/test_gpu_joint_cg.py}; shift 3 || true
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest $TESTS -m gpu -q -x --timeout 600 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -n 25 $OUT/${TAG}_pytest.log
python bench.py --steps 100 --warmup 10 --no-cpu-baseline "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
python - << PY
import json
d = json.load(open("$OUT/${TAG}_bench.json"))
print(d["ms_per_step"], d["value"] / 1e9, d["e2e"]["value"] / 1e9, {k: round(v["us"], 1) for k, v in d["kernels"].items()})
PY
tail -n 5 $OUT/${TAG}_bench.err
if [ "$KREGEX" != "none" ]; then
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 9 -c 6 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > $OUT/${TAG}_ncu_full.log 2>&1
fi
