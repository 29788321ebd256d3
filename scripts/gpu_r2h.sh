#!/bin/bash
# Round-2 visit H (final): the whole GPU suite, the results table (scripts/gpu_numbers.sh), an ncu capture of the
# warp-specialised sweep (what bounds it now that the chain is short), the launch list + sweep capture of the default step.
TAG=${1:-r2f}; OUT=gpurun_out; mkdir -p $OUT
rm -f $OUT/parity_r2.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=6 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -n 14 $OUT/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 $OUT/${TAG}_smoke.log
bash scripts/gpu_numbers.sh ${TAG}n
python scripts/make_results_table.py ${TAG}n > $OUT/${TAG}_table.md 2>&1; head -c 6000 $OUT/${TAG}_table.md
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lattice_sweep -s 6 -c 1 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/${TAG}_prof*
