python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/exp1.json 2> gpurun_out/exp1.err
python - << PY
import json
d = json.load(open("gpurun_out/exp1.json"))
print("EXP", d["ms_per_step"], {k: round(x["us"], 1) for k, x in d["kernels"].items()})
PY
