#!/usr/bin/env python
"""Markdown table of gpurun_out/<tag>_*.json bench lines (scripts/gpu_numbers.sh) for DESIGN.md."""
import glob
import json
import os
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "num"
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
rows = []
for path in sorted(glob.glob(os.path.join(root, f"{tag}_*.json"))):
    name = os.path.basename(path)[len(tag) + 1:-5]
    line = None
    for l in open(path):
        if l.startswith("{"):
            line = json.loads(l)
    if line is None:
        rows.append((name, None))
        continue
    rows.append((name, line))
print("| run | workload | joint/gemm | ms/step | Gcells/s | utt/s | e2e Gcells/s | lattice sweep alone (us, GB/s, frac of HBM; at a saturating batch) | reference GPU path ms/step (fp32 / fp16 autocast) |")
print("|---|---|---|---|---|---|---|---|---|")
for name, d in rows:
    if d is None:
        print(f"| {name} | (no JSON) | | | | | | |")
        continue
    c = d["config"]
    rf = d.get("roofline") or {}
    wl = f"B={c['B_per_gpu']} T={c['T']} U={c['U']} V={c['V']} H={c['H']}{' ragged' if c['ragged'] else ''}"
    if d.get("impl") == "reference":
        print(f"| {name} | {wl} | CPU oracle port, {d['cpu_baseline']['cores']} threads | {d['ms_per_step']:.1f} | "
              f"{d['value'] / 1e9:.6f} | {d['utterances_per_s']:.1f} | - | - | - |")
        continue
    print(f"| {name} | {wl} | {c['joint']}/{c['gemm']} | {d['ms_per_step']:.3f} | {d['value'] / 1e9:.3f} | "
          f"{d['utterances_per_s']:.0f} | {d['e2e']['value'] / 1e9:.3f} | "
          f"({rf.get('us_per_launch', 0):.0f}, {rf.get('achieved', 0):.0f}, {rf.get('frac', 0):.3f}; "
          f"B={(rf.get('saturating_batch') or {}).get('B', '-')}: {(rf.get('saturating_batch') or {}).get('us', 0):.0f} us, {(rf.get('saturating_batch') or {}).get('frac', 0):.3f}) | "
          f"{(d.get('gpu_baseline') or {}).get('ms_per_step', float('nan')):.1f} / {((d.get('gpu_baseline') or {}).get('fp16_autocast') or {}).get('ms_per_step', float('nan')):.1f} |")
    ks = d.get("kernels") or {}
    if ks:
        print("|  | kernels (us): " + ", ".join(f"{k} {v['us']:.0f}" for k, v in ks.items()) + " | | | | | | | |")
