#!/usr/bin/env python
"""Turns the ncu outputs a GPU visit brought back (gpurun_out/) into the small, committed summaries
under profiles/: one launch list (per-kernel time share) and one table of the counters that back
the roofline / tensor-pipe claims.  Needs only the `ncu` CLI (no GPU).

    python scripts/summarize_profiles.py <tag> [<tag> ...]      # e.g. r1j at2
"""
from __future__ import annotations

import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

COUNTERS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def ncu_raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        return [], [], []
    return rows[0], rows[1], rows[2:]


def summarize_full(tag):
    rep = os.path.join(OUT, f"{tag}_prof.ncu-rep")
    if not os.path.exists(rep):
        return None
    hdr, units, rows = ncu_raw(rep)
    seen, lines = set(), [f"# ncu --set full summary ({tag}_prof.ncu-rep, --clock-control none)\n"]
    for r in rows:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "?")
        if name in seen:
            continue
        seen.add(name)
        lines.append(f"\n## {name[:110]}\n\n| counter | value | unit |\n|---|---|---|")
        for key, label in COUNTERS:
            if key in d and d[key] != "":
                lines.append(f"| {label} (`{key}`) | {d[key]} | {u.get(key, '')} |")
        stalls = []
        for k, v in d.items():
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    if float(v) >= 0.2:
                        stalls.append((float(v), k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        if stalls:
            lines.append("\nwarp stall cycles per issued instruction: " +
                         ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)))
    path = os.path.join(PROF, f"{tag}_ncu_full.md")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    # DRAM traffic per launch of every kernel seen (bench.py's roofline.traffic reads this file)
    traffic_path = os.path.join(PROF, "dram_traffic.json")
    try:
        traffic = json.load(open(traffic_path))
    except Exception:
        traffic = {}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows:
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        try:
            tot = sum(float(d[k]) * scale.get(u.get(k, "byte"), 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        except (KeyError, ValueError):
            continue
        name = d.get("Kernel Name", "?")
        for key in ("proj_tc_bwd_kernel", "proj_tc_kernel", "cg_lse", "cg_grad", "lattice_sweep", "at_lse_tc", "at_grad_tc"):
            if key in name:
                traffic[key] = {"dram_bytes_per_launch": tot, "source": f"{tag}_prof.ncu-rep (ncu --set full)"}
                break
    with open(traffic_path, "w") as f:
        json.dump(traffic, f, indent=1, sort_keys=True)
    return path


def summarize_launches(tag):
    src = os.path.join(OUT, f"{tag}_launches.csv")
    if not os.path.exists(src):
        return None
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
    if not hi:
        return None
    hdr, data = rows[hi[0]], rows[hi[0] + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values()) or 1.0
    path = os.path.join(PROF, f"{tag}_launches.csv")
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "avg_us", "share_pct", "ours"])
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k[:160], v[0], f"{v[1] / v[0] / 1e3:.2f}", f"{100 * v[1] / tot:.2f}", int("rnntb200" in k)])
    return path


def main():
    os.makedirs(PROF, exist_ok=True)
    for tag in sys.argv[1:]:
        for fn in (summarize_full, summarize_launches):
            p = fn(tag)
            print(p or f"(nothing for {tag} / {fn.__name__})")


if __name__ == "__main__":
    main()
