"""Per-source-line instruction counts of one kernel from an ncu report (needs -lineinfo + --import-source on).
usage: python scripts/ncu_lines.py <report.ncu-rep> <kernel-regex> [top N]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      "regex:" + rx, "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, data = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) > 8 and r[0].isdigit() and r[2] == "-":  # a source line (not one of its SASS rows)
        ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
        try:
            data.append((int(r[ie]), int(r[sm]), fname, int(r[0]), r[1][:100]))
        except ValueError:
            pass
tot, tots = sum(d[0] for d in data), sum(d[1] for d in data)
print(f"total warp instructions {tot}, samples {tots}")
for d in sorted(data, reverse=True)[:top]:
    print(f"{d[0]:>10} {100 * d[0] / tot:5.1f}%  smp {100 * d[1] / max(tots, 1):5.1f}%  {d[2]}:{d[3]:<4} {d[4]}")

# optional: totals per line range of the main file, e.g. ranges=300-360,361-420
import os
rng = os.environ.get("RANGES")
if rng:
    main = max(set(d[2] for d in data), key=lambda f: sum(d[0] for d in data if d[2] == f))
    for part in rng.split(","):
        lo, hi = map(int, part.split("-"))
        n = sum(d[0] for d in data if d[2] == main and lo <= d[3] <= hi)
        sm = sum(d[1] for d in data if d[2] == main and lo <= d[3] <= hi)
        print(f"{main}:{lo}-{hi}: {n} ({100 * n / tot:.1f}%), samples {100 * sm / max(tots, 1):.1f}%")
    other = sum(d[0] for d in data if d[2] != main)
    print(f"other files (inlined helpers): {other} ({100 * other / tot:.1f}%)")
