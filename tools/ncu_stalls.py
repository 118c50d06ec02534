"""Stall-reason totals of one kernel from an ncu report's source page, overall and per CUDA source line.

    python tools/ncu_stalls.py REPORT KERNEL_REGEX [top_lines]
"""
import collections
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
hdr_idx = [i for i, l in enumerate(lines) if l.startswith('"Address"') or l.startswith('"#"') or l.startswith('"Line')]
start = next(i for i, l in enumerate(lines) if '"# Samples"' in l or "# Samples" in l)
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]
si = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
print("columns:", [h for h in hdr if not h.startswith("stall_")][:12])
tot = collections.Counter()
per_line = collections.defaultdict(collections.Counter)
key_col = 0
for r in rows[1:]:
    if len(r) <= si:
        continue
    try:
        n = int(r[si])
    except ValueError:
        continue
    for i in stall_cols:
        try:
            v = int(r[i] or 0)
        except ValueError:
            v = 0
        tot[hdr[i]] += v
        per_line[r[key_col] + " | " + (r[hdr.index("Source")][:100] if "Source" in hdr else "")][hdr[i]] += v
s = sum(tot.values()) or 1
print("total stall samples", s)
for k, v in tot.most_common():
    if v:
        print(f"  {k:32s} {v:8d} {100 * v / s:5.1f}%")
print("--- top lines")
for k, c in sorted(per_line.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    t = sum(c.values())
    print(f"{t:7d} {100 * t / s:5.1f}%  {k[:120]:120s} {c.most_common(3)}")
