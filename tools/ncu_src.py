"""Top stall lines of one kernel from an ncu report's source page: python tools/ncu_src.py REPORT KERNEL_REGEX [N]."""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}",
                      "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]
si = hdr.index("# Samples")
src = hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[1:]:
    if len(r) <= si or r[0].startswith("Kernel") or r[0] == "Address":
        continue
    try:
        n = int(r[si])
    except ValueError:
        continue
    data.append((n, r))
tot = sum(n for n, _ in data)
print(f"total samples {tot}")
for n, r in sorted(data, key=lambda x: -x[0])[:top]:
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{n:7d} {100 * n / tot:5.1f}%  {r[src][:90]:90s} {st}")
