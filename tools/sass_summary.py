"""Per-kernel SASS opcode evidence for profiles/: which kernels of libpcg.so carry tcgen05 (UTC*MMA), tensor-memory
loads / stores (LDTM / STTM), TMA (UTMALDG / UTMASTG / UBLKCP), legacy mma.sync (HMMA) ...

    python tools/sass_summary.py [libpcg.so] > profiles/r02_sass_opcodes.csv
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "perceptor_b200", "libpcg.so")
COLS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "HMMA", "MUFU.EX2", "MUFU.TANH",
        "RED", "ATOM", "LDGSTS", "BAR"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    counts[cur]["_total"] += 1
    for c in COLS:
        if op == c or op.startswith(c + ".") or (c in ("MUFU.EX2", "MUFU.TANH") and op.startswith(c)):
            counts[cur][c] += 1
demangled = subprocess.run(["cu++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
digest = subprocess.run(["sha256sum", lib], capture_output=True, text=True).stdout.split()[0]
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)} (sha256 {digest[:16]}...), instruction counts per kernel; "
      "UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAPF = cp.async.bulk.tensor load/store/prefetch, "
      "HMMA = mma.sync (legacy tensor path)")
print("kernel,instructions," + ",".join(COLS))
for (name, c), dm in zip(counts.items(), demangled):
    short = re.sub(r"\(.*", "", dm).replace("void ", "").replace("pcg::(anonymous namespace)::", "").replace("pcg::", "")
    print(f"\"{short}\",{c['_total']}," + ",".join(str(c[k]) for k in COLS))
