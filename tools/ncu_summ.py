"""Summarise ncu outputs into small CSVs for profiles/ (launch list -> per-kernel shares; full capture -> key metrics)."""
import collections
import csv
import re
import subprocess
import sys


def launches(path, out, note):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    n = 0
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        n += 1
        name = re.sub(r"\(.*", "", r[ki]).replace("void pcg::<unnamed>::", "").replace("pcg::<unnamed>::", "").replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# {note}\n# launches {n}, total {tot / 1e6:.2f} ms (cold-cache, serialised: compare shares)\n")
        f.write("ms,share_pct,launches,avg_us,kernel\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t / 1e6:.3f},{100 * t / tot:.2f},{c},{t / c / 1e3:.1f},{k[:110]}\n")


PAT = re.compile(r"^(Kernel Name|gpu__time_duration.sum|dram__bytes_(read|write).sum|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|"
                 r"gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|"
                 r"launch__grid_size|sm__throughput.avg.pct_of_peak_sustained_elapsed|lts__t_sector_hit_rate.pct|l1tex__throughput.avg.pct_of_peak_sustained_active|"
                 r"lts__throughput.avg.pct_of_peak_sustained_elapsed|smsp__inst_executed.sum)$")


def full(rep, out, note):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = [i for i, h in enumerate(hdr) if PAT.match(h)]
    stall = [i for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h]
    with open(out, "w") as f:
        f.write(f"# {note}\n")
        w = csv.writer(f)
        w.writerow([hdr[i] for i in keep] + ["top_stalls"])
        w.writerow([units[i] for i in keep] + [""])
        for r in rows[2:]:
            st = []
            for i in stall:
                try:
                    st.append((float(r[i].replace(",", "")), hdr[i].replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
            tot = sum(v for v, _ in st) or 1.0
            top = "; ".join(f"{h} {100 * v / tot:.0f}%" for v, h in sorted(st, reverse=True)[:5])
            w.writerow([r[i] for i in keep] + [top])


if __name__ == "__main__":
    kind, src, dst, note = sys.argv[1:5]
    (launches if kind == "launches" else full)(src, dst, note)
