"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list as a markdown table.
Usage: python tools/launch_summary.py launches.csv "<command line that was profiled>" > profiles/xxx.md"""
import csv, sys, collections, re

path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).strip()
    v = float(r[vi].replace(",", ""))
    us = v / 1e3 if r[ui] in ("ns", "nsecond") else (v if r[ui] in ("us", "usecond") else v * 1e3)
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1; t[1] += us
total = sum(v[1] for v in tot.values()); n = sum(v[0] for v in tot.values())
print("# ncu launch list of `%s`\n" % cmd)
print("`ncu --metrics gpu__time_duration.sum --clock-control none --csv` — cold-cache, serialised: compare SHARES with the")
print("in-situ profile (torch.profiler, no replay), not absolute times.  %d launches, %.1f ms.\n" % (n, total / 1e3))
print("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:24]:
    print("| `%s` | %d | %.3f | %.1f%% | %.1f |" % (k[:90], c, t / 1e3, 100 * t / total, t / c))
g = sum(t for k, (c, t) in tot.items() if "gemm" in k)
a = sum(t for k, (c, t) in tot.items() if "attn" in k)
print("\nGEMM share: %.1f%%; attention core share: %.1f%%." % (100 * g / total, 100 * a / total))
