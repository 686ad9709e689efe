"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total time, share."""
import csv, re, sys, collections
src, title = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("pose::", "").replace("(int)", "")
    t = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += t
tot = sum(v[1] for v in agg.values())
print(f"# {title}: per-kernel share of device time (ncu gpu__time_duration.sum, cold-cache serialised; shares, not absolutes)\n")
print(f"{sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms under ncu\n")
print("| kernel | launches | us | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if t / tot < 0.001:
        continue
    print(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f} % |")
