"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total / mean time, share."""
import csv, re, sys, collections
path = sys.argv[1]
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
hdr = rows[0]
ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"lsvs::\(anonymous namespace\)::|\(anonymous namespace\)::|void ", "", name)
    t = float(r[vi].replace(",", ""))
    t_us = t / 1e3 if r[ui] in ("ns", "nsecond") else (t if r[ui] in ("us", "usecond") else t * 1e3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t_us
tot = sum(v[1] for v in agg.values())
print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot/1e3:.2f} ms total (cold-cache, serialised; compare shares)")
print(f"{'kernel':70s} {'n':>6s} {'total_ms':>10s} {'mean_us':>9s} {'share':>7s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:6d} {t/1e3:10.3f} {t/n:9.1f} {100*t/tot:6.1f}%")
