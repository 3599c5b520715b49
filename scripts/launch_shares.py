"""ncu launch list (--metrics gpu__time_duration.sum --csv --log-file ...) -> per-kernel share of the profiled steps.
Usage: launch_shares.py launches.csv [title]"""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr = rows[0]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
iu = hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1e-3)
    name = r[ik].split("(")[0].strip()
    tot[name] += float(r[iv].replace(",", "")) * scale
    cnt[name] += 1
total = sum(tot.values())
print(f"{sys.argv[2] if len(sys.argv) > 2 else 'ncu launch list'} (gpu__time_duration.sum, --clock-control none): total {total:.1f} us")
for name, t in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{100 * t / total:6.2f} %  {t:10.1f} us  {cnt[name]:4d} launches  {name}")
