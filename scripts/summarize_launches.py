"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr = i; break
H = rows[hdr]
kn, mv, mu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    u = r[mu]
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9}.get(u, 1)
    name = re.sub(r"\(.*", "", r[kn])
    name = re.sub(r"<unnamed>::", "", name)
    agg[name][0] += 1; agg[name][1] += ns
tot = sum(v[1] for v in agg.values())
print(f"total {tot/1e6:.2f} ms over {sum(v[0] for v in agg.values())} launches")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 45]:
    print(f"{t/1e6:9.3f} ms {100*t/tot:5.1f}%  n={n:5d}  avg {t/n/1e3:8.1f} us  {name[:110]}")
