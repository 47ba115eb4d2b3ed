"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total time, share."""
import csv, re, sys
from collections import defaultdict

for path in sys.argv[1:]:
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[mv].replace(",", ""))
        v = v / 1000.0 if r[mu] in ("ns", "nsecond") else (v * 1000.0 if r[mu] in ("ms", "msecond") else v)
        name = re.sub(r"^void |\(anonymous namespace\)::|<unnamed>::", "", r[kn])
        name = re.sub(r"\(.*", "", name)
        tot[name][0] += 1
        tot[name][1] += v
    total = sum(v[1] for v in tot.values())
    print(f"## {path}: {sum(v[0] for v in tot.values())} launches, {total / 1000:.2f} ms (serialised, under ncu)")
    print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"| `{name[:70]}` | {n} | {us:.0f} | {100 * us / total:.1f}% | {us / n:.1f} |")
    print()
