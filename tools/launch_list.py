"""profiles/<name>.md from an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_list.py in.csv out.md "<title>" """
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[mv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu], 1e-6)
    name = r[kn].split("(")[0].replace("void ", "").replace("ssimu2::", "")
    tot[name][0] += 1; tot[name][1] += v
s = sum(v[1] for v in tot.values())
out = [f"# {sys.argv[3]}", "", "| kernel | launches | total ms | avg ms | share |", "|---|---|---|---|---|"]
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {n} | {t:.3f} | {t / n:.4f} | {100 * t / s:.1f} % |")
open(sys.argv[2], "w").write("\n".join(out) + "\n")
print("\n".join(out))
