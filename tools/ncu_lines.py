"""Summarise an ncu --page source --csv dump: executed warp instructions and stall samples per CUDA source line.
usage: ncu -i rep --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_lines.py src.csv [kernel-substr] [top]"""
import csv, sys, collections
path = sys.argv[1]; sel = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
# the dump is a sequence of blocks: "File Path", "Function Name", header, rows...
blocks = []; i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path":
        f = rows[i][1]; fn = rows[i + 1][1]; hdr = rows[i + 2]; j = i + 3; body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            body.append(rows[j]); j += 1
        blocks.append((f, fn, hdr, body)); i = j
    else:
        i += 1
agg = collections.defaultdict(lambda: [0, 0, ""])
tot = collections.Counter()
for f, fn, hdr, body in blocks:
    if sel not in fn: continue
    if "Instructions Executed" not in hdr: continue
    ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
    ln = hdr.index("Line No") if "Line No" in hdr else None
    srcs = [k for k, h in enumerate(hdr) if h == "Source"]
    if ln is None: continue
    for r in body:
        if len(r) <= ie or not r[ie]: continue
        try: n = int(r[ie]); s = int(r[sm] or 0)
        except ValueError: continue
        key = (fn[:40], f.split("/")[-1], r[ln])
        agg[key][0] += n; agg[key][1] += s
        if not agg[key][2]: agg[key][2] = r[srcs[0]].strip()[:90]
        tot[fn[:40]] += n
for fn, n in tot.items(): print(f"TOTAL {fn}: {n/1e6:.1f} M warp-instr")
for key, (n, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n/1e6:9.1f}M {s:7d}smp  {key[1]}:{key[2]:>4s}  {src}")
