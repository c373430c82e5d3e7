"""Opcode histogram weighted by executed warp instructions from an ncu source-page CSV (sass view).
usage: python tools/ncu_sass.py src.csv kernel-substr [top]"""
import csv, sys, collections, re
path, sel = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rows = list(csv.reader(open(path)))
i = 0; done = set()
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path" and i + 2 < len(rows):
        fn = rows[i + 1][1]; hdr = rows[i + 2]; j = i + 3; body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            body.append(rows[j]); j += 1
        i = j
        if sel in fn and fn not in done and "Address" in hdr:
            ad = hdr.index("Address"); ie = hdr.index("Instructions Executed")
            srcs = [k for k, h in enumerate(hdr) if h == "Source"]
            seen = {}
            for r in body:
                if len(r) > ie and r[ad] and r[ie].isdigit():
                    seen[r[ad]] = (r[srcs[-1]], int(r[ie]))
            if not seen: continue
            done.add(fn)
            c = collections.Counter(); tot = 0
            for a, (ins, n) in seen.items():
                op = re.sub(r"^@!?U?P\d+\s+", "", ins.strip()).split()[0].split(".")[0] if ins.strip() else "?"
                c[op] += n; tot += n
            print(fn[:60], "total", tot / 1e6, "M over", len(seen), "sass instrs")
            for op, n in c.most_common(top): print(f"   {op:10s} {n/1e6:9.1f}M {100*n/tot:5.1f}%")
    else:
        i += 1
