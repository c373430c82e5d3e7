"""Sum the warp-stall sampling columns of an ncu source-page CSV per kernel. usage: ncu_stalls.py src.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
i = 0; done = set()
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path" and i + 2 < len(rows):
        fn = rows[i + 1][1]; hdr = rows[i + 2]; j = i + 3; body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            body.append(rows[j]); j += 1
        i = j
        if fn in done or "Address" not in hdr: continue
        ad = hdr.index("Address")
        cols = [k for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        seen = set(); tot = collections.Counter()
        for r in body:
            if len(r) > ad and r[ad] and r[ad] not in seen:
                seen.add(r[ad])
                for k in cols:
                    try: tot[hdr[k]] += int(r[k])
                    except (ValueError, IndexError): pass
        if not tot: continue
        done.add(fn); s = sum(tot.values())
        print(fn[:50]); print("   " + "  ".join(f"{k[6:]}={100*v/s:.1f}%" for k, v in tot.most_common(9)))
    else:
        i += 1
