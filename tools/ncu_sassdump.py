"""Per-SASS-instruction dump of an `ncu --page source --csv --print-source sass` file:
address, warp-instructions executed, stall samples, the two dominant stall reasons.  usage: ncu_sassdump.py sass.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ad = hdr.index('Address'); src = hdr.index('Source'); ie = hdr.index('Instructions Executed'); sm = hdr.index('# Samples')
stc = [k for k, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
for r in rows[2:]:
    if len(r) <= ie: continue
    try: n = int(r[ie]); s = int(r[sm] or 0)
    except ValueError: continue
    st = sorted(((int(r[k] or 0), hdr[k][6:]) for k in stc), reverse=True)[:2]
    print(f"{r[ad][-5:]} {n:9d} {s:6d} {r[src][:80]:80s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")
