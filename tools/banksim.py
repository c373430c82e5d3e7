def wavefronts(addrs_bytes, width):
    per = {16: 8, 8: 16, 4: 32}[width]
    tot = 0
    for q0 in range(0, 32, per):
        banks = {}
        for l in range(q0, q0 + per):
            a = addrs_bytes[l]
            for w in range(width // 4):
                word = a // 4 + w
                banks.setdefault(word % 32, set()).add(word)
        tot += max((len(v) for v in banks.values()), default=0)
    return tot
kXR, kXInW = 12, 76
order=[0,2,3,1,4]
rowmap={0:[0,1,2,3,4,5],2:[2,3,0,1,4,5],3:[0,1,4,5,2,3],1:[0,1,2,3,4,5],4:[0,1,2,3,4,5]}
perm={0:0,1:3,2:2,3:1,4:4}
def sim(ch, oa, ob, hbpitch=68, hbplane=824):
    inplane = kXR * kXInW
    res = {}
    def lane_qrp(l):
        l = min(l, 29); q = order[l // 6]; return q, rowmap[q][l % 6]
    for name, which, rowadd in (("xa", "x", 0), ("xb", "x", 6), ("ya", "y", 0), ("yb", "y", 6)):
        addrs = []
        for l in range(32):
            q, rp = lane_qrp(l)
            px = 3 + ch if q in (1, 4) else ch
            py = ch if q == 0 else (3 + ch if q in (1, 2) else -1)
            pl = px if which == "x" else py
            a = (6*inplane*4 + (oa if rowadd == 0 else ob)) if pl < 0 else (pl * inplane + (rp + rowadd) * kXInW) * 4
            addrs.append(a)
        res[name] = wavefronts(addrs, 16)
    for name, rowadd in (("sa", 0), ("sb", 6)):
        addrs = [((perm[lane_qrp(l)[0]] * 3 + ch) * hbplane + (lane_qrp(l)[1] + rowadd) * hbpitch) * 4 for l in range(32)]
        res[name] = wavefronts(addrs, 16)
    return res
for c in range(3):
    b=None
    for oa in range(0,128,16):
        for ob in range(0,128,16):
            r=sim(c,oa,ob); t=sum(r.values())
            if b is None or t<b[0]: b=(t,oa,ob,r)
    print(c,b)
