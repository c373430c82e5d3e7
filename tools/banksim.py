"""Shared-memory bank model of the 128-bit accesses of k_hv's H warps (development aid + tests/test_hv_layout.py).

A 128-bit shared-memory instruction is served one quarter-warp (8 lanes) at a time; a quarter costs as many
wavefronts as the largest number of distinct 32-bit words that fall into one of the 32 banks.  The ideal for a full warp
is therefore 4.  `simulate()` evaluates the four tile loads (x / y operand, row pair member A / B) and the two tile stores
of one scan step for a channel, given the lane -> (quantity, row pair) map and the tile geometry of ssimu2_kernels.cuh."""
import re


def wavefronts(addrs_bytes, width=16):
    per = {16: 8, 8: 16, 4: 32}[width]
    tot = 0
    for q0 in range(0, 32, per):
        banks = {}
        for lane in range(q0, q0 + per):
            a = addrs_bytes[lane]
            for w in range(width // 4):
                word = a // 4 + w
                banks.setdefault(word % 32, set()).add(word)
        tot += max((len(v) for v in banks.values()), default=0)
    return tot


def nibbles(x, n):
    return [(x >> (4 * i)) & 7 for i in range(n)]


def kernel_params(path):
    """Pull the layout constants out of ssimu2_kernels.cuh so the test follows the kernel, not a copy of it."""
    s = open(path).read()
    def const(name):
        m = re.search(r"constexpr\s+\w+\s+%s\s*=\s*([^;]+);" % name, s)
        return m.group(1).strip()
    kXR = int(const("kXR"))
    kXInW = int(const("kXInW"))
    pitch = int(const("kXHbPitch"))
    plane = eval(const("kXHbPlane"), {"kXR": kXR, "kXHbPitch": pitch})
    order = nibbles(int(re.search(r"(?:const int )?q = \((0x[0-9a-fA-F]+) >> \(4 \* qidx\)\) & 7;", s).group(1), 16), 5)
    m = re.search(r"(?:const int )?rp = q == 2 \? \(\((0x[0-9a-fA-F]+) >> \(4 \* jj\)\) & 7\) : \(q == 3 \? \(\((0x[0-9a-fA-F]+) >>", s)
    rowmap = {q: list(range(6)) for q in range(5)}
    rowmap[2] = nibbles(int(m.group(1), 16), 6)
    rowmap[3] = nibbles(int(m.group(2), 16), 6)
    m = re.search(r"hv_slot\(int q\) \{ return q == 1 \? (\d) : \(q == 3 \? (\d) : q\); \}", s)
    slot = {0: 0, 1: int(m.group(1)), 2: 2, 3: int(m.group(2)), 4: 4}
    m = re.search(r"onesA = sbase \+ kXOffOnes \+ \(ch == 1 \? (\d+)u : (\d+)u\), onesB = sbase \+ kXOffOnes \+ \(ch == 1 \? (\d+)u : (\d+)u\)", s)
    ones = {"A": {1: int(m.group(1)), 0: int(m.group(2)), 2: int(m.group(2))}, "B": {1: int(m.group(3)), 0: int(m.group(4)), 2: int(m.group(4))}}
    return dict(kXR=kXR, kXInW=kXInW, pitch=pitch, plane=plane, order=order, rowmap=rowmap, slot=slot, ones=ones)


def simulate(P, ch, ones_base=None):
    """Wavefronts of {xa, xb, ya, yb, sa, sb}; the tile base and the ones rows are 128-byte aligned in the kernel."""
    inplane = P["kXR"] * P["kXInW"]
    if ones_base is None:
        ones_base = 6 * inplane * 4
        assert ones_base % 128 == 0

    def lane_qrp(lane):
        lane = min(lane, 29)
        q = P["order"][lane // 6]
        return q, P["rowmap"][q][lane % 6]

    res = {}
    for name, which, rowadd in (("xa", "x", 0), ("xb", "x", 6), ("ya", "y", 0), ("yb", "y", 6)):
        addrs = []
        for lane in range(32):
            q, rp = lane_qrp(lane)
            px = 3 + ch if q in (1, 4) else ch
            py = ch if q == 0 else (3 + ch if q in (1, 2) else -1)
            pl = px if which == "x" else py
            if pl < 0:
                addrs.append(ones_base + P["ones"]["A" if rowadd == 0 else "B"][ch])
            else:
                addrs.append((pl * inplane + (rp + rowadd) * P["kXInW"]) * 4)
        res[name] = wavefronts(addrs)
    for name, rowadd in (("sa", 0), ("sb", 6)):
        addrs = []
        for lane in range(32):
            q, rp = lane_qrp(lane)
            addrs.append(((P["slot"][q] * 3 + ch) * P["plane"] + (rp + rowadd) * P["pitch"]) * 4)
        res[name] = wavefronts(addrs)
    return res


if __name__ == "__main__":
    import os, sys
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "turbo_metrics_b200", "csrc",
                                                                 "ssimu2_kernels.cuh")
    P = kernel_params(path)
    print(P)
    for ch in range(3):
        print(ch, simulate(P, ch))
