"""Static look at a kernel's SASS: total instructions, and for every backward branch the size of the loop body and its
opcode mix.  usage: python tools/sass_loops.py <lib.so> <kernel-name-substring> [min_body]"""
import collections
import re
import subprocess
import sys

so, sel = sys.argv[1], sys.argv[2]
min_body = int(sys.argv[3]) if len(sys.argv) > 3 else 50
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, funcs = None, {}
for ln in txt.split("\n"):
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if sel not in name:
        continue
    print(f"== {name}: {len(ins)} instructions")
    addr_idx = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.\S+)?\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr_idx and i - addr_idx[tgt] >= min_body:
                body = ins[addr_idx[tgt]:i + 1]
                ops = collections.Counter()
                for _, tt in body:
                    parts = tt.split()
                    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
                    ops[op.split(".")[0]] += 1
                print(f"  loop {tgt:#x}..{a:#x}: {len(body)} instr; " + " ".join(f"{k}:{v}" for k, v in ops.most_common(28)))
