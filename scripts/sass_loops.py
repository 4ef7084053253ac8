"""Static look at the hot loops of a kernel: python scripts/sass_loops.py lib.so [function-substring]
Lists every backward branch (loop) of more than 100 instructions with its length and opcode mix -- the number to
watch before spending GPU time (the fast forward / backward cell loops of k_fb2 are the two longest)."""
import re, subprocess, sys
from collections import Counter
lib = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else "k_fb2ILi4ELb0ELb0"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None; ins = []
funcs = {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m and cur: funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for f, ins in funcs.items():
    if want not in f: continue
    print(f, len(ins), "instructions")
    addr = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr and i - addr[tgt] > 100:
                body = ins[addr[tgt]: i + 1]
                c = Counter(re.sub(r"^@!?U?P\d+\s+", "", x).split()[0].split(".")[0] for _, x in body)
                print("  loop %05x..%05x  %4d instr  " % (tgt, a, len(body)) + " ".join("%s:%d" % kv for kv in c.most_common(14)))
