"""Summarise the source page of an ncu report: python scripts/ncu_src.py rep.ncu-rep [top]
Groups SASS instructions into contiguous hot regions by executed count, prints opcode mix and stalls."""
import csv, subprocess, sys, re
from collections import Counter, defaultdict
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    ins.append(r)
tot = sum(int(r[ix["Instructions Executed"]]) for r in ins)
samples = sum(int(r[ix["# Samples"]]) for r in ins)
print("instructions", len(ins), "executed warp-instr", tot, "samples", samples)
def opc(src):
    s = re.sub(r"^@!?U?P\d+\s+", "", src.strip())
    return s.split()[0].split(".")[0]
c = Counter(); cs = Counter()
for r in ins:
    c[opc(r[ix["Source"]])] += int(r[ix["Instructions Executed"]])
    cs[opc(r[ix["Source"]])] += int(r[ix["# Samples"]])
print("opcode mix (% of executed | % of samples):")
for k, v in c.most_common(24):
    print("  %-8s %5.1f%%  %5.1f%%" % (k, 100.0 * v / tot, 100.0 * cs[k] / max(1, samples)))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
sc = Counter()
for r in ins:
    for h in stalls:
        sc[h] += int(r[ix[h]] or 0)
print("stall samples:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(1, samples)) for k, v in sc.most_common(10)))
# hot regions: split where executed count changes by > 30%
print("regions (start idx, n instr, executed per instr, share of executed, share of samples):")
i = 0
regs = []
while i < len(ins):
    e = int(ins[i][ix["Instructions Executed"]]); j = i
    s = 0; se = 0
    while j < len(ins) and abs(int(ins[j][ix["Instructions Executed"]]) - e) <= 0.3 * max(e, 1):
        s += int(ins[j][ix["# Samples"]]); se += int(ins[j][ix["Instructions Executed"]]); j += 1
    regs.append((i, j - i, e, se, s)); i = j
for (i, n, e, se, s) in sorted(regs, key=lambda t: -t[3])[: int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print("  @%4d n=%4d exec/instr=%.3e  %5.1f%% exec  %5.1f%% samples   first: %s" % (i, n, e, 100.0 * se / tot, 100.0 * s / max(1, samples), ins[i][ix["Source"]].strip()[:60]))
