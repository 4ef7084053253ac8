"""Instruction mix of the main loop of a probe kernel variant: python sasscount.py <binary> <pattern>"""
import re, subprocess, sys
from collections import Counter
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
f = False; L = []
for ln in out.splitlines():
    if "Function :" in ln:
        f = sys.argv[2] in ln
    elif f:
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", ln)
        if m: L.append(m.group(1).strip())
br = [i for i, l in enumerate(L) if re.search(r"\bBRA\b", l) and (l.startswith("@") or "BRA.U" in l)]
end = br[-1]
# loop start = branch target: approximate with the first fp64 op after the previous branch
prev = [i for i in br if i < end]
start = prev[-1] + 1 if prev else 0
body = L[start:end + 1]
div = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
c = Counter(re.sub(r"^@!?U?P\d+\s+", "", l).split()[0].split(".")[0] for l in body)
print("%d instrs (%.1f per unit): %s" % (len(body), len(body) / div, ", ".join("%s %.1f" % (k, v / div) for k, v in c.most_common(16))))
if len(sys.argv) > 4:
    print("\n".join(body))
