"""SASS evidence for the final build (no GPU needed): per kernel of libmvsb200.so, registers and the counts of the
mnemonics DESIGN.md's claims rest on -- UTCHMMA (tcgen05.mma), UTMALDG (TMA loads), UTCBAR / SYNCS (mbarriers), LDTM
(tcgen05.ld), FFMA2 (packed fp32), 256-bit global accesses (.ENL2.256), LDGSTS (cp.async), ACQBULK / PDL (griddepcontrol).

    python profiles/sass_mnemonics.py [libmvsb200.so] > profiles/sass_mnemonics_<round>.md
"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "wild_deep_mvs_b200", "libmvsb200.so")
KEYS = ["UTCHMMA", "UTMALDG", "UTCBAR", "SYNCS", "LDTM", "FFMA2", "FFMA", "HFMA2", "256", "LDGSTS", "LDS", "STS", "LDG", "STG", "RED", "ATOM",
        "BAR", "ACQBULK", "PREEXIT"]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True, check=True).stdout
regs = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+).*?SHARED:(\d+)", res):
    regs[m.group(1)] = (int(m.group(2)), int(m.group(3)))
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        op = m.group(1)
        base = op.split(".")[0]
        counts[cur][base] += 1
        if ".256" in op:
            counts[cur]["256"] += 1
        if base == "FFMA2":
            counts[cur]["FFMA"] -= 0
demangle = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print("# SASS mnemonic counts per kernel (`cuobjdump -sass`, final build)\n")
print("| kernel | regs | static smem | " + " | ".join(KEYS) + " |")
print("|---|---|---|" + "---|" * len(KEYS))
for (name, c), pretty in zip(counts.items(), demangle):
    pretty = re.sub(r"\(.*", "", pretty).replace("void ", "").replace("mvsb200::", "")
    r = regs.get(name, ("?", "?"))
    print("| `%s` | %s | %s | " % (pretty, r[0], r[1]) + " | ".join(str(c.get(k, 0)) if c.get(k, 0) else "" for k in KEYS) + " |")
