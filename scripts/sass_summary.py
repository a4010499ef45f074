"""Instruction-class counts per kernel from the SASS of the built library (cuobjdump -sass), for profiles/sass_summary.txt:
which kernels use bulk copies (UBLKCP) and mbarriers (SYNCS), 256-bit global loads (LDG.256, sm_100), the cluster barrier (UCGABAR / BAR.*CGA), shared-memory
atomics, local memory.     python scripts/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "huffman_b200", "libhuffb200.so")
CLASSES = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("UCGABAR", r"\bUCGABAR|\bCGABAR|BAR\.[A-Z.]*CGA"), ("CCTL", r"\bCCTL"),
           ("ATOMS", r"\bATOMS"), ("ATOMG/RED", r"\bATOMG|\bATOM\b|\bRED\b"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"),
           ("LDG", r"\bLDG"), ("LDG.256", r"\bLDG\.[A-Z0-9.]*256"), ("STG", r"\bSTG"), ("LDL", r"\bLDL"), ("STL", r"\bSTL"), ("BAR", r"\bBAR\."),
           ("SHFL", r"\bSHFL"), ("VOTE", r"\bVOTE"), ("MATCH", r"\bMATCH"), ("FLO/POPC/BREV", r"\bFLO|\bPOPC|\bBREV"),
           ("R2UR/REDUX", r"\bREDUX"), ("UTMA*", r"\bUTMALDG|\bUTMASTG")]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, counts, total = None, {}, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", name).replace("hf::", "").replace("void ", "")
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if kern and m:
        ins = m.group(1)
        total[kern] += 1
        for cname, pat in CLASSES:
            if re.search(pat, ins):
                counts[kern][cname] += 1
print("SASS instruction classes per kernel of huffman_b200/libhuffb200.so (sm_100a), static counts")
print("%-28s %6s " % ("kernel", "instr") + " ".join("%7s" % c[0][:7] for c in CLASSES))
for k in sorted(counts, key=lambda k: -total[k]):
    print("%-28s %6d " % (k[:28], total[k]) + " ".join("%7s" % (counts[k][c[0]] or ".") for c in CLASSES))
tot = collections.Counter()
for k in counts:
    tot.update(counts[k])
print("%-28s %6d " % ("all kernels", sum(total.values())) + " ".join("%7s" % (tot[c[0]] or ".") for c in CLASSES))
