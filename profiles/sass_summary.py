"""Opcode histogram per kernel of liben234gpu.so (cuobjdump -sass) -> profiles/sass_summary.txt.
What it shows: one sm_100a cubin; FP64 arithmetic (DFMA / DADD / DMUL) everywhere, no tensor-core (UTC*MMA, HMMA) and no TMA
(UTMALDG / UBLKCP) instructions — the path is FP64 and HBM-bound, so there is nothing for them to do;
peer-memory collectives are ordinary LDG / STG on mapped peer pointers plus MEMBAR.SC.SYS / fences."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "en234_fea_b200", "csrc", "liben234gpu.so")
out = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
kern, hist, arch = None, collections.OrderedDict(), set()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("en234k::", "").replace("void ", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
WATCH = ["DFMA", "DADD", "DMUL", "LDG", "STG", "LDS", "STS", "SHFL", "ATOMG", "RED", "MEMBAR", "BAR", "LDL", "STL",
         "HMMA", "UTCHMMA", "UTMALDG", "UBLKCP", "LDGSTS"]
with open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w") as f:
    f.write("cuobjdump -sass en234_fea_b200/csrc/liben234gpu.so: architectures %s, %d kernels\n" % (sorted(arch), len(hist)))
    f.write("%-58s %6s " % ("kernel", "instr") + " ".join("%7s" % w for w in WATCH) + "\n")
    for k, h in hist.items():
        f.write("%-58s %6d " % (k[:58], sum(h.values())) + " ".join("%7d" % sum(v for o, v in h.items() if o.startswith(w)) for w in WATCH) + "\n")
    tot = collections.Counter()
    for h in hist.values():
        tot.update(h)
    f.write("\nall kernels: " + ", ".join("%s %d" % (o, c) for o, c in tot.most_common(40)) + "\n")
    tc = [o for o in tot if o.startswith(("HMMA", "UTC", "UTMA", "UBLKCP", "DMMA", "IMMA", "QMMA"))]
    f.write("tensor-core / TMA opcodes present: %s\n" % (tc if tc else "none"))
print(open(os.path.join(ROOT, "profiles", "sass_summary.txt")).read()[:3000])
