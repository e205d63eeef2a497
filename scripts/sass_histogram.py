"""SASS opcode histogram per kernel of libhsidm_b200.so (cuobjdump -sass): evidence that the hot kernels are hand-written
tcgen05 / TMA / TMEM code (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store, LDTM = tcgen05.ld, UTCBAR =
tcgen05.commit, SYNCS = mbarrier ops).  Usage: python scripts/sass_histogram.py > profiles/r2_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hsi_dmgasr_b200", "libhsidm_b200.so")
KEY = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "FFMA", "LDG", "STG", "LDS", "STS",
       "ATOM", "RED", "MUFU", "BAR", "ACQBULK", "UCGABAR", "R2UR", "R2UR.BROADCAST", "REDUX"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = funcs.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P[T\d]+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        op = m.group(1)
        cur[op.split(".")[0]] += 1
        if op.startswith("UTCHMMA.2CTA"):
            cur["UTCHMMA.2CTA"] += 1
        if op.startswith("R2UR.BROADCAST"):
            cur["R2UR.BROADCAST"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
print("# SASS opcode histogram of libhsidm_b200.so (sm_100a), per kernel\n")
print("`cuobjdump -sass` of the shipped library, counted by `scripts/sass_histogram.py`.  UTCHMMA = `tcgen05.mma`, UTMALDG / UTMASTG = TMA "
      "tensor load / store, LDTM = `tcgen05.ld`, UTCBAR = `tcgen05.commit`, SYNCS = mbarrier operations, FFMA = CUDA-core fp32 FMA.  "
      "R2UR.BROADCAST in front of a UTCHMMA = the MMA issuer fell out of the uniform datapath (DESIGN.md section 4): 0 in every tensor-core kernel.  "
      "No cuBLAS / cuDNN / CUTLASS device code is linked (the library links only the static CUDA runtime).\n")
print("| kernel | instructions | " + " | ".join(KEY) + " |")
print("|---|---|" + "---|" * len(KEY))
tot = collections.Counter()
for (name, c), dm in zip(funcs.items(), demangle):
    short = dm.replace("hsidm::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("hsidm::", "")
    short = re.sub(r"^void ", "", short)
    short = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", short)        # drop the parameter list, keep the template arguments
    n = sum(v for k, v in c.items() if k not in ("UTCHMMA.2CTA", "R2UR.BROADCAST"))
    if n < 50:
        continue
    print(f"| `{short}` | {n} | " + " | ".join(str(c.get(k, 0)) for k in KEY) + " |")
    tot.update(c)
print(f"| **all kernels** | {sum(v for k, v in tot.items() if k not in ('UTCHMMA.2CTA', 'R2UR.BROADCAST'))} | " + " | ".join(str(tot.get(k, 0)) for k in KEY) + " |")
