"""Per-kernel SASS opcode summary of libnerfq.so (cuobjdump -sass): the instructions that prove what each kernel runs on
-- tcgen05 MMAs (UTCHMMA), TMEM loads/stores (LDTM/STTM), bulk async copies on the TMA engine (UBLKCP), mbarrier traffic
(SYNCS), tcgen05.commit (UTCBAR), vector global accesses, warp shuffles / REDUX for the scans and reductions.

    python profiles/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vanilla-nerf-model-compression-using-lsa-enhanced-nncodec_b200", "libnerfq.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "FENCE", "MEMBAR", "LDG.E.128", "LDG.E.ENL2.256", "LDG", "STG.E.128",
       "STG.E.ENL2.256", "STG", "LDS", "STS", "SHFL", "REDUX", "ATOMS", "RED", "ATOMG", "MUFU", "F2FP", "HFMA2", "FFMA", "LDL", "STL"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for o in OPS:
            if op == o or op.startswith(o + ".") or (("." in o) and op.startswith(o)):
                counts[kern][o] += 1
print("SASS opcode counts per kernel (sm_100a), libnerfq.so; LDG/STG rows include the vector forms listed separately")
print(f"{'kernel':70s} {'instr':>7s}  " + " ".join(f"{o}" for o in OPS))
for k, c in counts.items():
    cells = " ".join(f"{o}={c[o]}" for o in OPS if c[o])
    print(f"{k[:70]:70s} {total[k]:7d}  {cells}")
