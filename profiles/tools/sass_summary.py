"""cuobjdump -sass of libintel_b200.so -> per-kernel counts of the mnemonics that tell a Blackwell-native kernel from a
recompiled one (B200_PROFILING.md): UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), HMMA
(mma.sync), UTMA* / UBLKCP (TMA), LDGSTS (cp.async).   python profiles/tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "intel_sigir2023_b200", "libintel_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = {"UTCHMMA": r"\bUTC[A-Z]*MMA", "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "UTCBAR": r"\bUTCBAR", "HMMA": r"\bHMMA", "TMA": r"\bUTMA|\bUBLKCP",
       "LDGSTS": r"\bLDGSTS", "FFMA": r"\bFFMA", "MUFU": r"\bMUFU"}
cur, rows = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        rows[cur] = collections.Counter()
        continue
    if cur and re.search(r"/\*[0-9a-f]{4,6}\*/", line):
        rows[cur]["instr"] += 1
        for k, p in pat.items():
            if re.search(p, line):
                rows[cur][k] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(rows), capture_output=True, text=True).stdout.splitlines()
print(f"{'instr':>7} {'UTCHMMA':>8} {'LDTM':>5} {'STTM':>5} {'UTCBAR':>6} {'HMMA':>6} {'TMA':>4} {'LDGSTS':>6}  kernel")
tot = collections.Counter()
for (k, c), name in zip(rows.items(), demangle):
    tot.update(c)
    name = re.sub(r"\(.*", "", name)[:100]
    print(f"{c['instr']:7d} {c['UTCHMMA']:8d} {c['LDTM']:5d} {c['STTM']:5d} {c['UTCBAR']:6d} {c['HMMA']:6d} {c['TMA']:4d} {c['LDGSTS']:6d}  {name}")
print(f"{tot['instr']:7d} {tot['UTCHMMA']:8d} {tot['LDTM']:5d} {tot['STTM']:5d} {tot['UTCBAR']:6d} {tot['HMMA']:6d} {tot['TMA']:4d} {tot['LDGSTS']:6d}  TOTAL")
