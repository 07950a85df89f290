"""Counts the Blackwell-specific SASS mnemonics per kernel of libbp_b200.so (cuobjdump -sass): tcgen05.mma (UTCHMMA),
tensor-memory loads / stores (LDTM / STTM), bulk async copies (UBLKCP), tensor-map TMA (UTMALDG), packed FP32 FMA (FFMA2),
mbarrier try-waits (SYNCS.PHASECHK) — the evidence that the hot kernels run on the 5th-gen tensor cores."""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(__file__).resolve().parents[1] / "basic_pitch_b200" / "libbp_b200.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
keys = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTCBAR", "FFMA2", "SYNCS.PHASECHK", "R2UR"]
cur, counts, size = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter()
        size[cur] = 0
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        size[cur] += 1
        for k in keys:
            if k in line:
                counts[cur][k] += 1
print(f"{'kernel':64s} {'instrs':>7s} " + " ".join(f"{k:>9s}" for k in keys))
for k, c in counts.items():
    print(f"{k[:64]:64s} {size[k]:7d} " + " ".join(f"{c[x]:9d}" for x in keys))
print("\ntotals: " + ", ".join(f"{x} {sum(c[x] for c in counts.values())}" for x in keys))
