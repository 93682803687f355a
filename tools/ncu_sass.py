"""Dump SASS with executed counts / samples from an .ncu-rep: python tools/ncu_sass.py rep [min_count]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
iA, iS, iE, iN = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
for r in rows[hi + 1:]:
    if len(r) <= iE or not r[iA].startswith("0x"):
        continue
    a = int(r[iA], 16)
    base = base or a
    print(f"{a - base:05x} {int(r[iE]):10d} {int(r[iN]):6d}  {r[iS].strip()}")
