#!/usr/bin/env python
"""Print a SASS range of one kernel from `ncu --page source --csv` output with executed counts and samples.
    python tools/ncu_dump.py src.csv KERNEL_INDEX FIRST LAST"""
import sys
sys.path.insert(0, __import__("os").path.dirname(__file__))
from ncu_source import load
name, hdr, data = load(sys.argv[1], int(sys.argv[2]))
a, b = int(sys.argv[3]), int(sys.argv[4])
iS, iSamp, iEx = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
for i in range(a, min(b, len(data))):
    r = data[i]
    print(str(i).rjust(5), r[iS].strip()[:70].ljust(70), r[iSamp].rjust(6), r[iEx].rjust(10))
