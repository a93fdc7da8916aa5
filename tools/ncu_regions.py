#!/usr/bin/env python
"""Group the SASS of one kernel (ncu --page source --csv) into regions of similar execution count and print, per region,
the executed warp-instructions split into FP64-pipe and other instructions.
    python tools/ncu_regions.py src.csv KERNEL_INDEX [UNITS]   (UNITS: divide the counts, e.g. warp-steps)"""
import sys
sys.path.insert(0, __import__("os").path.dirname(__file__))
from ncu_source import load
name, hdr, data = load(sys.argv[1], int(sys.argv[2]))
units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
iS, iSamp, iEx = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
def op(r):
    t = r[iS].strip().split()
    o = t[1] if t[0].startswith("@") else t[0]
    return o.split(".")[0]
FP64 = {"DADD", "DMUL", "DFMA", "DSETP", "DMMA", "I2F", "F2F"}
regions = []
cur = None
for i, r in enumerate(data):
    ex = int(r[iEx])
    if cur is None or not (0.7 * cur["ref"] <= ex <= 1.4 * cur["ref"]) and ex > 0 or (ex == 0 and cur["ref"] > 0 and False):
        if cur is None or ex > 0:
            cur = {"a": i, "b": i, "ref": max(ex, 1), "fp64": 0, "other": 0, "samples": 0, "n": 0}
            regions.append(cur)
    cur["b"] = i
    cur["n"] += 1
    cur["samples"] += int(r[iSamp])
    if op(r) in FP64: cur["fp64"] += ex
    else: cur["other"] += ex
# merge tiny regions into neighbours for display
tot = sum(x["fp64"] + x["other"] for x in regions)
print(name[1][:90], "total executed", tot, "per unit", tot / units)
for x in regions:
    t = x["fp64"] + x["other"]
    if t < 0.004 * tot: continue
    print(f"[{x['a']:5d},{x['b']:5d}] n={x['n']:4d} exec/instr~{x['ref']:>11d}  fp64 {x['fp64']/units:9.2f}  other {x['other']/units:9.2f}  samples {x['samples']:8d}")
