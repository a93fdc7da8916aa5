#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output of one kernel: stall samples by SASS region.

    ncu -i rep.ncu-rep --page source --csv --kernel-id :::N > src.csv
    python tools/ncu_source.py src.csv [top]
"""
import csv
import sys


def load(path, which=0):
    rows = [r for r in csv.reader(open(path)) if r]
    starts = [i for i, r in enumerate(rows) if r[0] == "Kernel Name"] + [len(rows)]
    a, b = starts[which], starts[which + 1]
    hdr = rows[a + 1]
    data = [r for r in rows[a + 2:b] if len(r) == len(hdr)]
    return rows[a], hdr, data


def main():
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    name, hdr, data = load(sys.argv[1], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    iS, iSamp, iEx = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[iSamp]) for r in data)
    totex = sum(int(r[iEx]) for r in data)
    print(name[1][:100], "samples", tot, "instr", len(data), "executed", totex)
    agg = {}
    for r in data:
        for c, h in stall_cols:
            agg[h[6:]] = agg.get(h[6:], 0) + int(r[c])
    print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:10])
    idx = sorted(range(len(data)), key=lambda i: -int(data[i][iSamp]))[:top]
    for i in sorted(idx):
        r = data[i]
        st = sorted([(int(r[c]), h[6:]) for c, h in stall_cols if int(r[c]) > 0], reverse=True)[:3]
        print(str(i).rjust(5), r[iS].strip()[:52].ljust(52), r[iSamp].rjust(6), r[iEx].rjust(10), st)
    # executed-count histogram by opcode class
    cls = {}
    for r in data:
        op = r[iS].strip().split()[0] if not r[iS].strip().startswith("@") else r[iS].strip().split()[1]
        op = op.split(".")[0]
        cls[op] = cls.get(op, 0) + int(r[iEx])
    print("executed by opcode:", sorted(((v, k) for k, v in cls.items()), reverse=True)[:25])


if __name__ == "__main__":
    main()
