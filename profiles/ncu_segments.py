#!/usr/bin/env python
"""Split an `ncu --page source --csv` SASS dump at barrier instructions and report, per segment
(= per engine pass), sampled time share, executed instructions by class and the top stall reasons."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
segs = []
cur = dict(start=None, samples=0, inst=0, cls=collections.Counter(), stalls=collections.Counter(), first=None, n=0, wf=0, wf_ideal=0)
def cls_of(op):
    if op.startswith(("FFMA","FADD","FMUL","FSEL","FSETP","FMNMX","HFMA2")): return "fp"
    if op.startswith(("LDS","STS")): return "smem"
    if op.startswith(("LDG","STG","LD.","ST.","LD ","LDC","LDCU")) or op in ("LD","ST"): return "gmem/generic"
    if op.startswith(("BAR","UCGABAR","MEMBAR","ERRBAR","CCTL","FENCE")): return "sync"
    if op.startswith(("DADD","DFMA","DMUL","DSETP","F2F","I2F")): return "fp64/cvt"
    if op.startswith(("SHFL",)): return "shfl"
    return "int/other"
total = 0
for r in rows[2:]:
    src = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src.split()[0]
    smp = int(r[ix["# Samples"]] or 0); ins = int(r[ix["Instructions Executed"]] or 0)
    total += smp
    cur["samples"] += smp; cur["inst"] += ins; cur["cls"][cls_of(op)] += ins; cur["n"] += 1
    cur["wf"] += int(r[ix["L1 Wavefronts Shared"]] or 0); cur["wf_ideal"] += int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
    for s in stall_cols:
        cur["stalls"][s] += int(r[ix[s]] or 0)
    if op.startswith(("BAR", "UCGABAR_WAIT")):
        cur["end"] = op
        segs.append(cur)
        cur = dict(start=op, samples=0, inst=0, cls=collections.Counter(), stalls=collections.Counter(), first=None, n=0, wf=0, wf_ideal=0)
segs.append(cur)
print(f"total samples {total}")
for i, s in enumerate(segs):
    if s["samples"] < total * 0.004: continue
    top = ", ".join(f"{k.replace('stall_','')}:{v*100//max(s['samples'],1)}%" for k, v in s["stalls"].most_common(5))
    cl = ", ".join(f"{k}:{v/1e6:.1f}M" for k, v in s["cls"].most_common())
    print(f"seg {i:2d} sass_instrs={s['n']:5d} time={100*s['samples']/total:5.1f}%  executed={s['inst']/1e6:8.1f}M  [{cl}]  smem_wf={s['wf']/1e6:.0f}M ideal={s['wf_ideal']/1e6:.0f}M\n        stalls: {top}  (ends at {s.get('end')})")
