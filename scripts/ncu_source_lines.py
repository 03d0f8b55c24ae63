#!/usr/bin/env python
"""Instructions executed per SOURCE LINE of one kernel, from an `ncu --set full --import-source on` report
(kernels are compiled with -lineinfo).  This is how the per-line breakdowns quoted in DESIGN.md were obtained.

    python scripts/ncu_source_lines.py gpurun_out/prof.ncu-rep nn_fixed_query_rect [top N]

Columns: share of the kernel's warp instructions, share of the stall samples, average active lanes, file:line, source.
"""
import csv, sys, subprocess
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","-k","regex:"+kern],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None; hdr=None; agg=[]
for r in rows:
    if len(r)==2 and r[0] in ("File Path","File Name"): cur=r[1]; continue
    if r and r[0]=="Line No": hdr=r; continue
    if hdr and len(r)>8 and r[0].isdigit():
        ie=hdr.index("Instructions Executed"); sm=hdr.index("# Samples"); ti=hdr.index("Thread Instructions Executed")
        try: agg.append((int(r[ie]), int(r[sm]), cur.split('/')[-1], int(r[0]), r[1][:100], int(r[ti])))
        except Exception as e: pass
tot=sum(a[0] for a in agg); ts=sum(a[1] for a in agg); tt=sum(a[5] for a in agg)
print("total warp inst", tot, "thread inst", tt, "avg active lanes", tt/tot, "samples", ts)
for a in sorted(agg, reverse=True)[:top]:
    print(f"{100*a[0]/tot:5.1f}% inst {100*a[1]/ts:5.1f}% smp lanes {a[5]/max(a[0],1):4.1f} {a[2]}:{a[3]}  {a[4]}")
