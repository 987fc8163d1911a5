"""Per-region instruction and stall-sample breakdown of one kernel from an ncu report with --import-source on (SASS page): consecutive
SASS lines executed the same number of times per tile are one region.

    python profiles/region_breakdown.py gpurun_out/f1s_c3_8gib_r1b.ncu-rep <bytes scanned> [threshold]
"""
import csv,sys,subprocess
rep=sys.argv[1]; nbytes=float(sys.argv[2]); tilebytes=2048
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
# find header
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
hdr=rows[hi]; data=[r for r in rows[hi+1:] if len(r)==len(hdr)]
ia=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples')
tiles=nbytes/tilebytes
tot=sum(int(r[ia]) for r in data); stot=sum(int(r[isamp]) for r in data)
print('kernel',rows[0][1][:80]); print('total inst/tile',tot/tiles,'per 32B',tot/tiles/64,'samples',stot)
thr=float(sys.argv[3]) if len(sys.argv)>3 else 0.02
prev=None;start=0;acc=0;sacc=0
lines=[]
for i,r in enumerate(data):
    c=int(r[ia])/tiles; sm=int(r[isamp])
    key=c
    if prev is None: prev=key
    if abs(key-prev)>0.05*max(prev,1e-9)+0.01:
        lines.append((start,i-1,prev,acc,sacc)); start=i;acc=0;sacc=0;prev=key
    acc+=c;sacc+=sm
lines.append((start,len(data)-1,prev,acc,sacc))
for a,b,e,ac,sa in lines:
    if ac>1.0 or sa>stot*0.003:
        print(f"lines {a}-{b}: n={b-a+1} exec/tile={e:.3f} instr/tile={ac:.1f} samples={sa} ({100*sa/stot:.1f}%)  first: {data[a][1].strip()[:60]}")
