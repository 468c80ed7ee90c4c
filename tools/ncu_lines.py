"""Summarise an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
his=[i for i,r in enumerate(rows) if r and r[0]=="Line No"]
data={}
tot_inst=0; tot_s=0
for hn,hi in enumerate(his):
    hdr=rows[hi]
    fpath=rows[hi-2][1] if hi>=2 else "?"
    def col(name): return [i for i,h in enumerate(hdr) if h==name][0]
    ci=col("Instructions Executed"); csm=col("# Samples")
    stalls=[(h,i) for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    end=his[hn+1]-2 if hn+1<len(his) else len(rows)
    for r in rows[hi+1:end]:
        if not r or not r[0].isdigit(): continue
        num=lambda v: int(v) if v.strip().lstrip("-").isdigit() else 0
        inst=num(r[ci]); s=num(r[csm])
        key=(fpath.split('/')[-1],int(r[0]))
        d=data.setdefault(key,[r[1][:80],0,0,{}])
        d[1]+=inst; d[2]+=s
        for h,i in stalls: d[3][h]=d[3].get(h,0)+num(r[i])
        tot_inst+=inst; tot_s+=s
print("total inst",tot_inst,"samples",tot_s)
N=int(sys.argv[2]) if len(sys.argv)>2 else 25
print("--- top by instructions")
for k,d in sorted(data.items(),key=lambda kv:-kv[1][1])[:N]:
    print(f"{k[0][:18]:18s} L{k[1]:4d} inst={d[1]/tot_inst*100:5.1f}% samp={d[2]/max(tot_s,1)*100:5.1f}%  {d[0]}")
print("--- top by samples")
for k,d in sorted(data.items(),key=lambda kv:-kv[1][2])[:N]:
    top=sorted(d[3].items(),key=lambda kv:-kv[1])[:3]
    print(f"{k[0][:18]:18s} L{k[1]:4d} samp={d[2]/max(tot_s,1)*100:5.1f}% inst={d[1]/tot_inst*100:5.1f}% {d[0][:46]:46s} {[(a[6:],b) for a,b in top]}")
