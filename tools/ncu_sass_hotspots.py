import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
hdr=rows[hi]; data=[r for r in rows[hi+1:] if len(r)>10]
ci={h:i for i,h in enumerate(hdr)}
tot_inst=sum(int(r[ci['Instructions Executed']]) for r in data)
tot_samp=sum(int(r[ci['# Samples']]) for r in data)
tot_thr=sum(int(r[ci['Thread Instructions Executed']]) for r in data)
print("SASS instrs",len(data),"warp-inst %.1fM"%(tot_inst/1e6),"avg threads %.2f"%(tot_thr/tot_inst),"samples",tot_samp)
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg={s:0 for s in stalls}
for r in data:
    for s in stalls: agg[s]+=int(r[ci[s]])
print("stall totals:", {k:round(100*v/tot_samp,1) for k,v in sorted(agg.items(),key=lambda kv:-kv[1])[:8]})
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.006
for k,r in enumerate(data):
    ie=int(r[ci['Instructions Executed']]); at=r[ci['Avg. Threads Executed']]; smp=int(r[ci['# Samples']])
    src=r[ci['Source']].strip()[:60]
    st=sorted(((int(r[ci[s]]),s) for s in stalls),reverse=True)[:2]
    if smp>=tot_samp*thr or any(x in src for x in ('BRA','LDG','LDL','STL','BSSY','EXIT','LDS','STS')):
        print(f"{k:4d} {src:60s} exec={ie/1e6:7.1f}M thr={at:>5s} samp={100*smp/tot_samp:5.2f}% {st[0][1]}={st[0][0]}")
