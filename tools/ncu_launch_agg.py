import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hdr]; ni=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[hdr+1:]:
    if len(r)<=vi: continue
    v=float(r[vi].replace(",",""))
    if r[ui]=="ns": v/=1e3
    elif r[ui]=="ms": v*=1e3
    k=r[ni].split("(")[0][:40]; agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
print("total us",round(tot),"launches",sum(v[0] for v in agg.values()))
for k,v in sorted(agg.items(), key=lambda x:-x[1][1])[:14]: print(f"{k:42s} n={v[0]:5d} us={v[1]:10.1f} avg={v[1]/v[0]:7.1f}")
