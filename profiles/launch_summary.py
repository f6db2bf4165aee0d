import csv, collections, re, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; data=rows[hdr+1:]
ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
tot=collections.Counter(); cnt=collections.Counter()
for r in data:
    if len(r)<=vi: continue
    name=re.sub(r'\(.*','',r[ki]); name=re.sub(r'^void ','',name)[:100]
    v=float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1e3
    elif r[ui]=='ms': v*=1e3
    tot[name]+=v; cnt[name]+=1
T=sum(tot.values())
print("total us",round(T,1), "launches",sum(cnt.values()))
for k,v in tot.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 25): print(f"{v:9.1f} us {100*v/T:5.1f}% n={cnt[k]:4d} {k}")
