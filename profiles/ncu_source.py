import csv, subprocess, sys
rep=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 25
raw=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
h=rows[1]; data=rows[2:]
ws=h.index('Warp Stall Sampling (All Samples)'); si=h.index('Source'); ie=h.index('Instructions Executed')
tot=sum(int(r[ws]) for r in data if r[ws].isdigit())
print('total samples',tot,'instructions',len(data))
print("-- sequential listing with samples (only lines with >=1% or memory ops)")
for idx,r in enumerate(data):
    s=int(r[ws]) if r[ws].isdigit() else 0
    src=r[si].strip()
    if s>=0.01*tot or any(k in src for k in ('LDG','STG','SHFL','BRA','LDL','STL')):
        print(f"{idx:4d} {s:6d} {100*s/tot:5.1f}% ex={r[ie]:>7} {src[:100]}")
