import re, sys
rows=[]
tot=ti=0
for ln in open(sys.argv[1]):
    m=re.match(r'total samples (\d+), warp instructions (\d+)',ln)
    if m: tot,ti=int(m.group(1)),int(m.group(2))
    m=re.match(r'(\S+):\s*(\d+)\s+samples\s+(\d+).*inst\s+(\d+)',ln)
    if m: rows.append((m.group(1),int(m.group(2)),int(m.group(3)),int(m.group(4))))
def rng(f,a,b):
    s=sum(r[2] for r in rows if r[0]==f and a<=r[1]<=b); i=sum(r[3] for r in rows if r[0]==f and a<=r[1]<=b); return s,i
div=float(sys.argv[2]) if len(sys.argv)>2 else 148*89.7
for spec in sys.argv[3:]:
    name,f,a,b=spec.split(',')
    s,i=rng(f,int(a),int(b)); print(f'{name:14s} samples {100*s/tot:5.1f}%  inst {100*i/ti:5.1f}%  inst/iter/CTA {i/div:8.0f}')
