"""Development aid: aggregates the per-instruction counters of an `ncu --set full --import-source on` report
(ncu -i rep --page source --csv --print-source sass > file.csv) per loop: share of executed warp instructions and
of stall samples, executions per warp.  usage: python scripts/ncu_source_agg.py file.csv [warps in the launch]"""
import csv,re,sys
rows=list(csv.reader(open(sys.argv[1])))
nw=float(sys.argv[2]) if len(sys.argv)>2 else 36864
# sections
secs=[];cur=None
for r in rows:
    if r and r[0]=='Address': h=r; cur=[]; secs.append(cur); continue
    if cur is not None and len(r)==len(h) and r[0].startswith('0x'): cur.append(dict(zip(h,r)))
data=secs[0]
print('sections',len(secs),'instrs',len(data))
base=int(data[0]['Address'],16)
ex=[int(d['Instructions Executed']) for d in data]
smp=[int(d['# Samples']) for d in data]
src=[d['Source'].strip() for d in data]
tot=sum(ex); ts=sum(smp)
print('total warp inst',tot,'samples',ts)
loops=[]
for i,s in enumerate(src):
    if 'BRA' in s:
        m=re.search(r'0x([0-9a-f]+)',s)
        if m:
            v=int(m.group(1),16)
            t=(v-base)//16 if v>=base else v//16
            if 0<=t<i and i-t>30: loops.append((t,i))
for (a,b) in loops:
    e=sum(ex[a:b+1]); s=sum(smp[a:b+1])
    print(f"loop {a}-{b}: inst {e/tot*100:.1f}%  samples {s/ts*100:.1f}%  per-warp {e/nw:.0f}")
# coarse profile in chunks of 100 instrs
print('chunks: start inst% samp%')
for a in range(0,len(data),100):
    e=sum(ex[a:a+100]); s=sum(smp[a:a+100])
    if e/tot>0.004 or s/ts>0.004: print(a, f"{e/tot*100:.1f} {s/ts*100:.1f}")
import pickle
pickle.dump((src,ex,smp),open(sys.argv[1]+'.pkl','wb'))
