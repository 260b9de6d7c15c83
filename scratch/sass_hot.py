import csv, sys, collections, re
path=sys.argv[1]; which=int(sys.argv[2]) if len(sys.argv)>2 else 0
rows=list(csv.reader(open(path)))
starts=[i for i,r in enumerate(rows) if r and r[0]=='Kernel Name']
seg=rows[starts[which]:(starts[which+1] if which+1<len(starts) else len(rows))]
print(seg[0][1], 'instances', len(starts))
hdr=seg[1]; ci=hdr.index('Instructions Executed'); cs=hdr.index('# Samples')
ins=[]
for r in seg[2:]:
    try: ins.append((r[0], r[1].strip(), int(r[ci]), int(r[cs])))
    except: pass
tot=sum(x[2] for x in ins); stot=sum(x[3] for x in ins)
print('static',len(ins),'executed',tot,'samples',stot)
by=collections.Counter(); bys=collections.Counter()
for a,s,n,sm in ins:
    m=re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_\.]+)', s); op=m.group(2).split('.')[0] if m else s
    by[op]+=n; bys[op]+=sm
for op,n in by.most_common(40): print(f'{op:12s} {n:10d} {100*n/tot:5.1f}%  samples {100*bys[op]/max(stot,1):5.1f}%')
if len(sys.argv)>3:
    for i,(a,s,n,sm) in enumerate(ins): print(f'{i:5d} {n:8d} {sm:5d}  {s}')
