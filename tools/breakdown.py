import csv, collections, re, sys
f=sys.argv[1]
lines=[l for l in open(f) if not l.startswith('==')]
rows=[]
for x in csv.DictReader(lines):
    if x.get('Metric Name')=='gpu__time_duration.sum':
        rows.append((x['Kernel Name'], float(x['Metric Value'].replace(',','')), x['Grid Size']))
idx=[i for i,r_ in enumerate(rows) if 'adamw_kernel' in r_[0]]
a,b=idx[-2]+1, idx[-1]+1
step=rows[a:b]
tot=sum(r_[1] for r_ in step)
print("kernels in last step:", len(step), "sum(us)=%.1f"%(tot/1000))
agg=collections.OrderedDict()
for n,t,g in step:
    k=re.sub(r'\(.*','',n).replace('void ','').replace('mmssl::','')[:60]
    agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=t
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{t/1000:9.1f} us  {100*t/tot:5.1f}%  x{c:3d}  {k}")
if len(sys.argv)>2:
    for n,t,g in step: print("%8.1f %-16s %s"%(t/1000,g,re.sub(r'\(.*','',n).replace('void ','').replace('mmssl::','')[:70]))
