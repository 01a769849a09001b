import csv,re,sys
from collections import defaultdict
def load(path):
    lines=[l for l in open(path) if l.startswith('"')]
    rows=[r for r in csv.DictReader(lines) if r['Metric Name']=='gpu__time_duration.sum']
    def ns(r):
        v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
        return v*{'ns':1,'us':1e3,'ms':1e6}.get(u,1)
    names=[re.sub(r'\(anonymous namespace\)::|<unnamed>::|void ','',r['Kernel Name']).split('(')[0] for r in rows]
    idx=[i for i,n in enumerate(names) if n.startswith('dropout_mask')]
    a=idx[-1]
    return [(names[i],ns(rows[i])/1e3,rows[i]['Grid Size']) for i in range(a,len(rows))]
for path in sys.argv[1:]:
    st=load(path)
    print(path,len(st),'launches, sum us',round(sum(t for _,t,_ in st),1))
    agg=defaultdict(lambda:[0,0.0])
    for n,t,g in st: agg[n][0]+=1; agg[n][1]+=t
    for n,(c,t) in sorted(agg.items(),key=lambda kv:-kv[1][1])[:22]: print(f'  {t:8.1f} us {c:4d} {t/c:6.1f}  {n[:80]}')
