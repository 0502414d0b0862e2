import csv, sys, collections, os
path=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 60
agg=collections.OrderedDict()
cur=None;hdr=None
for r in csv.reader(open(path)):
    if not r: continue
    if r[0]=="File Path": cur=os.path.basename(r[1]); hdr=None; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; continue
    if hdr is None or not r[0].isdigit(): continue
    d=dict(zip(hdr,r))
    try:
        ins=int(r[hdr.index("Instructions Executed")]); smp=int(r[hdr.index("# Samples")]); thr=int(r[hdr.index("Thread Instructions Executed")])
    except ValueError: continue
    k=(cur,int(r[0]))
    a=agg.setdefault(k,[0,0,0,r[1]]); a[0]+=ins; a[1]+=smp; a[2]+=thr
tot=sum(a[0] for a in agg.values()); tots=sum(a[1] for a in agg.values())
print("total",tot,tots)
mode=sys.argv[3] if len(sys.argv)>3 else "top"
items=list(agg.items())
if mode=="top": items=sorted(items,key=lambda kv:-kv[1][0])[:top]
else: items=[kv for kv in items if kv[0][0]==mode]; items.sort(key=lambda kv: kv[0][1])
for (f,ln),a in items:
    print("%5.2f%% ins %5.2f%% smp lanes %4.1f %s:%d  %s"%(100*a[0]/tot,100*a[1]/max(1,tots),a[2]/max(1,a[0]),f,ln,a[3].strip()[:110]))
