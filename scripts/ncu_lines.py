import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if 'Warp Stall Sampling (All Samples)' in r][0]
hdr=rows[hi]
stallcols=[(i,h) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
isamp=hdr.index('Warp Stall Sampling (All Samples)'); iex=hdr.index('Instructions Executed')
agg={}
for r in rows[hi+1:]:
    if len(r)<len(hdr) or not r[0].strip(): continue
    try: s=int(r[isamp]); e=int(r[iex])
    except: continue
    a=agg.setdefault(r[0],[0,0,r[1][:70],collections.Counter()])
    a[0]+=s; a[1]+=e
    for i,h in stallcols:
        try: a[3][h]+=int(r[i])
        except: pass
tot=sum(v[0] for v in agg.values())
allr=collections.Counter()
for v in agg.values(): allr.update(v[3])
print("total samples",tot, [(k.replace('stall_',''),round(c/tot,3)) for k,c in allr.most_common(8)])
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[2]) if len(sys.argv)>2 else 25]:
    print(f"{k:>5s} {v[0]/tot:5.3f} inst={v[1]:9d} {v[2]:70s} {[(a.replace('stall_',''),b) for a,b in v[3].most_common(3)]}")
