"""Per-CUDA-source-line totals (stall samples, warp instructions) from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.  usage: ncu_srclines.py file.csv [min_frac]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
minf = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
agg = collections.OrderedDict()
hdr = None
for r in rows:
    if r and r[0] == 'Line No' and 'Address' in r:
        hdr = r
        isamp, iex = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == '':
        continue
    try:
        s, e = int(r[isamp]), int(r[iex])
    except ValueError:
        continue
    a = agg.setdefault(int(r[0]), [0, 0, r[1][:120]])
    a[0] += s
    a[1] += e
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
print("total samples", tot, "warp instructions", toti)
for k in sorted(agg):
    v = agg[k]
    if v[0] > tot * minf or v[1] > toti * minf:
        print(f"{k:5d} samp={v[0] / tot:6.3f} inst={v[1] / toti:6.3f}  {v[2]}")
