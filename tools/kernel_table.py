import json, sys
d=json.load(open(sys.argv[1]))
n=int(sys.argv[2]) if len(sys.argv)>2 else 30
filt=sys.argv[3] if len(sys.argv)>3 else ''
print('per call eager ms', round(d['per_call_ms_eager'],2))
print(f"{'kernel':18s} {'shape':38s} {'n':>3s} {'tot_ms':>8s} {'share':>6s} {'avg_us':>9s} {'GB/s':>8s} {'TF/s':>7s}")
for k in [k for k in d['kernels'] if filt in k['kernel']][:n]:
    print(f"{k['kernel']:18s} {k['shape']:38s} {k['launches']:3d} {k['total_ms']:8.2f} {k['share']:6.3f} {k['avg_us']:9.1f} {str(k['GBps']):>8s} {str(k['TFLOPs']):>7s}")
agg={}
for k in d['kernels']:
    agg[k['kernel']]=agg.get(k['kernel'],0)+k['total_ms']
for k,v in sorted(agg.items(), key=lambda kv:-kv[1]): print(f"{k:20s} {v:8.2f} ms  {v/d['per_call_ms_eager']:.3f}")
