# per-kernel durations of full-size eager steps (ncu launch list; cold-cache, serialised launches)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/kt.csv python tools/profile_step.py 40 ${1:-0} > /dev/null 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/kt.csv') if not l.startswith('==')]
agg=collections.defaultdict(list)
for x in csv.DictReader(lines):
    try: v=float(x['Metric Value'].replace(',',''))
    except: continue
    u=x['Metric Unit']; v = v/1e6 if u in ('nsecond','ns') else v/1e3 if u in ('usecond','us') else v
    agg[x['Kernel Name'][:44]].append(v)
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:4]:
    v2=v[5:] if len(v)>10 else v
    print(f"{k:46s} n={len(v):3d} mean(after 5)={sum(v2)/len(v2):8.4f} ms")
PY
