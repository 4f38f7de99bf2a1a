mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "hub or bit_exact or oracle" 2>&1 | tail -3 )
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/probe.csv python scripts/hub_probe.py > gpurun_out/probe.log 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/probe.csv') if not l.startswith('==')]
agg=collections.OrderedDict()
k=0
for row in csv.DictReader(lines):
    n=row['Kernel Name']
    for key in ('walk_hub2','walk_small','snapshot'):
        if key in n:
            if key=='snapshot': k+=1
            agg.setdefault((k-1)//3, {}).setdefault(key, []).append(float(row['Metric Value'].replace(',',''))/1000)
for c,d in agg.items():
    print('case', 'ABCD'[c], {kk: [round(x,1) for x in v] for kk,v in d.items()})
PY
