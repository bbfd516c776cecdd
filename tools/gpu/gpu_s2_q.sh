#!/bin/bash
mkdir -p gpurun_out
python tools/query_wall.py 10000 10 30 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_query_v2.csv python tools/query_wall.py 10000 10 1 > gpurun_out/ncu_qwall.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_query_v2.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
ks=[(int(r[ii]), r[ki], float(r[vi].replace(',',''))) for r in rows[1:]]
# the last call = the kernels after the last row_prep of the query batch: take the last 60 launches
tail=ks[-70:]
tot=0
for i,n,v in tail:
    print(i, n[:70], v)
PY
