"""Summarise an ncu --csv launch list (gpu__time_duration etc.): python tools/ncu_list.py file.csv [first_id]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
d = collections.OrderedDict()
for r in rows[hdr + 1:]:
    x = dict(zip(h, r))
    d.setdefault((x['ID'], x['Kernel Name'][:48]), {})[x['Metric Name']] = float(x['Metric Value'].replace(',', ''))
for (i, n), m in d.items():
    if int(i) < first:
        continue
    g = lambda k: m.get(k, 0.0)
    print(f"{i:>3} {n:48s} {g('gpu__time_duration.sum') / 1e3:9.1f}us inst={g('smsp__inst_executed.sum') / 1e6:8.1f}M "
          f"rd={g('dram__bytes_read.sum') / 1e6:8.1f}MB wr={g('dram__bytes_write.sum') / 1e6:8.1f}MB "
          f"warps={g('sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f}% "
          f"issue={g('smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f}%")
