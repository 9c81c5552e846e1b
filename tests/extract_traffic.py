"""Summarise an `ncu --set full` capture of the K1 kernels: per-kernel duration, DRAM bytes, tensor-pipe activity.
usage: python tests/extract_traffic.py <raw.csv from `ncu -i rep --page raw --csv`> <out.json>"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
h, body = rows[0], rows[2:]
ix = {n: i for i, n in enumerate(h)}
def col(r, name):
    for k, i in ix.items():
        if k == name:
            return float(r[i].replace(',', '')) if r[i] not in ('', 'n/a') else 0.0
    return 0.0
units = dict(zip(h, rows[1]))
def to_bytes(v, u):
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
def to_us(v, u):
    return v * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1, 'msecond': 1e3}.get(u, 1)
out = {'kernels': []}
tot = 0.0
for r in body:
    name = r[ix['Kernel Name']].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
    rd = to_bytes(col(r, 'dram__bytes_read.sum'), units['dram__bytes_read.sum'])
    wr = to_bytes(col(r, 'dram__bytes_write.sum'), units['dram__bytes_write.sum'])
    k = {'kernel': name, 'duration_us': round(to_us(col(r, 'gpu__time_duration.sum'), units['gpu__time_duration.sum']), 2),
         'dram_read_bytes': int(rd), 'dram_write_bytes': int(wr),
         'tensor_pipe_active_pct': col(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'),
         'issue_active_pct': col(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
         'registers': int(col(r, 'launch__registers_per_thread'))}
    out['kernels'].append(k)
    tot += rd + wr
out['dram_bytes_per_call'] = int(tot)
json.dump(out, open(sys.argv[2], 'w'), indent=1)
print(json.dumps(out, indent=1))
