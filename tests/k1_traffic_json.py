"""profiles/r2_k1_traffic.json from an `ncu --set full` capture of ONE k1_fused launch (bench.py reads `dram_bytes_per_launch`
for the `roofline.traffic` key).
usage: ncu -i gpurun_out/r2_k1_fused.ncu-rep --page raw --csv > raw.csv ; python tests/k1_traffic_json.py raw.csv out.json "<provenance>" """
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
body = [r for r in rows[2:] if 'k1_fused' in r[h.index('Kernel Name')]]
r = body[-1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum']
m = {k: {'value': r[h.index(k)], 'unit': units[h.index(k)]} for k in want if k in h}
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
tot = sum(float(m[k]['value'].replace(',', '')) * scale.get(m[k]['unit'], 1) for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
out = {'kernel': 'k1_fused', 'shape': {'N': 22680, 'HW': 1620, 'n_obj': 1}, 'dram_bytes_per_launch': int(tot),
       'captured_at': sys.argv[3] if len(sys.argv) > 3 else 'ncu --set full --clock-control none', 'metrics': m}
json.dump(out, open(sys.argv[2], 'w'), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != 'metrics'}))
