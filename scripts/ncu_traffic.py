"""profiles/conv_traffic.json from an `ncu --set full` capture of all conv-family launches of one eager step:
mean (dram__bytes_read.sum + dram__bytes_write.sum) per launch, plus the per-launch table.
    python scripts/ncu_traffic.py gpurun_out/<capture>.ncu-rep profiles/conv_traffic.json"""
import csv, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
tot, n, table = 0.0, 0, []
for r in rows[2:]:
    name = r[ix['Kernel Name']]
    if 'conv_' not in name or 'simt' in name or 'conv_stem_tc' in name:
        continue
    rd = float(r[ix['dram__bytes_read.sum']].replace(',', '')) * scale[units[ix['dram__bytes_read.sum']]]
    wr = float(r[ix['dram__bytes_write.sum']].replace(',', '')) * scale[units[ix['dram__bytes_write.sum']]]
    tot += rd + wr
    n += 1
    table.append({'kernel': name.split('(')[0].replace('void ', ''), 'grid': r[ix['Grid Size']], 'dram_read': rd, 'dram_write': wr,
                  'time_us': float(r[ix['gpu__time_duration.sum']].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[units[ix['gpu__time_duration.sum']]],
                  'tensor_pct': float(r[ix['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']])})
json.dump({'source': rep, 'csrc_sha': bench.csrc_sha(), 'launches': n, 'dram_bytes_per_launch': tot / max(n, 1), 'dram_bytes_per_step': tot, 'per_launch': table},
          open(out, 'w'), indent=1)
print('launches', n, 'mean DRAM bytes/launch %.1f MB' % (tot / max(n, 1) / 1e6))
