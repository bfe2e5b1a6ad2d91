"""One line per profiled launch from an .ncu-rep: the counters the roofline / DESIGN tables quote.
    python scripts/ncu_table.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [('Kernel Name', 'kernel', 34), ('Grid Size', 'grid', 12), ('gpu__time_duration.sum', 'time', 9),
        ('dram__bytes_read.sum', 'dram_rd', 10), ('dram__bytes_write.sum', 'dram_wr', 10),
        ('lts__t_bytes.sum', 'l2_bytes', 10),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%', 8),
        ('sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active', 'hmma%', 6),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%', 6),
        ('lts__t_sector_hit_rate.pct', 'l2hit%', 7),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%', 7),
        ('launch__registers_per_thread', 'regs', 5)]
print(' '.join('%-*s' % (w, n) for _, n, w in cols))
for r in rows[2:]:
    out = []
    for key, n, w in cols:
        if key not in ix:
            out.append('%-*s' % (w, '-')); continue
        v = r[ix[key]]
        if key == 'Kernel Name':
            v = v.replace('void ', '').split('(')[0][-w:]
        else:
            try:
                f = float(v.replace(',', ''))
                u = units[ix[key]]
                v = ('%.4g' % f) + ({'Mbyte': 'M', 'Gbyte': 'G', 'Kbyte': 'K', 'byte': 'B', 'us': 'us', 'ms': 'ms', 'ns': 'ns'}.get(u, ''))
            except ValueError:
                pass
        out.append('%-*s' % (w, v[:w]))
    print(' '.join(out))
