#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep) into the few numbers DESIGN.md / bench.py quote.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_xxx.md "title" [alg_bytes_per_launch]

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU)."""
import csv
import io
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of ncu peak'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU (MUFU) pipe %'),
    ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'FMA pipe %'),
    ('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'ALU pipe %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
]


def main():
    rep, out, title = sys.argv[1:4]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ['# %s' % title, '', 'Source report: `%s` (`ncu --set full --clock-control none --import-source on`, '
             '1x B200, one capture per kernel; times under ncu are cold-cache).' % rep, '']
    for r in data:
        name = r[idx['Kernel Name']].split('(')[0].replace('void ', '')
        lines.append('## `%s`' % name)
        lines.append('')
        lines.append('| metric | value |')
        lines.append('|---|---|')
        vals = {}
        for k, label in KEYS:
            if k in idx:
                vals[k] = r[idx[k]]
                lines.append('| %s | %s %s |' % (label, r[idx[k]], units[idx[k]]))
        try:
            rd = float(vals['dram__bytes_read.sum'].replace(',', ''))
            wr = float(vals['dram__bytes_write.sum'].replace(',', ''))
            ur, uw = units[idx['dram__bytes_read.sum']], units[idx['dram__bytes_write.sum']]
            scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
            tot = rd * scale[ur] + wr * scale[uw]
            dur = float(vals['gpu__time_duration.sum'].replace(',', ''))
            du = {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1}[units[idx['gpu__time_duration.sum']]]
            lines.append('| DRAM traffic (read+write) | %.1f MB -> %.0f GB/s |' % (tot / 1e6, tot / (dur * du) / 1e9))
        except Exception:
            pass
        lines.append('')
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines[:60]))


if __name__ == '__main__':
    main()
