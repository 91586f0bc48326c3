#!/usr/bin/env python
"""ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <cmd>`) ->
markdown table of kernels with launch counts, total / average time and the share of this library's kernels.

    python profiles/summarize_launches.py gpurun_out/launches.csv profiles/rNN_launches.md "title"
"""
import csv
import sys


def main():
    src, out, title = sys.argv[1:4]
    rows = [r for r in csv.reader(open(src, errors='replace')) if len(r) >= 15 and r[0].isdigit()]
    order, agg = [], {}
    for r in rows:
        name, ns = r[4], float(r[14])
        short = name.replace('void ', '').split('(')[0]
        if short not in agg:
            agg[short] = [0, 0.0]
            order.append(short)
        agg[short][0] += 1
        agg[short][1] += ns
    ours = sum(v[1] for k, v in agg.items() if k.startswith('rn::'))
    lines = ['# %s' % title, '',
             '`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (`%s`). Per-launch times under ncu are '
             'cold-cache and serialised: the kernels\' SHARE of the step is what must agree with the live event times '
             'in `bench.py`.' % src.replace('gpurun_out/', 'profiles/'), '',
             '| kernel | launches | total us | avg us | share of librndense time |', '|---|---:|---:|---:|---:|']
    for k in order:
        n, t = agg[k]
        share = '%.1f %%' % (100.0 * t / ours) if k.startswith('rn::') else '(torch)'
        lines.append('| `%s` | %d | %.1f | %.1f | %s |' % (k[:70], n, t / 1e3, t / n / 1e3, share))
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines[-12:]))


if __name__ == '__main__':
    main()
