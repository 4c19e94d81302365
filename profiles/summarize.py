#!/usr/bin/env python
"""Turns the raw profiler outputs brought back in gpurun_out/ into the committed summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_r1.csv > profiles/r1_launches.md
  python profiles/summarize.py ncu gpurun_out/prof_rows.ncu-rep gpurun_out/prof_reduce.ncu-rep > profiles/r1_ncu_rows_reduce.md
"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict


def launches(path):
    rows = list(csv.DictReader(l for l in open(path) if not l.startswith('==')))
    idx = [i for i, r in enumerate(rows) if 'adam_kernel' in r['Kernel Name']]
    start, end = (idx[-2] + 1, idx[-1] + 1) if len(idx) >= 2 else (0, len(rows))
    step = rows[start:end]
    agg = OrderedDict()
    total = 0.0
    for r in step:
        name = re.sub(r'\(.*', '', r['Kernel Name']).replace('rd::', '').replace('void ', '')
        us = float(r['Metric Value']) / 1e3
        total += us
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    print('# ncu launch list, one train step (batch 64, 3-ch 256x256, depth 5) -- `ncu --metrics gpu__time_duration.sum '
          '--clock-control none`')
    print(f'\nLaunches in the step: {len(step)}; serialised cold-cache total {total / 1e3:.2f} ms '
          '(compare SHARES with bench.py, not absolutes).\n')
    print('| kernel | launches | total us | share |')
    print('|---|---:|---:|---:|')
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| `{name}` | {n} | {us:.1f} | {100 * us / total:.1f} % |')
    print('\n## Every launch of the step, in order\n')
    print('| # | kernel | grid | us |')
    print('|---:|---|---|---:|')
    for i, r in enumerate(step):
        name = re.sub(r'\(.*', '', r['Kernel Name']).replace('rd::', '').replace('void ', '')
        print(f"| {i} | `{name}` | {r['Grid Size']} | {float(r['Metric Value']) / 1e3:.1f} |")


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum.per_second', 'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread',
        'sm__cycles_elapsed.max', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__shared_mem_per_block_dynamic']


def ncu(paths):
    print('# ncu --set full captures (clock-control none), raw-page extract\n')
    for path in paths:
        out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        ci = {h: i for i, h in enumerate(hdr)}
        print(f'## {path.split("/")[-1]}\n')
        cols = [w for w in WANT if w in ci]
        print('| kernel | grid | ' + ' | '.join(f'{c} [{units[ci[c]]}]' for c in cols) + ' |')
        print('|---|---|' + '---:|' * len(cols))
        for r in rows[2:]:
            name = re.sub(r'\(.*', '', r[ci['Kernel Name']]).replace('void ', '')
            print(f"| `{name}` | {r[ci['Grid Size']]} | " + ' | '.join(r[ci[c]] for c in cols) + ' |')
        print()


def step(md_path):
    """Re-format the table written by `summarize.py ncu` for a --metrics capture of one whole step (run on the GPU box,
    only the small .md travels back) into the committed per-launch + per-kernel summary."""
    rows = [l for l in open(md_path).read().splitlines() if l.startswith('| `')]
    hdr = ['kernel', 'grid', 'time us', 'DRAM rd MB', 'DRAM wr MB', 'DRAM %', 'tensor pipe %', 'SM %', 'L2 hit %', 'regs',
           'warps active %', 'dyn smem KB']
    print('# ncu metrics of every kernel launch of one train step (batch 64, 3-ch 256x256, depth 5)\n')
    print('Command (on the B200 box): `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,'
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,'
          'sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,launch__registers_per_thread,'
          f'sm__warps_active.avg.pct_of_peak_sustained_active,launch__shared_mem_per_block_dynamic --clock-control none -s {len(rows)} '
          f'-c {len(rows)} -o /tmp/step python profiles/one_step.py 2` (profiles/capture_step.sh), then `summarize.py ncu /tmp/step.ncu-rep` there and '
          f'`summarize.py step` here.  {len(rows)} consecutive launches = one full step (the window starts at the second step). '
          'ncu times are cold-cache and serialised: compare SHARES with `bench.py`, not absolutes.\n')
    print('| # | ' + ' | '.join(hdr) + ' |')
    print('|---:|---|---|' + '---:|' * 10)
    agg, tot = OrderedDict(), 0.0
    for i, l in enumerate(rows):
        c = [x.strip() for x in l.strip('|').split('|')]
        name = c[0].replace('rd::', '')
        t, rd, wr = float(c[2]), float(c[3]) * 1e3, float(c[4]) * 1e3
        vals = [name, c[1], f'{t:.1f}', f'{rd:.1f}', f'{wr:.1f}', f'{float(c[5]):.1f}', f'{float(c[6]):.1f}',
                f'{float(c[7]):.1f}', f'{float(c[8]):.1f}', c[9], f'{float(c[10]):.1f}', f'{float(c[11]):.1f}']
        print(f'| {i} | ' + ' | '.join(vals) + ' |')
        a = agg.setdefault(name.replace('`', ''), [0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += t; a[2] += rd + wr; a[3] += float(c[6]) * t; a[4] += float(c[5]) * t
        tot += t
    print('\n## Per kernel\n')
    print('| kernel | launches | total us | share | DRAM traffic MB / launch | time-weighted tensor pipe % | time-weighted DRAM % |')
    print('|---|---:|---:|---:|---:|---:|---:|')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| `{k}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f} % | {a[2] / a[0]:.1f} | {a[3] / a[1]:.1f} | {a[4] / a[1]:.1f} |')
    print(f'\nTotal {tot / 1e3:.2f} ms over {len(rows)} launches.')


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2])
    elif sys.argv[1] == 'step':
        step(sys.argv[2])
    else:
        ncu(sys.argv[2:])
