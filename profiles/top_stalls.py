#!/usr/bin/env python
"""Top warp-stall locations of one kernel from an ncu report (run on the GPU box; only this text travels back).
  python profiles/top_stalls.py report.ncu-rep [n]"""
import csv
import subprocess
import sys

path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r]
h = rows[hi[0]]
ci = {c: i for i, c in enumerate(h)}
body = rows[hi[0] + 1:(hi[1] if len(hi) > 1 else len(rows))]


def f(r, c):
    try:
        return float(r[ci[c]])
    except (ValueError, IndexError):
        return 0.0


print('total samples', sum(f(r, '# Samples') for r in body))
for r in sorted(body, key=lambda r: -f(r, '# Samples'))[:n]:
    stalls = {c: f(r, c) for c in h if c.startswith('stall_') and 'Not' not in c}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    print(r[ci['Address']][-5:], r[ci['Source']][:84].ljust(84), int(f(r, '# Samples')), top)
