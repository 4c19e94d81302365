#!/usr/bin/env python
"""Top warp-stall source locations of the kernels in an ncu report (run on the GPU box; only this text travels back).
  python profiles/top_stalls.py report.ncu-rep [n_lines] [name_filter]
For every kernel whose name contains `name_filter`: duration, DRAM %, SM %, then the n hottest source/SASS lines with
their two dominant stall reasons."""
import csv
import io
import subprocess
import sys

path = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
flt = sys.argv[3] if len(sys.argv) > 3 else ''

raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
kernels = rows[2:]
want = ['gpu__time_duration.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active']
for k, r in enumerate(kernels):
    name = r[ci['Kernel Name']]
    if flt and flt not in name:
        continue
    print(f'\n=== [{k}] {name[:110]} grid {r[ci["Grid Size"]]}')
    print('   ' + ', '.join(f'{w.split(".")[0]}={r[ci[w]]}' for w in want if w in ci))
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--launch-skip', str(k), '--launch-count', '1'],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, x in enumerate(srows) if 'Source' in x and '# Samples' in x]
    if not hi:
        print('   (no source page)')
        continue
    h = srows[hi[0]]
    si = {c: i for i, c in enumerate(h)}
    body = srows[hi[0] + 1:(hi[1] if len(hi) > 1 else len(srows))]

    def f(x, c):
        try:
            return float(x[si[c]])
        except (ValueError, IndexError, KeyError):
            return 0.0
    total = sum(f(x, '# Samples') for x in body)
    print('   total samples', total)
    agg = {}
    for x in body:
        for c in h:
            if c.startswith('stall_') and 'Not' not in c:
                agg[c] = agg.get(c, 0.0) + f(x, c)
    print('   stall totals:', sorted(((round(v), c) for c, v in agg.items() if v > 0), reverse=True)[:8])
    for x in sorted(body, key=lambda x: -f(x, '# Samples'))[:n]:
        stalls = {c: f(x, c) for c in h if c.startswith('stall_') and 'Not' not in c}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
        print('   ', x[si['Address']][-5:], x[si['Source']][:90].ljust(90), int(f(x, '# Samples')), top)
    # samples per code region (64 instructions), in address order, with the marker opcodes the region contains: tells the
    # warp roles apart (UTMALDG = TMA producer, UTCHMMA = MMA issuer, LDTM = epilogue) when the kernel is warp-specialised
    if len(sys.argv) > 4 and sys.argv[4] == 'regions':
        marks = ('UTMALDG', 'UTMASTG', 'UTCHMMA', 'UTCBAR', 'LDTM', 'SYNCS', 'SHFL', 'STS', 'LDS', 'STG', 'LDG', 'BAR')
        ordered = sorted(body, key=lambda x: x[si['Address']])
        print('   regions (64 instructions each): start address, samples, share, two dominant stalls, marker opcodes')
        for b0 in range(0, len(ordered), 64):
            blk = ordered[b0:b0 + 64]
            smp = sum(f(x, '# Samples') for x in blk)
            if smp < 0.004 * total:
                continue
            st = {}
            for x in blk:
                for c in h:
                    if c.startswith('stall_') and 'Not' not in c:
                        st[c] = st.get(c, 0.0) + f(x, c)
            top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
            present = [m for m in marks if any(m in x[si['Source']] for x in blk)]
            print(f'    {blk[0][si["Address"]][-5:]}  {int(smp):6d}  {100 * smp / total:5.1f}%  '
                  f'{[(c[6:], int(v)) for c, v in top]}  {" ".join(present)}')
