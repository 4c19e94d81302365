#!/usr/bin/env python
"""Opcode evidence from the built library: `cuobjdump -sass` of resdepth_b200/_lib/libresdepth_b200.so, per kernel the
instruction count and the counts of the Blackwell-specific mnemonics (tcgen05 MMA = UTCHMMA / UTCQMMA / UTCIMMA..., TMEM
loads = LDTM, TMA = UTMALDG / UTMASTG / UTMAPF, tcgen05.commit = UTCBAR, mbarrier = SYNCS, cp.async = LDGSTS, packed fp32
FMA = FFMA2) plus registers per thread from `cuobjdump -res-usage`.

    python profiles/sass_summary.py > profiles/r2_sass_summary.md      (build container; no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'resdepth_b200', '_lib', 'libresdepth_b200.so')
KEYS = ['UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'UTCOMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'UTCBAR', 'UTCATOMSWS', 'USETMAXREG',
        'SYNCS', 'LDGSTS', 'FFMA2', 'HMMA', 'STG', 'LDG', 'STS', 'LDS', 'SHFL']


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    res = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r'Function (\S+):', line)
        if m:
            cur = m.group(1)
        m = re.search(r'REG:(\d+).*?SHARED:(\d+)', line)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)))
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m and cur:
            op = m.group(1)
            counts[cur]['__total__'] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[cur][k] += 1
                    break
    names = demangle(list(counts))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print('# SASS opcode summary of `libresdepth_b200.so` (sm_100a)\n')
    print('`python profiles/sass_summary.py` (= `cuobjdump -sass` + `-res-usage`, counted per kernel). Library totals: '
          + ', '.join(f'{tot[k]} `{k}`' for k in KEYS if tot[k]) + f'; {len(counts)} kernels, {tot["__total__"]} instructions.\n')
    show = [k for k in KEYS if tot[k]]
    print('| kernel | regs | static smem B | instr | ' + ' | '.join(show) + ' |')
    print('|---|---:|---:|---:|' + '---:|' * len(show))
    for fn, c in sorted(counts.items(), key=lambda kv: -(kv[1]['UTCHMMA'] + kv[1]['UTCQMMA']) * 100000 - kv[1]['__total__']):
        name = re.sub(r'^void ', '', names.get(fn, fn))
        name = re.sub(r'\(.*', '', name).replace('rd::', '')
        r = regs.get(fn, ('', ''))
        print(f'| `{name[:90]}` | {r[0]} | {r[1]} | {c["__total__"]} | ' + ' | '.join(str(c[k]) if c[k] else '' for k in show) + ' |')


if __name__ == '__main__':
    sys.exit(main())
