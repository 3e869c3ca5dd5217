"""Summarise an `ncu --page source --csv` dump: instruction mix by opcode class and stall-sample totals.

    ncu -i X.ncu-rep --page source --csv > x_sass.csv ; python profiles/sass_summary.py x_sass.csv
"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
ops, stalls, samples_by_op = Counter(), Counter(), Counter()
total = 0
for r in rows[2:]:
    if len(r) <= iex or not r[iex]:
        continue
    n = int(float(r[iex]))
    op = r[isrc].split()[0] if not r[isrc].startswith('@') else r[isrc].split()[1]
    op = op.split('.')[0]
    ops[op] += n
    total += n
    samples_by_op[op] += int(float(r[ismp] or 0))
    for i in stall_cols:
        if r[i]:
            stalls[hdr[i]] += int(float(r[i]))
print(f'total warp instructions executed: {total:.4g}   static SASS instructions: {len(rows) - 2}')
print('--- opcode mix (warp instr, % of total, stall samples)')
for op, n in ops.most_common(28):
    print(f'{op:12s} {n:14.4g} {100 * n / total:6.2f}%  samples {samples_by_op[op]}')
print('--- stall samples')
ts = sum(stalls.values())
for k, v in stalls.most_common(10):
    print(f'{k:28s} {v:10d} {100 * v / ts:6.2f}%')
