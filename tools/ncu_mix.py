"""Instruction mix of one kernel from `ncu -i X.ncu-rep --page source --csv` output."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia = hdr.index('Source'); ie = hdr.index('Instructions Executed'); it = hdr.index('Thread Instructions Executed'); isamp = hdr.index('# Samples')
ops = collections.Counter(); thr = collections.Counter(); samp = collections.Counter()
n = 0; tot = 0
for r in rows[2:]:
    if len(r) < 10 or r[0] in ('Kernel Name', 'Address'):
        if r and r[0] == 'Kernel Name' and n > 0:
            break
        continue
    n += 1
    src = r[ia].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2).split('.')[0] if m else src
    e = int(r[ie] or 0); t = int(r[it] or 0)
    ops[op] += e; thr[op] += t; samp[op] += int(r[isamp] or 0); tot += e
print('static instrs', n, 'executed warp-instrs', tot, 'samples', sum(samp.values()))
for op, c in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f'{op:12s} {c:12d} {100*c/tot:5.1f}%  avgthr {thr[op]/max(c,1):5.1f}  samples {samp[op]}')
