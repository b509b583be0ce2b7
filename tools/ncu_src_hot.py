"""Hot spots of an `ncu --page source --csv` export: opcode histogram and the top-sampled SASS lines."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
print(rows[0][1][:80])
hdr = rows[1]
iS = hdr.index('Source'); iN = hdr.index('# Samples'); iE = hdr.index('Instructions Executed')
data = [r for r in rows[2:] if len(r) > iE and r[iN].isdigit()]
tot = sum(int(r[iN]) for r in data); totE = sum(int(r[iE]) for r in data)
print('samples', tot, 'inst', totE)
ce = collections.Counter(); cs = collections.Counter()
for r in data:
    op = [x for x in r[iS].split() if not x.startswith('@')]
    op = op[0].split('.')[0] if op else '?'
    ce[op] += int(r[iE]); cs[op] += int(r[iN])
for op, v in cs.most_common(12):
    print(f"{op:10s} samples {v:6d} {100*v/tot:5.1f}%   exec {ce[op]:9d} {100*ce[op]/totE:5.1f}%")
print('--- top lines')
top = sorted(range(len(data)), key=lambda i: -int(data[i][iN]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for i in sorted(top):
    print(i, data[i][iN], data[i][iE], data[i][iS].strip()[:90])
