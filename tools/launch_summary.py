"""Summarise an ncu gpu__time_duration launch list (csv): the LAST training step (the kernels between the last two Adam
launches) by kernel; `python tools/launch_summary.py FILE [n]` with n = number of trailing launches instead."""
import csv, re, sys, collections
path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
if len(sys.argv) > 2:
    rows = rows[-int(sys.argv[2]):]
else:
    ad = [i for i, r in enumerate(rows) if "adam" in r[4]]
    rows = rows[ad[-2] + 1:ad[-1] + 1] if len(ad) >= 2 else rows
tot = 0; c = collections.Counter(); k = collections.Counter()
for r in rows:
    name = re.sub(r'[\(<].*', '', r[4]).replace('void ', '').replace('mmdfn::', '')
    if 'umma_gemm' in r[4]:
        name += re.search(r'<(\d)', r[4]).group(0)
    t = int(r[-1]) / 1000; tot += t; c[name] += t; k[name] += 1
print("launches %d  total %.1f us (serialised, cold-cache ncu times)" % (len(rows), tot))
for key, v in c.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 16):
    print(f"{key:40s} {k[key]:4d} {v:9.1f} us {100*v/tot:5.1f}%")
