"""Summarise an `ncu --page source --print-source cuda,sass --csv` export per CUDA source line:
instructions executed, stall samples.  usage: ncu_lines.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
cur_file = None
out = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
    if len(r) > 8 and r[0] == 'Line No':
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0] not in ('', 'Line No'):
        d = dict(zip(hdr, r))
        try:
            out.append((cur_file, int(r[0]), r[1].strip()[:90], int(d['Instructions Executed']), int(d['# Samples'])))
        except ValueError:
            pass
tot_i = sum(o[3] for o in out); tot_s = sum(o[4] for o in out)
print('total instr %d samples %d' % (tot_i, tot_s))
for o in sorted(out, key=lambda o: -o[4])[:top]:
    print('%5.1f%% samp %5.1f%% inst  %s:%d  %s' % (100.0 * o[4] / max(tot_s, 1), 100.0 * o[3] / max(tot_i, 1), o[0], o[1], o[2]))

# aggregate by file / line ranges given as extra args  name:file:lo-hi
import re
if len(sys.argv) > 3:
    print('--- regions')
    for spec in sys.argv[3:]:
        name, f, rng = spec.split(':')
        lo, hi = map(int, rng.split('-'))
        si = sum(o[3] for o in out if o[0] == f and lo <= o[1] <= hi)
        ss = sum(o[4] for o in out if o[0] == f and lo <= o[1] <= hi)
        print('%-14s %5.1f%% samp %5.1f%% inst' % (name, 100.0 * ss / max(tot_s, 1), 100.0 * si / max(tot_i, 1)))
    files = {}
    for o in out:
        files.setdefault(o[0], [0, 0])
        files[o[0]][0] += o[3]; files[o[0]][1] += o[4]
    for f, (i, s_) in sorted(files.items(), key=lambda kv: -kv[1][1]):
        print('file %-28s %5.1f%% samp %5.1f%% inst' % (f, 100.0 * s_ / max(tot_s, 1), 100.0 * i / max(tot_i, 1)))
