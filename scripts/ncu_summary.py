"""Pick the metrics the round summaries quote out of an ncu report:
   python scripts/ncu_summary.py gpurun_out/x.ncu-rep [metric-list-file] > profiles/rNN_x_full.txt
(metric names default to the ones profiles/r01_bench_fused_full.txt lists)"""
import csv, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
names_file = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, 'profiles', 'r01_bench_fused_full.txt')
names = []
for line in open(names_file):
    t = line.split()
    if t and not line.startswith('#') and ('__' in t[0] or t[0].startswith('launch')) and t[0] not in names:
        names.append(t[0])
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}
print('# kernel: %s' % vals[col['Kernel Name']])
for n in names:
    if n in col:
        print('%s [%s] = %s' % (n, units[col[n]], vals[col[n]]))
