"""Turn the outputs of scripts/profile_round.sh (gpurun_out/*_TAG.*) into the tracked summaries under profiles/:
   python scripts/make_profiles.py r01f 'build description'"""
import csv, collections, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
desc = sys.argv[2] if len(sys.argv) > 2 else tag
G = os.path.join(ROOT, 'gpurun_out'); P = os.path.join(ROOT, 'profiles')

def last_json(path):
    for line in reversed(open(path).read().strip().splitlines()):
        if line.startswith('{'):
            return json.loads(line)

json.dump(last_json(os.path.join(G, 'bench_%s.log' % tag)), open(os.path.join(P, 'r01_bench_1gpu.json'), 'w'), indent=1)
json.dump(last_json(os.path.join(G, 'bench_ref_%s.log' % tag)), open(os.path.join(P, 'r01_bench_reference_arm.json'), 'w'), indent=1)

# launch list
src = os.path.join(G, 'launches_%s.csv' % tag)
lines = [l for l in open(src) if l.startswith('"')]
open(os.path.join(P, 'r01_bench_launches.csv'), 'w').writelines(lines)
agg = collections.OrderedDict()
for r in csv.DictReader(lines):
    a = agg.setdefault(r['Kernel Name'], [0, 0.0])
    a[0] += 1; a[1] += float(r['Metric Value']) / 1e3
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, 'r01_bench_launch_list_summary.txt'), 'w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum --clock-control none -c 400: python bench.py --steps 2 --warmup 1 --no-cpu (round 1, %s)\n' % desc)
    f.write('# kernel, launches, total us, share   (cold-cache, serialised launches: the SHARES are what compare with the CUDA-event shares of bench.py)\n')
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('%-75s %5d %12.1f us %6.2f%%\n' % (k[:75], n, us, 100 * us / tot))

for rep, out, what in [('fused_%s' % tag, 'r01_bench_fused_full.txt', 'k_fused'), ('dc_%s' % tag, 'r01_bench_dc_warp_full.txt', 'k_dc_warp')]:
    txt = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_summary.py'), os.path.join(G, rep + '.ncu-rep'),
                          os.path.join(ROOT, 'scripts', 'ncu_metric_names.txt')], capture_output=True, text=True).stdout
    open(os.path.join(P, out), 'w').write('# ncu --set full --clock-control none --import-source on, %s launch of: python bench.py --steps 1 --warmup 1 --no-cpu (4096 circuits, %s)\n' % (what, desc) + txt)

csvp = '/tmp/%s_src.csv' % tag
open(csvp, 'w').write(subprocess.run(['ncu', '-i', os.path.join(G, 'fused_%s.ncu-rep' % tag), '--page', 'source', '--print-source', 'cuda,sass', '--csv'],
                                     capture_output=True, text=True).stdout)
hot = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_lines.py'), csvp, '40'], capture_output=True, text=True).stdout
open(os.path.join(P, 'r01_fused_final_source_hotspots.txt'), 'w').write(
    '# ncu --set full --import-source on, k_fused of bench.py (4096 circuits), %s: source lines by stall samples\n' % desc + hot)
shutil.copy(os.path.join(G, 'configs_%s.json' % tag), os.path.join(P, 'r01_other_configs.json'))
shutil.copy(os.path.join(G, 'deveval_%s.json' % tag), os.path.join(P, 'r01_device_eval_stage.json'))
print(open(os.path.join(P, 'r01_bench_launch_list_summary.txt')).read())
print(open(os.path.join(P, 'r01_bench_fused_full.txt')).read())
