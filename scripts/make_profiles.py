"""Turn the outputs of scripts/profile_round.sh (gpurun_out/*_TAG.*) into the tracked summaries under profiles/:
   python scripts/make_profiles.py TAG 'build description' [round-prefix, default r02]
Writes profiles/<rNN>_bench_1gpu.json, _bench_reference_arm.json, _bench_launches.csv, _bench_launch_list_summary.txt,
_bench_fused_full.txt, _bench_dc_warp_full.txt, _fused_source_hotspots.txt, _other_configs.json, _device_eval_stage.json and
<rNN>_ncu_summary.json (the per-kernel ncu figures bench.py attaches to its roofline objects)."""
import csv, collections, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
desc = sys.argv[2] if len(sys.argv) > 2 else tag
rnd = sys.argv[3] if len(sys.argv) > 3 else 'r02'
G = os.path.join(ROOT, 'gpurun_out'); P = os.path.join(ROOT, 'profiles')


def out(name):
    return os.path.join(P, '%s_%s' % (rnd, name))


def last_json(path):
    if not os.path.exists(path):
        return None
    for line in reversed(open(path).read().strip().splitlines()):
        if line.startswith('{'):
            return json.loads(line)


for src, dst in (('bench_%s.log', 'bench_1gpu.json'), ('bench_ref_%s.log', 'bench_reference_arm.json')):
    j = last_json(os.path.join(G, src % tag))
    if j:
        json.dump(j, open(out(dst), 'w'), indent=1)

# launch list
src = os.path.join(G, 'launches_%s.csv' % tag)
if os.path.exists(src):
    lines = [l for l in open(src) if l.startswith('"')]
    open(out('bench_launches.csv'), 'w').writelines(lines)
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        a = agg.setdefault(r['Kernel Name'], [0, 0.0])
        a[0] += 1; a[1] += float(r['Metric Value']) / 1e3
    tot = sum(v[1] for v in agg.values())
    with open(out('bench_launch_list_summary.txt'), 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none -c 400: python bench.py --steps 2 --warmup 1 --no-cpu (%s)\n' % desc)
        f.write('# kernel, launches, total us, share   (cold-cache, serialised launches: the SHARES are what compare with the CUDA-event shares of bench.py;\n')
        f.write('# serialised = the DC kernel cannot run beside the HB kernel here, so its share is its stand-alone duration)\n')
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('%-75s %5d %12.1f us %6.2f%%\n' % (k[:75], n, us, 100 * us / tot))
        # the step's own kernels only (the list also holds the FP64-peak microbenchmark, the device-evaluation stage
        # kernel of roofline_device_eval and torch's fills / copies, which run outside the timed steps)
        own = ('k_fused', 'k_dc_', 'k_prepare', 'k_order', 'k_init', 'k_finish')
        step = {k: v for k, v in agg.items() if any(o in k for o in own)}
        stot = sum(v[1] for v in step.values())
        f.write('# shares among the kernels of a step (compare roofline.share_serialised of bench.py):\n')
        for k, (n, us) in sorted(step.items(), key=lambda kv: -kv[1][1]):
            f.write('#   %-71s %5d %12.1f us %6.2f%%\n' % (k[:71], n, us, 100 * us / stot))
    print(open(out('bench_launch_list_summary.txt')).read())


def num(x):
    try:
        return float(str(x).replace(',', ''))
    except Exception:
        return None


def raw_metrics(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return {}
    hdr, units, vals = rows[0], rows[1], rows[2]
    scale = {'byte': 1., 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1., 'ms': 1e3, 's': 1e6}   # -> bytes, us
    out = {}
    for h, u, v in zip(hdr, units, vals):
        x = num(v)
        out[h] = x * scale[u] if (x is not None and u in scale) else v
    return out


summary = {}
for rep, name, what, key in [('fused_%s' % tag, 'bench_fused_full.txt', 'k_fused', 'k_fused'),
                             ('dc_%s' % tag, 'bench_dc_warp_full.txt', 'k_dc_warp', 'k_dc_warp'),
                             ('deveval_%s' % tag, 'device_eval_full.txt', 'k_stage_device_eval', 'k_stage_device_eval')]:
    path = os.path.join(G, rep + '.ncu-rep')
    if not os.path.exists(path):
        continue
    txt = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_summary.py'), path,
                          os.path.join(ROOT, 'scripts', 'ncu_metric_names.txt')], capture_output=True, text=True).stdout
    open(out(name), 'w').write('# ncu --set full --clock-control none --import-source on, one %s launch (%s)\n' % (what, desc) + txt)
    m = raw_metrics(path)
    rd, wr = num(m.get('dram__bytes_read.sum')), num(m.get('dram__bytes_write.sum'))
    summary[key] = {
        'source': 'profiles/%s_%s' % (rnd, name),
        'dram_bytes_per_launch': (rd + wr) if rd is not None and wr is not None else None,
        'gpu_time_us': num(m.get('gpu__time_duration.sum')),
        'fp64_pipe_pct': num(m.get('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active')),
        'dmma_pipe_pct': num(m.get('sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active')),
        'issue_active_pct': num(m.get('smsp__issue_active.avg.pct_of_peak_sustained_active')),
        'warps_active_pct': num(m.get('sm__warps_active.avg.pct_of_peak_sustained_active')),
        'dram_throughput_pct': num(m.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')),
        'stall_barrier_per_issue': num(m.get('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio')),
        'note': 'dram bytes: dram__bytes_read.sum + dram__bytes_write.sum of the captured launch (units as ncu reports them are '
                'converted to bytes; gpu_time_us = gpu__time_duration.sum of the captured launch in microseconds',
    }
    print(open(out(name)).read())
if summary:
    json.dump(summary, open(out('ncu_summary.json'), 'w'), indent=1)

rep = os.path.join(G, 'fused_%s.ncu-rep' % tag)
if os.path.exists(rep):
    csvp = '/tmp/%s_src.csv' % tag
    open(csvp, 'w').write(subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'],
                                         capture_output=True, text=True).stdout)
    hot = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_lines.py'), csvp, '40'], capture_output=True, text=True).stdout
    open(out('fused_source_hotspots.txt'), 'w').write(
        '# ncu --set full --import-source on, k_fused of bench.py (4096 circuits), %s: source lines by stall samples\n' % desc + hot)
for src, dst in (('configs_%s.json', 'other_configs.json'), ('deveval_%s.json', 'device_eval_stage.json')):
    if os.path.exists(os.path.join(G, src % tag)):
        shutil.copy(os.path.join(G, src % tag), out(dst))
