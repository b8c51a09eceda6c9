#!/usr/bin/env python
"""bench.py -- HB solves/sec on BASELINE.json's configuration 4.

Workload (config.workload): the reference-pinned DiffAmp netlist (tests/HB_DiffAmp.py:13-69 of the reference,
resistor loads, 4 Gummel-Poon BJTs with junction capacitances, N=10, K=10, W=210), a (vbias, vin) grid, every point
an independent COLD start (hb.run(y), V0=None): DC operating point, 8+ source-stepping levels, Newton to the
reference's default tolerances.  One "step" = one pass over the whole grid.  The ActiveLoad netlist BASELINE.json
names needs PNP devices, which the reference's HB does not implement; `--workload activeload` runs it
with the PNP extension instead.

    python bench.py [--gpus N] [--steps K] [--warmup W]            (torchrun for N > 1, one rank per GPU)
    python bench.py --scaling strong ...                           (the ONE 4096-point sweep split over the ranks)
    python bench.py --impl reference ...                           (CPU arm: the oracle port on all host cores)

--scaling weak (default): 64 x 64 points PER GPU -- a 64 x (64 N) (vbias x vin) grid whose points are dealt round-robin to the ranks.
value  : solves/s with the parameter arrays resident in HBM (DeviceBatch.solve: DC + ladder + Newton on device,
         results packed and gathered with one NCCL all_gather_into_tensor for N > 1)
e2e    : solves/s through hb.run_sweep() (hb.distributed = True for N > 1): host arrays in, host arrays out
roofline: the kernel with the largest device time share, from CUDA events recorded around every launch; `frac`
         counts the floating-point work the kernel EXECUTES in its linear solves against the live FP64 peak
roofline_device_eval: the stand-alone device-evaluation stage kernel against the measured HBM bandwidth
parity : GPU results against the oracle on a seeded random sample of the grid, outside the timed region
cpu_baseline: the numpy oracle (oracle/hb_oracle.py, kind "port": the reference is Python and cannot travel)
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

GRID = 64


def grid_points(ncols, workload):
    """(vbias, vin) of a 64 x ncols grid (64 bias points, ncols amplitudes over the same range), flattened vbias-major.
    Weak scaling refines the AMPLITUDE axis (ncols = 64 N): dealt round-robin, every rank then holds what one GPU holds at
    N = 1 -- all 64 bias points with 64 amplitudes each -- instead of 64 N bias points with 8 amplitudes each."""
    if workload == 'activeload':
        vb = np.linspace(1.3, 1.7, GRID)
        vin = np.linspace(20e-6, 1e-3, ncols)
    else:
        vb = np.linspace(1.0, 1.4, GRID)
        vin = np.linspace(100e-6, 50e-3, ncols)
    VB, VIN = np.meshgrid(vb, vin, indexing='ij')
    return VB.ravel(), VIN.ravel()


def ncu_summary(name):
    """ncu figures of this round's capture of a kernel (profiles/r02_ncu_summary.json, written by
    scripts/make_profiles.py from the .ncu-rep of the same bench command), or None"""
    path = os.path.join(ROOT, 'profiles', 'r02_ncu_summary.json')
    if not os.path.exists(path):
        return None
    try:
        return json.load(open(path)).get(name)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------ CPU arm
def _oracle_point(args):
    os.environ['OMP_NUM_THREADS'] = '1'
    workload, vb, vin = args
    import circuits
    from oracle import hb_oracle as ho
    y = ho.FlatNetlist()
    if workload == 'activeload':
        h = circuits.diffamp_activeload(y, vin=vin / 2, vbias=vb)
        hb = ho.HBOracle('HB1', h['freq'], h['K'], pnp_extension=True)
    else:
        h = circuits.diffamp(y, vin=vin / 2, vbias=vb)
        hb = ho.HBOracle('HB1', h['freq'], h['K'])
    conv = hb.run(y)[0]
    V = hb.V[:, 0].copy() if conv else None
    return bool(conv), hb.total_newton, V, [l[1] for l in hb.level_log]


def cpu_sample(workload, VB, VIN, idx, cores):
    """Oracle cold solves of the grid points idx on `cores` worker processes: (solves/s, seconds, results)."""
    import multiprocessing as mp
    jobs = [(workload, float(VB[i]), float(VIN[i])) for i in idx]
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        pool.map(_oracle_point, jobs[:cores])            # warm the workers (imports)
        t = time.perf_counter()
        out = pool.map(_oracle_point, jobs, chunksize=1)
        dt = time.perf_counter() - t
    return len(jobs) / dt, dt, out


def sample_indices(total, n):
    """fixed random subset of the grid (numpy.random.default_rng(0))"""
    return np.sort(np.random.default_rng(0).choice(total, size=min(n, total), replace=False))


def parity_report(res_V, res_levels, res_conv, idx, oracle_out):
    """GPU vs oracle on the sampled points: worst relative phasor error, identical per-level iteration counts"""
    worst, iters_equal, conv_equal = 0.0, True, True
    for j, i in enumerate(idx):
        conv_o, _, Vo, lev_o = oracle_out[j]
        conv_equal = conv_equal and (bool(res_conv[i] > 0) == bool(conv_o))
        if not conv_o or res_conv[i] <= 0:
            continue
        worst = max(worst, float(np.max(np.abs(res_V[i] - Vo)) / np.max(np.abs(Vo))))
        if res_levels is not None:
            it = res_levels[i]
            iters_equal = iters_equal and (it[it > 0].tolist() == list(lev_o))
    return {'points': int(len(idx)), 'worst_rel': worst, 'iters_equal': bool(iters_equal), 'converged_equal': bool(conv_equal),
            'tolerance': 1e-9, 'ok': bool(worst < 1e-9 and iters_equal and conv_equal),
            'sample': 'numpy.random.default_rng(0) subset of the bench grid, oracle cold solves, compared outside the timed region'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    npts = min(512, max(64, 16 * cores))
    VB, VIN = grid_points(GRID, args.workload)
    idx = sample_indices(len(VB), npts)
    vals = []
    for s in range(args.warmup + args.steps):
        v, dt, _ = cpu_sample(args.workload, VB, VIN, idx, cores)
        if s >= args.warmup:
            vals.append((v, dt))
    value = sum(npts for _ in vals) / sum(dt for _, dt in vals)
    sample = '%d of the 4096 grid points (default_rng(0) subset), cold start each, %d worker processes, 1 BLAS thread each' % (
        npts, cores)
    line = {
        'impl': 'reference', 'metric': 'HB solves/sec', 'value': value, 'unit': 'solves/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * statistics.mean(dt for _, dt in vals),
        'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.workload, args.gpus, args.scaling),
        'cpu_baseline': {'value': value, 'unit': 'solves/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'solves/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(workload, world, scaling):
    name = ('HB_DiffAmp_ActiveLoad netlist + PNP extension' if workload == 'activeload' else
            'HB_DiffAmp (cfg 4, reference-pinned resistor-load variant of the DiffAmp bias/amplitude sweep)')
    total = GRID * GRID * (world if scaling == 'weak' else 1)
    return {'workload': '%s: %d-point (vbias x vin) sweep of cold-start HB solves (%s), N=10 nodes, K=10 harmonics, '
                        'W=210, reltol 1e-3 / abstol 1e-6 / maxiter 100' % (
                            name, total, '64x64 per GPU' if scaling == 'weak' else 'the one 64x64 sweep split over the GPUs'),
            'points_per_gpu': total // world, 'points_total': total, 'start': 'cold (V0=None, DC on GPU)',
            'parallelism': 'independent circuits dealt round-robin to the ranks (b % N); one NCCL all_gather_into_tensor of '
                           'the packed results; no collective inside the Newton loop',
            'l2': 'explicit flush: a 256 MiB buffer is rewritten before every timed step (untimed); per-step HBM '
                  'working set is parameters + results (~80 MB), the Newton state itself stays on chip'}


# ------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,' \
            'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
            'clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower() == 'active' for r in self.rows)]
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def device_eval_roofline(plan, topo, db, lib, L, hbm_gbs):
    """The stand-alone device-evaluation stage kernel (yhb_stage_device_eval: time samples + raw model parameters in,
    14 model outputs per BJT sample out, all through HBM) on 4x the grid's circuits at their converged waveforms,
    against the measured HBM bandwidth.  Algorithmic bytes: S (48 D_diode + 160 D_bjt) per circuit."""
    import torch
    rep = 4
    B = db.B * rep
    W, S = plan.W, plan.S
    nd, nb, nc = len(topo.diodes), len(topo.bjts), len(topo.cubics)
    t = {n: v.repeat((rep,) + (1,) * (v.dim() - 1)).contiguous() for n, v in db.t.items()}
    p = L.Params()
    p.batch = B
    for n, v in t.items():
        setattr(p, n, v.data_ptr())
    V = db.V.repeat(rep, 1).contiguous()
    vt = torch.empty_like(V)
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.yhb_stage_ifft(plan.handle, B, V.data_ptr(), vt.data_ptr(), st))
    vtprev = (vt * 0.999).contiguous()                   # limiter reference: the 'previous iterate'
    dout = torch.empty((B, max(nd, 1), 2, S), dtype=torch.float64, device=V.device)
    bout = torch.empty((B, max(nb, 1), 14, S), dtype=torch.float64, device=V.device)
    cout = torch.empty((B, max(nc, 1), 2, S), dtype=torch.float64, device=V.device)

    def launch():
        L.check(lib.yhb_stage_device_eval(plan.handle, ctypes.byref(p), vt.data_ptr(), vtprev.data_ptr(), dout.data_ptr(),
                                          bout.data_ptr(), cout.data_ptr(), st))
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    alg = B * S * (48 * nd + 160 * nb)
    ach = alg / (ms * 1e-3) / 1e9
    out = {'kernel': 'k_stage_device_eval', 'bound': 'hbm', 'achieved': ach, 'peak': hbm_gbs, 'unit': 'GB/s',
           'frac': ach / hbm_gbs, 'avg_launch_ms': ms, 'circuits': B, 'algorithmic_bytes_per_launch': alg,
           'working_set_bytes': int(2 * V.nbytes + bout.nbytes + dout.nbytes),
           'note': 'inputs + outputs (%.0f MB) exceed the 126 MB L2; 20 back-to-back launches timed with CUDA events. '
                   'The kernel is bound by fp64 instruction issue (one Gummel-Poon sample with junction charges is ~2 000 '
                   'warp instructions per 32 samples: 25 divisions, 3 exp, 2 tanh, 1-2 pow at the reference\'s operation '
                   'order), not by HBM; inside k_fused the same sweep runs from shared memory.' % (
                       (2 * V.nbytes + bout.nbytes + dout.nbytes) / 1e6)}
    ncu = ncu_summary('k_stage_device_eval')
    if ncu:
        out['ncu'] = ncu
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import circuits
    import yalrf_b200
    from yalrf_b200 import _dist, _engine as E, _lib as L
    from yalrf_b200._flatten import ParamBlock
    from yalrf_b200.Analyses import MultiToneHarmonicBalance

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    lib = L.lib()
    L.check(lib.yhb_set_device(local))

    y = yalrf_b200.YalRF('Differential Amplifier')
    if args.workload == 'activeload':
        h = circuits.diffamp_activeload(y)
    else:
        h = circuits.diffamp(y)
    hb = MultiToneHarmonicBalance('HB1', h['freq'], h['K'])
    hb.pnp_extension = args.workload == 'activeload'
    # the whole job's grid; this rank solves the points b % world == rank
    VB_all, VIN_all = grid_points(GRID * (world if args.scaling == 'weak' else 1), args.workload)
    Btot = len(VB_all)
    mine = _dist.shard_indices(Btot, rank, world)
    VB, VIN = VB_all[mine], VIN_all[mine]

    def sweeps_of(vb, vin):
        return [(h['i2'], 'ac', vin / 2), (h['i4'], 'ac', vin / 2), (h['i3'], 'dc', vb), (h['i5'], 'dc', vb)]
    B = len(VB)

    plan, topo = hb._setup(y)
    pb = ParamBlock(topo, hb.freq, B)
    for dev, field, values in sweeps_of(VB, VIN):
        pb.set(dev, field, values)
    db = E.DeviceBatch(plan, pb)
    lay = _dist.PackLayout(plan, levels=False)
    nmax = _dist.shard_count(Btot, 0, world)
    gathered = torch.empty((world * nmax, lay.R), dtype=torch.float64, device='cuda') if world > 1 else None
    ev = {k: [] for k in ('t0', 'solved', 'done')}

    def step(timed=False):
        if timed:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
        db.solve(hb.options)
        if timed:
            e[1].record()
        if world > 1:  # results only: no collective inside the Newton loop
            packed = _dist.pack_device(db, lay)
            if packed.shape[0] < nmax:
                packed = torch.cat([packed, torch.zeros((nmax - packed.shape[0], lay.R), dtype=packed.dtype, device='cuda')])
            dist.all_gather_into_tensor(gathered, packed)
        if timed:
            e[2].record()
            for k, x in zip(('t0', 'solved', 'done'), e):
                ev[k].append(x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

    def flush_l2():
        flush_buf.fill_(1)

    for _ in range(args.warmup):
        flush_l2()
        step()
    barrier()
    lib.yhb_profile_enable(1)
    lib.yhb_profile_reset()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    for _ in range(args.steps):      # exactly K timed steps; the L2 flush between them is outside the event pairs
        flush_l2()
        step(timed=True)
    barrier()
    ms = sum(a.elapsed_time(b) for a, b in zip(ev['t0'], ev['done']))
    ms_solve = sum(a.elapsed_time(b) for a, b in zip(ev['t0'], ev['solved']))
    ms_comm = sum(a.elapsed_time(b) for a, b in zip(ev['solved'], ev['done']))
    clocks = sampler.stop() if rank == 0 else None
    prof = L.profile_read()
    lib.yhb_profile_enable(0)
    t = torch.tensor([ms, ms_solve, ms_comm], dtype=torch.float64, device='cuda')
    per_rank = None
    if world > 1:
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = {'solve_ms_per_step': [float(x[1]) / args.steps for x in allt],
                    'comm_ms_per_step': [float(x[2]) / args.steps for x in allt],
                    'note': 'device time per step on every rank (CUDA events): DC + Newton kernels, then pack + '
                            'all_gather_into_tensor; comm includes the wait for the slowest rank'}
        ms = max(float(x[0]) for x in allt)
    cnt = torch.stack([db.converged.clamp(min=0).sum(), db.lu_count.sum(), db.newton_iters.sum()]).to(torch.float64)
    if world > 1:
        dist.all_reduce(cnt)
    nconv, lus_all, newton_all = (int(x) for x in cnt.tolist())
    lus = int(db.lu_count.sum().item())          # this rank's own, for its kernel roofline
    value = Btot * args.steps / (ms * 1e-3)

    # ---- end to end: the call a user makes, host arrays in / out (sharded over the ranks for N > 1)
    hb.distributed = world > 1
    hb.gather_root = 0 if world > 1 else None   # the gathered results are read back to the host on rank 0 only
    hb.report_levels = False   # the 768 B per circuit of per-level reports are optional output; parity fetches them below
    sw_all = sweeps_of(VB_all, VIN_all)
    for _ in range(max(1, min(args.warmup, 2))):
        res = hb.run_sweep(y, sw_all)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = hb.run_sweep(y, sw_all)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e = Btot * args.steps / float(tt.item())
    per_circuit_in = sum(getattr(pb, n).nbytes for n in ('tones', 'lin_val', 'src_val', 'diode_par', 'bjt_par', 'cubic_par')) // B
    h2d = int(per_circuit_in * B)                      # this rank's parameter upload
    d2h = 0
    if rank == 0:
        per_circuit_out = (res.V.nbytes + res.Vf.nbytes + res.converged.nbytes + res.newton_iters.nbytes +
                           res.lu_count.nbytes + res.err.nbytes) // Btot
        d2h = int(per_circuit_out * Btot)   # N > 1: rank 0 reads the gathered results of the whole job

    # ---- parity on a seeded random sample of the grid (outside the timed region) + CPU baseline beside it
    parity = cpu = None
    if rank == 0:
        cores = os.cpu_count() or 1
        hb.report_levels = True
        hb.distributed = False
        npts = min(512, max(64, 16 * cores)) if (world == 1 and not args.no_cpu) else 64
        idx = sample_indices(Btot, npts)
        gres = hb.run_sweep(y, sweeps_of(VB_all[idx], VIN_all[idx]))     # the sampled points again, with level reports
        v, cdt, out = cpu_sample(args.workload, VB_all, VIN_all, idx, cores)
        parity = parity_report(gres.V, gres.level_iters, gres.converged, np.arange(len(idx)), out)
        # and the big batch's own phasors at the same points (same circuits, another batch: must agree to rounding)
        parity['batch_vs_sample_rel'] = float(np.max(np.abs(res.V[idx] - gres.V)) / np.max(np.abs(gres.V)))
        hb.gather_root = None
        if world == 1 and not args.no_cpu:
            cpu = {'value': v, 'unit': 'solves/s', 'cores': cores, 'kind': 'port',
                   'sample': '%d of the 4096 grid points (default_rng(0) subset), cold start each, %d worker processes, '
                             '1 BLAS thread each, %.1f s wall' % (npts, cores, cdt)}

    if rank == 0:
        W = plan.W
        launches = sum(n for n, _ in prof.values())
        kind = max(prof, key=lambda k: prof[k][1])
        n_l, ms_l = prof[kind]
        peak = np.zeros(1)
        L.check(lib.yhb_fp64_peak(peak.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), None))
        hbm = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(
            os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
        total_ms = sum(m for _, m in prof.values())
        fl_ref, fl_core, fl_exec = plan.flops()
        kname = {'fused': 'k_fused', 'lu': 'k_solve_core', 'eval': 'k_eval', 'dc': 'k_dc_warp'}.get(kind, kind)
        # the DC kernel runs BESIDE the HB kernel (side stream), so the kernels' own durations add up to more than the step:
        # share_of_step = this kernel's duration / step duration (CUDA events); share_serialised = its share of the sum of all
        # kernel durations, with the DC kernel at its stand-alone duration -- the figure an ncu launch list (serialised
        # launches, profiles/r02_bench_launch_list_summary.txt) shows
        dc_alone = ncu_summary('k_dc_warp')
        dc_alone_ms = (dc_alone or {}).get('gpu_time_us')
        ser_total = sum(m for k, (_, m) in prof.items() if k != 'dc') + (
            dc_alone_ms * 1e-3 * prof['dc'][0] / 2 if dc_alone_ms and 'dc' in prof else prof.get('dc', (0, 0.))[1])
        roof = {'kernel': kname, 'share_of_step': ms_l / ms if ms else None,
                'share_serialised': ms_l / ser_total if ser_total else None,
                'launches': n_l, 'avg_launch_ms': ms_l / max(n_l, 1), 'traffic': None}
        if kind in ('lu', 'fused'):
            per_s = lus * args.steps / (ms_l * 1e-3) / 1e12
            ach_exec, ach_ref, ach_core = per_s * fl_exec, per_s * fl_ref, per_s * fl_core
            roof.update({
                'bound': 'tensor', 'achieved': ach_exec, 'peak': float(peak[0]), 'unit': 'TFLOP/s',
                'frac': ach_exec / float(peak[0]),
                'flops_per_solve_step': {'executed': fl_exec, 'dense_core': fl_core, 'reference': fl_ref},
                'algorithmic_saving': {'achieved_reference_tflops': ach_ref, 'frac_of_peak': ach_ref / float(peak[0]),
                                       'achieved_dense_core_tflops': ach_core,
                                       'note': 'SURVEY 8d algorithmic LU flops (2/3 W^3 + 2 W^2, W=%d: the dense system the '
                                               'reference hands to LAPACK) x factorisations / kernel time.  This is how much '
                                               'of the reference\'s arithmetic the structure (peeling to Wc=%d, block-sparse '
                                               'LU over node blocks) removes -- NOT a hardware utilisation.' % (W, plan.core()[1])},
                'note': 'FP64 pipes (DFMA + DMMA).  achieved = multiply-adds x 2 the kernel EXECUTES in its Newton-step linear '
                        'solves (yhb_plan_flops: block-sparse LU over node blocks incl. Schur updates and substitutions) x '
                        'factorisations / CUDA-event time of the whole launch, which also runs device evaluation, DFTs, '
                        'assembly and peeled rows; peak = DFMA microbenchmark run live (MEASURED_PEAKS.json has no fp64 '
                        'figure).  The kernel is latency bound (dependent-issue chains of the pivot steps), see ncu.'})
            ncu = ncu_summary(kname)
            if ncu:
                roof['ncu'] = ncu
                roof['traffic'] = ncu.get('dram_bytes_per_launch')
        else:
            S = plan.S
            newton = int(db.newton_iters.sum().item())
            bytes_eval = newton * args.steps * S * (48 * len(topo.diodes) + 160 * len(topo.bjts))
            ach = bytes_eval / (ms_l * 1e-3) / 1e9
            roof.update({'bound': 'hbm', 'achieved': ach, 'peak': hbm, 'unit': 'GB/s', 'frac': ach / hbm})
        devroof = device_eval_roofline(plan, topo, db, lib, L, hbm)
        line = {
            'metric': 'HB solves/sec', 'value': value, 'unit': 'solves/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': args.scaling,
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args.workload, world, args.scaling),
            'converged': '%d/%d' % (nconv, Btot), 'newton_iters_per_solve': newton_all / Btot, 'lu_per_solve': lus_all / Btot,
            'e2e': {'value': e2e, 'unit': 'solves/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'path': 'hb.run_sweep (hb.distributed=%s, report_levels=False): pinned host arrays -> H2D -> DC + solve '
                            '-> %sD2H%s' % (world > 1, 'pack + NCCL all_gather_into_tensor -> ' if world > 1 else '',
                                         ' of the whole job\'s results on rank 0 (hb.gather_root = 0)' if world > 1 else '')},
            'gpu_launches': launches, 'kernels': {k: {'launches': n, 'ms': round(m, 3)} for k, (n, m) in prof.items() if n},
            'roofline': roof, 'roofline_device_eval': devroof, 'parity': parity, 'per_rank': per_rank,
            'cpu_baseline': cpu, 'clocks': clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='diffamp', choices=['diffamp', 'activeload'])
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: 64x64 points per GPU; strong: the one 4096-point sweep split over the GPUs')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg (parity still runs on 64 points)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
