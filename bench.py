#!/usr/bin/env python
"""bench.py -- HB solves/sec on BASELINE.json's configuration 4.

Workload (config.workload): the reference-pinned DiffAmp netlist (tests/HB_DiffAmp.py:13-69 of the reference,
resistor loads, 4 Gummel-Poon BJTs with junction capacitances, N=10, K=10, W=210), 64 x 64 grid of
(vbias, vin) points per GPU, every point an independent COLD start (hb.run(y), V0=None): DC operating point,
8+ source-stepping levels, Newton to the reference's default tolerances.  One "step" = one pass over the
whole grid.  The ActiveLoad netlist BASELINE.json names needs PNP devices, which the reference's HB does not
implement; `--workload activeload` runs it with the PNP extension instead.

    python bench.py [--gpus N] [--steps K] [--warmup W]            (torchrun for N > 1, one rank per GPU)
    python bench.py --impl reference ...                           (CPU arm: the oracle port on all host cores)

value  : solves/s with the parameter arrays resident in HBM (DeviceBatch.solve: DC + ladder + Newton on device)
e2e    : solves/s through hb.run_sweep() -> yhb_run_host: host arrays in, host arrays out, copies inside
roofline: the kernel with the largest device time share, from CUDA events recorded around every launch
cpu_baseline: the numpy oracle (oracle/hb_oracle.py, kind "port": the reference is Python and cannot travel)
"""
import argparse
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
# dram__bytes_read.sum + dram__bytes_write.sum of one k_fused launch over the 4096-point grid (ncu --set full, profiles/)
FUSED_TRAFFIC_BYTES = {'diffamp': 49096192, 'activeload': None}


def grid_points(rank, world, workload):
    """(vbias, vin) of this rank's 64 x 64 grid: ranks split a (64 * world) x 64 grid along vbias."""
    if workload == 'activeload':
        vb_all = np.linspace(1.3, 1.7, GRID * world)
        vin = np.linspace(20e-6, 1e-3, GRID)
    else:
        vb_all = np.linspace(1.0, 1.4, GRID * world)
        vin = np.linspace(100e-6, 50e-3, GRID)
    vb = vb_all[rank * GRID:(rank + 1) * GRID]
    VB, VIN = np.meshgrid(vb, vin, indexing='ij')
    return VB.ravel(), VIN.ravel()


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
    return bool(conv), hb.total_newton


def cpu_sample(workload, npoints, cores):
    """Oracle cold solves of `npoints` grid points spread over the sweep, on `cores` worker processes."""
    import multiprocessing as mp
    VB, VIN = grid_points(0, 1, workload)
    idx = np.linspace(0, len(VB) - 1, npoints).astype(int)
    jobs = [(workload, float(VB[i]), float(VIN[i])) for i in idx]
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        pool.map(_oracle_point, jobs[:cores])            # warm the workers (imports)
        t = time.perf_counter()
        out = pool.map(_oracle_point, jobs, chunksize=1)
        dt = time.perf_counter() - t
    return len(jobs) / dt, dt, sum(1 for c, _ in out if c), sum(n for _, n in out)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    npts = min(512, max(64, 16 * cores))
    vals = []
    for s in range(args.warmup + args.steps):
        v, dt, nconv, nnewton = cpu_sample(args.workload, npts, cores)
        if s >= args.warmup:
            vals.append((v, dt))
    value = sum(npts for _ in vals) / sum(dt for _, dt in vals)
    sample = '%d of the 4096 grid points (evenly strided), cold start each, %d worker processes, 1 BLAS thread each' % (
        npts, cores)
    line = {
        'impl': 'reference', 'metric': 'HB solves/sec', 'value': value, 'unit': 'solves/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * statistics.mean(dt for _, dt in vals),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.workload, 1),
        'cpu_baseline': {'value': value, 'unit': 'solves/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'solves/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(workload, world):
    name = ('HB_DiffAmp_ActiveLoad netlist + PNP extension' if workload == 'activeload' else
            'HB_DiffAmp (cfg 4, reference-pinned resistor-load variant of the DiffAmp bias/amplitude sweep)')
    return {'workload': '%s: %dx%d (vbias x vin) cold-start HB solves per GPU, N=10 nodes, K=10 harmonics, W=210, '
                        'reltol 1e-3 / abstol 1e-6 / maxiter 100' % (name, GRID, GRID),
            'points_per_gpu': GRID * GRID, 'points_total': GRID * GRID * world, 'start': 'cold (V0=None, DC on GPU)',
            'parallelism': 'independent circuits sharded over ranks; NCCL all_gather of Vf + converged only',
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


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import circuits
    import yalrf_b200
    from yalrf_b200 import _engine as E, _lib as L
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
    VB, VIN = grid_points(rank, world, args.workload)
    sweeps = [(h['i2'], 'ac', VIN / 2), (h['i4'], 'ac', VIN / 2), (h['i3'], 'dc', VB), (h['i5'], 'dc', VB)]
    B = len(VB)

    plan, topo = hb._setup(y)
    pb = ParamBlock(topo, hb.freq, B)
    for dev, field, values in sweeps:
        pb.set(dev, field, values)
    db = E.DeviceBatch(plan, pb)
    gather = None
    if world > 1:
        gather = [torch.empty_like(db.Vf) for _ in range(world)]
        gconv = [torch.empty_like(db.converged) for _ in range(world)]

    def step():
        db.solve(hb.options)
        if world > 1:  # results only: no collective inside the Newton loop
            dist.all_gather(gather, db.Vf)
            dist.all_gather(gconv, db.converged)

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
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for e0, e1 in evs:      # exactly K timed steps; the L2 flush between them is outside the event pairs
        flush_l2()
        e0.record()
        step()
        e1.record()
    barrier()
    ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    clocks = sampler.stop() if rank == 0 else None
    prof = L.profile_read()
    lib.yhb_profile_enable(0)
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    nconv = int(db.converged.sum().item())
    lus = int(db.lu_count.sum().item())
    newton = int(db.newton_iters.sum().item())
    value = B * world * args.steps / (ms * 1e-3)

    # ---- end to end: the call a user makes, host arrays in / out
    for _ in range(max(1, min(args.warmup, 2))):
        res = hb.run_sweep(y, sweeps)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = hb.run_sweep(y, sweeps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e = B * world * args.steps / float(tt.item())
    pbf = res.params_bytes
    h2d = int(pbf)
    d2h = int(res.V.nbytes + res.Vf.nbytes + res.converged.nbytes + res.newton_iters.nbytes + res.lu_count.nbytes +
              res.level_iters.nbytes + res.level_alpha.nbytes + res.err.nbytes)

    if rank == 0:
        W = plan.W
        launches = sum(n for n, _ in prof.values())
        kind = max(prof, key=lambda k: prof[k][1])
        n_l, ms_l = prof[kind]
        peak = np.zeros(1)
        L.check(lib.yhb_fp64_peak(peak.ctypes.data_as(__import__('ctypes').POINTER(__import__('ctypes').c_double)), None))
        total_ms = sum(m for _, m in prof.values())
        roof = {'kernel': kind, 'share_of_device_time': ms_l / total_ms if total_ms else None,
                'launches': n_l, 'avg_launch_ms': ms_l / max(n_l, 1), 'traffic': None}
        if kind in ('lu', 'fused'):
            wc = res.core_W if hasattr(res, 'core_W') else W
            per_lu = (2.0 / 3.0) * W ** 3 + 2.0 * W ** 2                 # the system the reference factors
            per_lu_exec = (2.0 / 3.0) * wc ** 3 + 2.0 * wc ** 2          # what is left after structural peeling
            ach = lus * args.steps * per_lu / (ms_l * 1e-3) / 1e12
            ach_exec = lus * args.steps * per_lu_exec / (ms_l * 1e-3) / 1e12
            roof.update({'bound': 'tensor', 'achieved': ach, 'peak': float(peak[0]), 'unit': 'TFLOP/s',
                         'frac': ach / float(peak[0]), 'achieved_executed': ach_exec,
                         'traffic': FUSED_TRAFFIC_BYTES.get(args.workload),
                         'note': 'FP64 pipes (DFMA + DMMA).  achieved = SURVEY 8d algorithmic LU flops (2/3 W^3 + 2 W^2, '
                                 'W=%d, the system the reference hands to LAPACK) x factorisations / CUDA-event time of '
                                 'the whole fused kernel (which also does device evaluation, the DFTs, assembly and the '
                                 'peeled rows): the fraction says how much of the reference\'s arithmetic the structure '
                                 'removes, not how busy the pipe is.  achieved_executed counts a dense factorisation of '
                                 'the Wc=%d core that structural peeling leaves; the block-sparse LU over node blocks '
                                 'that k_fused<1> runs on it executes about 1/8 of that again (DiffAmp: 3 + 1 node blocks '
                                 'of 21 pivots on 10 of 16 blocks).  peak = DFMA microbenchmark run live '
                                 '(MEASURED_PEAKS.json has no fp64 figure; DMMA m8n8k4 measured at 36.8 TFLOP/s, '
                                 'profiles/r01_microbench.txt).  ncu (profiles/r01_bench_fused_full.txt): issue slots '
                                 '39 %% busy, FP64 pipe 9.2 %%, DMMA pipe 2.7 %% (transforms), 3 CTAs/SM by shared '
                                 'memory and registers: the kernel is bound by the dependent-issue latency of the '
                                 'factorisation warps.  traffic = dram bytes of one launch from ncu --set full.'
                                 % (W, wc)})
        else:
            hbm = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(
                os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
            S = plan.S
            bytes_eval = newton * args.steps * S * (48 * len(topo.diodes) + 160 * len(topo.bjts))
            ach = bytes_eval / (ms_l * 1e-3) / 1e9
            roof.update({'bound': 'hbm', 'achieved': ach, 'peak': hbm, 'unit': 'GB/s', 'frac': ach / hbm})
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu:
            npts = min(512, max(64, 16 * cores))
            v, cdt, cconv, cnewton = cpu_sample(args.workload, npts, cores)
            cpu = {'value': v, 'unit': 'solves/s', 'cores': cores, 'kind': 'port',
                   'sample': '%d of the 4096 grid points (evenly strided), cold start each, %d worker processes, '
                             '1 BLAS thread each, %.1f s wall' % (npts, cores, cdt)}
        line = {
            'metric': 'HB solves/sec', 'value': value, 'unit': 'solves/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(args.workload, world),
            'converged': '%d/%d' % (nconv, B), 'newton_iters_per_solve': newton / B, 'lu_per_solve': lus / B,
            'e2e': {'value': e2e, 'unit': 'solves/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': launches, 'kernels': {k: {'launches': n, 'ms': round(m, 3)} for k, (n, m) in prof.items() if n},
            'roofline': roof, 'cpu_baseline': cpu, 'clocks': clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='diffamp', choices=['diffamp', 'activeload'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
