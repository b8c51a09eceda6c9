"""Timings of the other BASELINE.json configs on the GPU (not the bench line): python scripts/bench_configs.py"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import circuits, yalrf_b200
from yalrf_b200.Analyses import MultiToneHarmonicBalance

def timeit(fn, reps=3):
    fn(); fn(); torch.cuda.synchronize()   # two warm-up calls: plan creation, then the page-locked result buffers
    t = time.perf_counter()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps, out

res = {}
# config 1: HB_Diode single + 4096-point I1.ac sweep
y = yalrf_b200.YalRF('d'); h = circuits.hb_diode(y)
hb = MultiToneHarmonicBalance('HB1', h['freq'], h['K'])
dt, _ = timeit(lambda: hb.run(y)); res['cfg1_single_ms'] = dt * 1e3
ac = np.linspace(0.5, 5, 4096)
dt, r = timeit(lambda: hb.run_sweep(y, [(h['i1'], 'ac', ac)]), reps=5); res['cfg1_sweep4096_solves_per_s'] = 4096 / dt; res['cfg1_converged'] = int(r.converged.sum())
# config 3: two tones [3,3] and [8,8]
for nh in ([3, 3], [8, 8]):
    y = yalrf_b200.YalRF('d'); h = circuits.hb_diode_two_tones(y)
    hb = MultiToneHarmonicBalance('HB1', h['freq'], nh)
    dt, _ = timeit(lambda: hb.run(y), reps=2); res['cfg3_%dx%d_single_s' % tuple(nh)] = dt
    res['cfg3_%dx%d_core' % tuple(nh)] = hb._plan.core()
    for B in (64, 512):   # batch over I1.ac = I2.ac in linspace(0.1, 1) (no batch size named)
        ac = np.linspace(0.1, 1.0, B)
        dt, r = timeit(lambda: hb.run_sweep(y, [(h['i1'], 'ac', ac), (h['i2'], 'ac', ac)]), reps=1)
        res['cfg3_%dx%d_sweep%d_solves_per_s' % (nh[0], nh[1], B)] = B / dt
        res['cfg3_%dx%d_sweep%d_converged' % (nh[0], nh[1], B)] = int(r.converged.sum())
# config 2: Peltz oscillator
y = yalrf_b200.Netlist('o'); h = circuits.peltz(y)
hb = MultiToneHarmonicBalance('HB1')
t = time.perf_counter(); hb.run_oscillator(y, h['f0'], h['K'], h['V0'], h['node']); res['cfg2_peltz_oscillator_s'] = time.perf_counter() - t
res['cfg2_outer_iterations'] = int(hb.last_osc.outer_iterations[0]); res['cfg2_fosc'] = hb.fosc; res['cfg2_Vosc'] = hb.Vosc; res['cfg2_yosc'] = hb.osc_yosc
f0 = np.linspace(70e3, 90e3, 64); V0 = np.linspace(0.05, 0.8, 64)
dt, r = timeit(lambda: hb.run_oscillator_batch(y, f0, h['K'], V0, h['node']), reps=2); res['cfg2_batch64_oscillators_per_s'] = 64 / dt; res['cfg2_batch64_converged'] = int(r.converged.sum())
# config 2, the reference's heaviest oscillator workload: the Colpitts divider-ratio sweep of tests/HBOscillators.ipynb cell 8
# (20 run_oscillator calls, K = 20, 110 s in the reference) as one batched call
y = yalrf_b200.Netlist('o2'); h = circuits.colpitts_nvec(y)
hb = MultiToneHarmonicBalance('HB1'); hb.options['maxiter'] = 100
nvec = np.geomspace(0.05, 0.6, 20)
sw = [(h['c1'], 'C', h['ct'] / (1 - nvec)), (h['c2'], 'C', h['ct'] / nvec)]
dt, r = timeit(lambda: hb.run_oscillator_batch(y, np.full(20, h['f0']), h['K'], h['V0'](nvec), h['node'], sweeps=sw, maxiter=60), reps=2)
res['cfg2_colpitts_nvec20_s'] = dt; res['cfg2_colpitts_nvec20_converged'] = int(r.converged.sum())
res['cfg2_colpitts_nvec20_outer_iterations_max'] = int(r.outer_iterations.max()); res['cfg2_colpitts_nvec20_yosc_max'] = float(np.max(r.yosc))
print(json.dumps(res, indent=1))
