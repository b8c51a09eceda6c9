"""The reference's Colpitts divider-ratio sweep (tests/HBOscillators.ipynb cell 8: a loop of 20 run_oscillator calls, 110 s
there) as ONE batched GPU call: python scripts/osc_nvec.py"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import circuits, yalrf_b200
from yalrf_b200.Analyses import MultiToneHarmonicBalance
y = yalrf_b200.Netlist('Oscillator Parameter Sweep')
h = circuits.colpitts_nvec(y)
hb = MultiToneHarmonicBalance('HB1')
hb.options['maxiter'] = 100
nvec = np.geomspace(0.05, 0.6, 20)
sweeps = [(h['c1'], 'C', h['ct'] / (1 - nvec)), (h['c2'], 'C', h['ct'] / nvec)]
V0 = h['V0'](nvec)
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    r = hb.run_oscillator_batch(y, np.full(20, h['f0']), h['K'], V0, h['node'], sweeps=sweeps, maxiter=60)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
nc = r.netlist.get_node_idx('nc') - 1
print(json.dumps({'seconds': dt, 'converged': int(r.converged.sum()), 'outer_iterations': r.outer_iterations.tolist(),
                  'fosc_MHz': (r.fosc / 1e6).round(4).tolist(), 'Vosc': r.Vosc.round(5).tolist(),
                  'vtank': np.abs(r.Vf[:, nc, 1]).round(5).tolist(), 'yosc_max': float(np.max(r.yosc)),
                  'engine': hb._plan.engine() if hasattr(hb, '_plan') else None}))
