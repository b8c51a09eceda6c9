"""Peltz-sized circuits (dense core Wc = 42: k_fused<6>): a 4096-point capacitor sweep of the Peltz oscillator with its probe
fixed, device-timed, and the 64-start-point oscillator batch; for A/B runs: python scripts/cfg2_value.py"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import circuits, yalrf_b200
from yalrf_b200 import _engine as E
from yalrf_b200._flatten import ParamBlock
from yalrf_b200.Analyses import MultiToneHarmonicBalance
y = yalrf_b200.Netlist('o'); h = circuits.peltz(y)
circuits.add_probe(y, 'HB1', h['node'], 71e3, 0.7)
hb = MultiToneHarmonicBalance('HB1', 71e3, h['K'])
plan, topo = hb._setup(y)
B = 4096
pb = ParamBlock(topo, hb.freq, B); pb.set(h['c1'], 'C', np.linspace(8e-9, 12e-9, B))
db = E.DeviceBatch(plan, pb)
for _ in range(3): db.solve(hb.options)
torch.cuda.synchronize()
ms = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); db.solve(hb.options); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
y2 = yalrf_b200.Netlist('o'); h2 = circuits.peltz(y2)
hb2 = MultiToneHarmonicBalance('HB1')
f0 = np.linspace(70e3, 90e3, 64); V0 = np.linspace(0.05, 0.8, 64)
hb2.run_oscillator_batch(y2, f0, h2['K'], V0, h2['node']); torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(3): r = hb2.run_oscillator_batch(y2, f0, h2['K'], V0, h2['node'])
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 3
print(json.dumps({'lib': os.environ.get('YHB_LIB', 'libyhb.so'), 'engine': plan.engine(), 'core': plan.core(),
                  'peltz_sweep4096_solves_per_s': B * 5 / (sum(ms) * 1e-3), 'ms': [round(m, 2) for m in ms],
                  'converged': int((db.converged > 0).sum()), 'Vsum': float(db.V.abs().sum()),
                  'osc_batch64_per_s': 64 / dt, 'osc_converged': int(r.converged.sum())}))
