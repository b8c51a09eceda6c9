"""BASELINE config 1 (tests/HB_Diode.py): single cold solve and a 4096-point I1.ac sweep, for A/B runs of the small-circuit
variants of the fused kernel: python scripts/cfg1_value.py"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import circuits, yalrf_b200
from yalrf_b200 import _engine as E, _lib as L
from yalrf_b200._flatten import ParamBlock
from yalrf_b200.Analyses import MultiToneHarmonicBalance
y = yalrf_b200.YalRF('d'); h = circuits.hb_diode(y)
hb = MultiToneHarmonicBalance('HB1', h['freq'], h['K'])
hb.run(y); torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5): hb.run(y)
torch.cuda.synchronize(); single = (time.perf_counter() - t) / 5
ac = np.linspace(0.5, 5, 4096)
plan, topo = hb._setup(y)
pb = ParamBlock(topo, hb.freq, 4096); pb.set(h['i1'], 'ac', ac)
db = E.DeviceBatch(plan, pb)
for _ in range(3): db.solve(hb.options)
torch.cuda.synchronize()
ms = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); db.solve(hb.options); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
print(json.dumps({'lib': os.environ.get('YHB_LIB', 'libyhb.so'), 'cfg1_single_ms': single * 1e3, 'cfg1_sweep4096_device_ms': ms,
                  'cfg1_sweep4096_solves_per_s': 4096 * 5 / (sum(ms) * 1e-3), 'converged': int((db.converged > 0).sum()),
                  'newton': int(db.newton_iters.sum()), 'Vsum': float(db.V.abs().sum())}))
