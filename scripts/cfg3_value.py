"""BASELINE config 3 at [3,3] (two-tone diode mixer, dense core Wc = 49: k_fused<8>): 4096-point amplitude sweep, device-timed,
for A/B runs: python scripts/cfg3_value.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import circuits, yalrf_b200
from yalrf_b200 import _engine as E
from yalrf_b200._flatten import ParamBlock
from yalrf_b200.Analyses import MultiToneHarmonicBalance
y = yalrf_b200.YalRF('d'); h = circuits.hb_diode_two_tones(y)
hb = MultiToneHarmonicBalance('HB1', h['freq'], [3, 3])
plan, topo = hb._setup(y)
B = 4096
ac = np.linspace(0.1, 1.0, B)
pb = ParamBlock(topo, hb.freq, B); pb.set(h['i1'], 'ac', ac); pb.set(h['i2'], 'ac', ac)
db = E.DeviceBatch(plan, pb)
for _ in range(3): db.solve(hb.options)
torch.cuda.synchronize()
ms = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); db.solve(hb.options); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
print(json.dumps({'lib': os.environ.get('YHB_LIB', 'libyhb.so'), 'core': plan.core(), 'two_tone_3x3_sweep4096_solves_per_s': B * 5 / (sum(ms) * 1e-3),
                  'ms': [round(m, 2) for m in ms], 'converged': int((db.converged > 0).sum()), 'Vsum': float(db.V.abs().sum())}))
