"""Device-timed solves/s of the bench workload (4096 cold DiffAmp solves) for A/B runs of library builds:
   YHB_LIB=yalrf_b200/libyhb_variant.so python scripts/ab_value.py [steps]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import circuits
import yalrf_b200
from yalrf_b200 import _engine as E, _lib as L
from yalrf_b200._flatten import ParamBlock
from yalrf_b200.Analyses import MultiToneHarmonicBalance
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
y = yalrf_b200.YalRF('Differential Amplifier')
h = circuits.diffamp(y)
hb = MultiToneHarmonicBalance('HB1', h['freq'], h['K'])
VB, VIN = bench.grid_points(64, 'diffamp')
if os.environ.get('AB_DISTINCT'):      # AB_DISTINCT=n: n distinct bias points (4096: no two DC problems alike, a Monte-Carlo-like batch)
    nd = int(os.environ['AB_DISTINCT'])
    VB = np.linspace(1.0, 1.4, nd)[np.random.default_rng(0).permutation(len(VB)) % nd]
plan, topo = hb._setup(y)
pb = ParamBlock(topo, hb.freq, len(VB))
for dev, f, v in [(h['i2'], 'ac', VIN / 2), (h['i4'], 'ac', VIN / 2), (h['i3'], 'dc', VB), (h['i5'], 'dc', VB)]:
    pb.set(dev, f, v)
db = E.DeviceBatch(plan, pb)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(3):
    flush.fill_(1); db.solve(hb.options)
torch.cuda.synchronize()
lib = L.lib()
lib.yhb_profile_enable(1); lib.yhb_profile_reset()
ms = []
for _ in range(steps):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); db.solve(hb.options); b.record(); torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
prof = L.profile_read()
print(json.dumps({'lib': os.environ.get('YHB_LIB', 'libyhb.so'), 'env': {k: v for k, v in os.environ.items() if (k.startswith('YHB_') or k.startswith('AB_')) and k != 'YHB_LIB'},
                  'solves_per_s': len(VB) * steps / (sum(ms) * 1e-3), 'ms': [round(m, 3) for m in ms],
                  'converged': int((db.converged > 0).sum()), 'newton': int(db.newton_iters.sum()), 'lus': int(db.lu_count.sum()),
                  'kernels_ms_per_step': {k: round(m / steps, 3) for k, (n, m) in prof.items() if n},
                  'Vsum': float(db.V.abs().sum())}))
