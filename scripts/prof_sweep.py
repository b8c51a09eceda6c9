"""Small DiffAmp cold-start sweep for profiling runs (ncu): python scripts/prof_sweep.py [points] [repeats]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import circuits
import yalrf_b200
from yalrf_b200 import _engine as E
from yalrf_b200._flatten import ParamBlock
from yalrf_b200.Analyses import MultiToneHarmonicBalance

n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
y = yalrf_b200.YalRF('Differential Amplifier')
h = circuits.diffamp(y)
hb = MultiToneHarmonicBalance('HB1', h['freq'], h['K'])
vin = np.linspace(100e-6, 50e-3, n)
vb = np.linspace(1.0, 1.4, n)[np.random.default_rng(0).permutation(n)]
plan, topo = hb._setup(y)
pb = ParamBlock(topo, hb.freq, n)
for dev, f, v in [(h['i2'], 'ac', vin / 2), (h['i4'], 'ac', vin / 2), (h['i3'], 'dc', vb), (h['i5'], 'dc', vb)]:
    pb.set(dev, f, v)
db = E.DeviceBatch(plan, pb)
for r in range(reps):
    torch.cuda.synchronize(); t = time.perf_counter()
    db.solve(hb.options)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print('rep %d: %d solves in %.2f ms (%.0f solves/s), converged %d' % (r, n, dt * 1e3, n / dt, int(db.converged.sum())))
