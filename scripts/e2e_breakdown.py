"""Where the end-to-end time of hb.run_sweep() goes (DiffAmp 64x64 cold sweep): python scripts/e2e_breakdown.py"""
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

y = yalrf_b200.YalRF('Differential Amplifier')
h = circuits.diffamp(y)
hb = MultiToneHarmonicBalance('HB1', h['freq'], h['K'])
vin = np.linspace(100e-6, 50e-3, 64); vb = np.linspace(1.0, 1.4, 64)
VB, VIN = (a.ravel() for a in np.meshgrid(vb, vin, indexing='ij'))
sweeps = [(h['i2'], 'ac', VIN / 2), (h['i4'], 'ac', VIN / 2), (h['i3'], 'dc', VB), (h['i5'], 'dc', VB)]
for _ in range(3):
    res = hb.run_sweep(y, sweeps)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = T(); plan, topo = hb._setup(y)
    t1 = T(); pb = ParamBlock(topo, hb.freq, 4096)
    for dev, f, v in sweeps: pb.set(dev, f, v)
    pb.finalize()
    t2 = T(); r = E.BatchResult(plan, 4096, pinned=True)
    t3 = T(); res = E.run_host(plan, pb, hb.options)
    t4 = T(); res2 = hb.run_sweep(y, sweeps)
    t5 = T()
    print('setup %.2f ms, param block %.2f ms, result alloc %.2f ms, run_host (incl. its own result alloc) %.2f ms; run_sweep total %.2f ms'
          % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)))
