"""Scratch diagnostics run on the GPU box (not a test)."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import circuits
import yalrf_b200
from yalrf_b200.Analyses import MultiToneHarmonicBalance

g = np.load(os.path.join(ROOT, 'tests/golden/peltz_probe_fixed.npz'))
y = yalrf_b200.Netlist('Oscillator')
circuits.peltz(y)
y.add_iac('P.I', 'P.n1', 'gnd', ac=0.7, freq=71e3)
y.add_gyrator('P.G', 'P.n1', 'P.n2', 'gnd', 'gnd', 1)
y.add_idealharmonicfilter('P.IHF', 'P.n2', 'nb', 71e3)
hb = MultiToneHarmonicBalance('HB1', 71e3, 10)
conv, freqs, Vf, _, _ = hb.run(y)
V = hb.V.reshape(6, 21); G = g['V'].reshape(6, 21)
print('nodes', y.node_idx_to_name)
for n in range(6):
    print(n, 'maxabs diff %.3e' % np.abs(V[n] - G[n]).max(), 'max|V| %.3e' % np.abs(G[n]).max(), 'rows', np.argsort(-np.abs(V[n]-G[n]))[:3])

# timing: cfg4 reference-pinned 64x64 cold sweep
y = yalrf_b200.YalRF('Differential Amplifier')
h = circuits.diffamp(y)
for nside in (8, 32, 64):
    vin = np.linspace(100e-6, 50e-3, nside); vb = np.linspace(1.0, 1.4, nside)
    VB, VIN = np.meshgrid(vb, vin, indexing='ij'); VB = VB.ravel(); VIN = VIN.ravel()
    hb = MultiToneHarmonicBalance('HB1', h['freq'], h['K'])
    sw = [(h['i2'], 'ac', VIN / 2), (h['i4'], 'ac', VIN / 2), (h['i3'], 'dc', VB), (h['i5'], 'dc', VB)]
    for rep in range(2):
        t = time.time(); res = hb.run_sweep(y, sw); dt = time.time() - t
        print('B=%d rep %d: %.3f s  -> %.1f solves/s; converged %d/%d; newton iters min/mean/max %d/%.1f/%d; lus mean %.1f' % (
            len(VB), rep, dt, len(VB) / dt, res.converged.sum(), len(VB), res.newton_iters.min(), res.newton_iters.mean(),
            res.newton_iters.max(), res.lu_count.mean()))
