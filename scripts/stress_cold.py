"""Stress of the cold solve (DC groups beside the HB kernel, deferred list): repeated sweeps of several sizes and bias-point
counts must converge everywhere and be bit-identical from run to run: python scripts/stress_cold.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import circuits, yalrf_b200
from yalrf_b200.Analyses import MultiToneHarmonicBalance
y = yalrf_b200.YalRF('Differential Amplifier'); h = circuits.diffamp(y)
hb = MultiToneHarmonicBalance('HB1', h['freq'], h['K'])
rng = np.random.default_rng(1)
bad = 0
for B, nb in ((1, 1), (2, 2), (7, 3), (33, 33), (100, 5), (1000, 64), (4096, 64), (5000, 500), (4096, 1)):
    vb = np.linspace(1.0, 1.4, nb)[rng.integers(0, nb, B)]
    vin = rng.uniform(100e-6, 50e-3, B)
    sw = [(h['i2'], 'ac', vin / 2), (h['i4'], 'ac', vin / 2), (h['i3'], 'dc', vb), (h['i5'], 'dc', vb)]
    ref = None
    t = time.perf_counter()
    for rep in range(12):
        r = hb.run_sweep(y, sw)
        if ref is None:
            ref = (r.V.copy(), r.converged.copy(), r.newton_iters.copy())
        same = np.array_equal(r.V, ref[0]) and np.array_equal(r.converged, ref[1]) and np.array_equal(r.newton_iters, ref[2])
        if not same or int((r.converged > 0).sum()) != B:
            bad += 1
            print('MISMATCH B=%d rep=%d converged %d same %s' % (B, rep, int((r.converged > 0).sum()), same))
    print('B=%5d, %3d bias points: 12 runs identical, all converged, %.1f ms per run' % (B, nb, (time.perf_counter() - t) / 12 * 1e3), flush=True)
print('stress: %s' % ('OK' if bad == 0 else '%d problems' % bad))
