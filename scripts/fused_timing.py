"""Phase timing of the fused engine from a -DYHB_TIMING build:
   YHB_OUT=yalrf_b200/libyhb_timing.so bash yalrf_b200/csrc/build.sh -DYHB_TIMING
   YHB_LIB=yalrf_b200/libyhb_timing.so python scripts/fused_timing.py [points]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import circuits
import yalrf_b200
from yalrf_b200 import _engine as E, _lib as L
from yalrf_b200._flatten import ParamBlock
from yalrf_b200.Analyses import MultiToneHarmonicBalance

n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
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
db.solve(hb.options)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 64)()
L.check(L.lib().yhb_debug_timing(buf, 1))
db.solve(hb.options)
torch.cuda.synchronize()
L.check(L.lib().yhb_debug_timing(buf, 0))
t = np.array(list(buf), dtype=np.float64)
newton = float(db.newton_iters.sum()); lus = float(db.lu_count.sum())
print('circuits %d, newton iterations %.0f, solves %.0f' % (n, newton, lus))
names = {0: 'evaluation', 1: 'newton_pre', 2: 'assembly', 3: 'core solve', 4: 'newton_post', 5: 'bookkeeping'}
tot = sum(t[k] for k in names)
for k, nm in names.items():
    per = t[k] / (newton if k in (0, 5) else lus)
    print('%-14s %5.1f%% of CTA time, %8.0f cycles per %s' % (nm, 100 * t[k] / tot, per, 'iteration' if k in (0, 5) else 'solve'))
if plan.block_sparse():
    print('block-sparse core LU (%d levels), cycles per solve: factor level 0 %.0f, deeper levels %.0f, schur %.0f, back substitution %.0f'
          % ((plan.block_sparse(),) + tuple(t[k] / lus for k in (8, 9, 11, 12))))
npanel = lus * 21
if not plan.block_sparse():
  print('per panel (Wc=84, 21 panels): factor %.0f cycles; barrier wait owner %.0f / next owner %.0f / others(sum over warps) %.0f; '
      'update next owner %.0f / others (sum over warps) %.0f' % tuple(t[k] / npanel for k in (8, 10, 11, 12, 13, 14)))
print('evaluation sub-phases, cycles per iteration: idft %.0f, device sweep %.0f, current/waveform sums %.0f, dfts %.0f, residual+test %.0f'
      % tuple(t[k] / newton for k in (16, 17, 18, 19, 20)))
