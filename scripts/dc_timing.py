"""Phase timing of k_dc from a -DYHB_TIMING build (bench grid: 64x64 DiffAmp cold sweep):
   YHB_OUT=yalrf_b200/libyhb_timing.so bash yalrf_b200/csrc/build.sh -DYHB_TIMING
   YHB_LIB=yalrf_b200/libyhb_timing.so python scripts/dc_timing.py"""
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

y = yalrf_b200.YalRF('Differential Amplifier')
h = circuits.diffamp(y)
hb = MultiToneHarmonicBalance('HB1', h['freq'], h['K'])
vin = np.linspace(100e-6, 50e-3, 64); vb = np.linspace(1.0, 1.4, 64)
VB, VIN = (a.ravel() for a in np.meshgrid(vb, vin, indexing='ij'))
plan, topo = hb._setup(y)
pb = ParamBlock(topo, hb.freq, 4096)
for dev, f, v in [(h['i2'], 'ac', VIN / 2), (h['i4'], 'ac', VIN / 2), (h['i3'], 'dc', VB), (h['i5'], 'dc', VB)]:
    pb.set(dev, f, v)
db = E.DeviceBatch(plan, pb)
lib = L.lib()
st = torch.cuda.current_stream().cuda_stream
def dc():
    L.check(lib.yhb_dc_solve(plan.handle, C.byref(db.params), db.ws.data_ptr(), db.ws_bytes, db.Vdc.data_ptr(),
                             db.dcinfo.data_ptr(), st))
for _ in range(3): dc()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); dc(); e1.record(); torch.cuda.synchronize()
info = db.dcinfo.cpu().numpy()
it = info[:, 1]
print('k_dc: %.3f ms for 4096 circuits; converged %d; newton iterations min/mean/max %d/%.1f/%d; histogram %s'
      % (e0.elapsed_time(e1), info[:, 0].sum(), it.min(), it.mean(), it.max(),
         np.histogram(it, bins=[0, 25, 50, 75, 100, 150, 200, 250, 400, 10000])[0].tolist()))
buf = (C.c_ulonglong * 64)()
if lib.yhb_debug_timing(buf, 1) == 0:
    dc(); torch.cuda.synchronize()
    L.check(lib.yhb_debug_timing(buf, 0))
    t = np.array(list(buf), dtype=np.float64)
    names = {24: 'zero + device evaluation', 25: 'ordered stamping', 26: 'A + Anl', 27: 'linear solve', 28: 'checks + limiter state'}
    for k, nm in names.items():
        print('%-26s %8.0f cycles per Newton iteration' % (nm, t[k] / it.sum()))
    print('total %.0f cycles per iteration' % (sum(t[k] for k in names) / it.sum()))
