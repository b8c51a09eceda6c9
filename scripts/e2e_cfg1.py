"""Where the end-to-end time of a 4096-point HB_Diode sweep goes: python scripts/e2e_cfg1.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import circuits, yalrf_b200
from yalrf_b200 import _engine as E, _lib as L
from yalrf_b200._flatten import ParamBlock
from yalrf_b200.Analyses import MultiToneHarmonicBalance
y = yalrf_b200.YalRF('d'); h = circuits.hb_diode(y)
hb = MultiToneHarmonicBalance('HB1', h['freq'], h['K'])
ac = np.linspace(0.5, 5, 4096)
for _ in range(3): hb.run_sweep(y, [(h['i1'], 'ac', ac)])
def T():
    torch.cuda.synchronize(); return time.perf_counter()
lib = L.lib()
for rep in range(3):
    t0 = T(); plan, topo = hb._setup(y)
    t1 = T(); pb = ParamBlock(topo, hb.freq, 4096); pb.set(h['i1'], 'ac', ac); pb.finalize()
    t2 = T()
    lib.yhb_profile_enable(1); lib.yhb_profile_reset()
    res = E.run_host(plan, pb, hb.options)
    t3 = T(); prof = L.profile_read(); lib.yhb_profile_enable(0)
    hb.report_levels = False
    t4 = T(); res2 = hb.run_sweep(y, [(h['i1'], 'ac', ac)]); t5 = T()
    hb.report_levels = True
    print('setup %.2f ms, param block %.2f ms, run_host %.2f ms (kernels: %s); run_sweep without level reports %.2f ms' % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), {k: round(m, 2) for k, (n, m) in prof.items() if n}, 1e3 * (t5 - t4)))
