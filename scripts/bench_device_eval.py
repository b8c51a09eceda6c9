"""Device-evaluation stage kernel (yhb_stage_device_eval) on the config-4 shape: effective HBM bandwidth.
Algorithmic bytes per circuit and evaluation = S * (48 D_diode + 160 D_bjt)  (SURVEY 8d).
    python scripts/bench_device_eval.py [circuits] [repeats]"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import circuits, yalrf_b200
from yalrf_b200 import _engine as E, _lib as L
from yalrf_b200._flatten import ParamBlock

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
y = yalrf_b200.YalRF('x')
h = circuits.diffamp(y)
plan, topo = E.get_plan(y, h['freq'], h['K'])
pb = ParamBlock(topo, h['freq'], B).finalize()
t = {n: torch.from_numpy(getattr(pb, n)).cuda() for n in ('tones', 'lin_val', 'src_val', 'diode_par', 'bjt_par', 'cubic_par')}
p = L.Params(); p.batch = B
for n, v in t.items(): setattr(p, n, v.data_ptr())
g = torch.Generator(device='cuda').manual_seed(0)
# plausible bias: bases ~0.75 V, collectors ~3 V, emitters ~0 (per node rows of W = N*S)
vt = torch.rand((B, plan.N, plan.S), dtype=torch.float64, device='cuda', generator=g) * 0.2
vt[:, 3] += 0.7; vt[:, 5] += 0.7; vt[:, 9] += 0.65; vt[:, 6] += 3.0; vt[:, 7] += 3.0
vtp = vt + 0.01 * torch.randn(vt.shape, dtype=torch.float64, device='cuda', generator=g)
nb = len(topo.bjts)
out = torch.empty((B, nb, 14, plan.S), dtype=torch.float64, device='cuda')
lib = L.lib()
def run():
    L.check(lib.yhb_stage_device_eval(plan.handle, C.byref(p), vt.data_ptr(), vtp.data_ptr(), None, out.data_ptr(), None, None))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
alg = B * plan.S * 160 * nb
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
print(json.dumps({'kernel': 'k_stage_device_eval', 'circuits': B, 'bjt_samples': B * nb * plan.S, 'ms': ms,
                  'algorithmic_bytes': alg, 'achieved_gbs': alg / ms / 1e6, 'peak_gbs': peak, 'frac': alg / ms / 1e6 / peak,
                  'samples_per_s': B * nb * plan.S / ms * 1e3, 'nan': bool(torch.isnan(out).any().item())}))
