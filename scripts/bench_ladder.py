"""BASELINE config 5: the diode / diode-connected-BJT ladder on the block-Thomas engine.
   python scripts/bench_ladder.py STAGES K1 K2 [cje] [ac] [exact] [batch]
   batch > 1: that many ladders with independently jittered device parameters (seeds 0 .. batch-1) solved as ONE batch --
   SURVEY 8e: a single ladder is one chain of dependent block rows ("replicas only"), scale-out is a batch of ladders
   cje: junction capacitance of the BJT stages (SURVEY: 1e-12; the reference's accumulated dQdV stalls with it, so
   either cje = 0 or exact = 1, the re-zeroed-dQdV extension); ac: amplitude of the two tones (SURVEY: 1.0)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import circuits
import yalrf_b200
from yalrf_b200 import _lib as L
from yalrf_b200.Analyses import MultiToneHarmonicBalance

stages, k1, k2 = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cje = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-12
ac = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
exact = len(sys.argv) > 6 and sys.argv[6] == '1'
batch = int(sys.argv[7]) if len(sys.argv) > 7 else 1
ys = []
for seed in range(batch):
    y = yalrf_b200.Netlist('ladder')
    h = circuits.ladder(y, stages=stages, cje=cje, ac=ac, seed=seed)
    ys.append(y)
y = ys[0]
hb = MultiToneHarmonicBalance('HB1', h['freq'], [k1, k2])
hb.block_thomas = True
hb.exact_jacobian = exact
print('cje %g, ac %g, exact_jacobian %s, batch %d' % (cje, ac, exact, batch), flush=True)
L.lib().yhb_profile_enable(1); L.lib().yhb_profile_reset()
torch.cuda.synchronize(); t0 = time.perf_counter()
if batch == 1:
    conv, freqs, Vf, _, _ = hb.run(y)
    res = hb.last
else:
    res = hb.run_batch(ys)
    conv = bool((res.converged > 0).all())
    Vf = res.Vf[0]
torch.cuda.synchronize(); dt = time.perf_counter() - t0
prof = L.profile_read()
Nc, Wc, fused = hb._plan.core()
S = hb.S
lus = int(res.lu_count.sum()); newton = int(res.newton_iters.sum())
flops = lus * (14.0 / 3.0) * Nc * S ** 3
lu_ms = prof['lu'][1]
print('ladder %d stages, [%d,%d]: S=%d N=%d W=%d, core %d blocks of %d (Wc=%d), engine %s' % (stages, k1, k2, S, hb.N, hb.W, Nc, S, Wc, hb._plan.engine()))
print('final sum|F| = %s' % np.array2string(res.err, precision=3))
print('converged %s (%d of %d) in %.2f s wall (plan + DC + solve); levels of circuit 0 %s; %d Newton iterations, %d factorisations over the batch' % (
    conv, int((res.converged > 0).sum()), batch, dt, res.levels(0), newton, lus))
print('kernels: ' + ', '.join('%s %d launches %.1f ms' % (k, n, m) for k, (n, m) in prof.items() if n))
print('block-Thomas: %.3g flop per factorisation (14/3 Nc S^3), %.2f TFLOP/s over the solve launches (%.3f s per Newton round of %d circuits)' % (
    flops / max(lus, 1), flops / (lu_ms * 1e-3) / 1e12 if lu_ms else 0., lu_ms * 1e-3 / max(1, prof['lu'][0] // 2), batch))
if conv:
    print('|Vf(n%d)| first bins: %s' % (stages, np.array2string(np.abs(Vf[y.get_node_idx('n%d' % stages) - 1][:6]), precision=6)))
