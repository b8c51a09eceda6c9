import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch, circuits, yalrf_b200
from yalrf_b200.Analyses import MultiToneHarmonicBalance
stages, k, cje = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
res = {}
for bt in (True, False):
    y = yalrf_b200.Netlist('ladder'); h = circuits.ladder(y, stages=stages, cje=cje)
    hb = MultiToneHarmonicBalance('HB1', h['freq'], [k, k]); hb.block_thomas = bt
    hb.options['maxiter'] = int(sys.argv[4]) if len(sys.argv) > 4 else 100
    t0 = time.perf_counter(); conv = hb.run(y)[0]; torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(hb._plan.engine(), 'conv', conv, 'levels', hb.last.levels()[:12], 'newton', int(hb.last.newton_iters[0]), '%.2f s' % dt, flush=True)
    res[bt] = hb.V.copy() if conv else None
if res[True] is not None and res[False] is not None:
    print('rel diff', np.max(np.abs(res[True] - res[False])) / np.max(np.abs(res[False])))
