"""Small solves that touch every kernel of libyhb, for compute-sanitizer (scripts/sanitize.sh):
   python scripts/sanitize_case.py CASE [CASE ...]     CASE in fused_bs | fused_dense | staged | gjp | thomas | osc | stages"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import circuits
import yalrf_b200
from yalrf_b200.Analyses import MultiToneHarmonicBalance as MTHB



def diffamp_sweep(hb, n):
    y = yalrf_b200.YalRF('Differential Amplifier')
    h = circuits.diffamp(y)
    vin = np.linspace(100e-6, 50e-3, n)
    vb = np.linspace(1.0, 1.4, n)
    return hb.run_sweep(y, [(h['i2'], 'ac', vin / 2), (h['i4'], 'ac', vin / 2), (h['i3'], 'dc', vb), (h['i5'], 'dc', vb)])


for case in sys.argv[1:]:
    if case == 'fused_bs':          # k_prepare, k_order, k_dc_warp<12>, k_init, k_fused<1> (block-sparse compact variant), k_finish
        hb = MTHB('HB1', 10e6, 10)
        r = diffamp_sweep(hb, 6)
        print(case, 'converged', int((r.converged > 0).sum()), 'of', r.B)
    elif case == 'fused_dense':     # k_fused<3> (HB_Diode), k_fused<11> (DiffAmp, dense tensor-core Gauss-Jordan), k_fused<8> (two tones)
        y = yalrf_b200.YalRF('Diode Testbench'); h = circuits.hb_diode(y)
        print(case, 'hb_diode', MTHB('HB1', h['freq'], h['K']).run(y)[0])
        hb = MTHB('HB1', 10e6, 10); hb.block_sparse = False
        r = diffamp_sweep(hb, 3)
        print(case, 'diffamp dense', int((r.converged > 0).sum()), 'of', r.B)
        y = yalrf_b200.YalRF('Diode Testbench'); h = circuits.hb_diode_two_tones(y)
        print(case, 'two tones [3,3]', MTHB('HB1', h['freq'], [3, 3]).run(y)[0])
    elif case == 'staged':          # k_eval, k_assemble_core, k_solve_core (one-CTA elimination, dense and banded), k_dc (general engine)
        os.environ['YHB_DC_ENGINE'] = 'block'
        hb = MTHB('HB1', 10e6, 10); hb.engine = 'staged'
        r = diffamp_sweep(hb, 2)
        print(case, 'diffamp staged', int((r.converged > 0).sum()), 'of', r.B)
        y = yalrf_b200.Netlist('ladder'); h = circuits.ladder(y, stages=6, cje=0., ac=0.5)
        hb = MTHB('HB1', h['freq'], [1, 1]); hb.engine = 'staged'; hb.block_thomas = False
        print(case, 'ladder banded', hb.run(y)[0])
    elif case == 'gjp':             # k_gjp_panel<1>, k_gjp_update, k_gjp_finish on a dense core (two tones [5,5]: Wc = 121 < 128 -> force via [6,6])
        y = yalrf_b200.YalRF('Diode Testbench'); h = circuits.hb_diode_two_tones(y)
        hb = MTHB('HB1', h['freq'], [6, 6])
        print(case, 'two tones [6,6]', hb.run(y)[0], hb._plan.engine(), hb._plan.core())
    elif case == 'thomas':          # k_assemble_btd, k_btd_schur, k_gjp_*, k_btd_back
        y = yalrf_b200.Netlist('ladder'); h = circuits.ladder(y, stages=4, cje=0., ac=0.5)
        hb = MTHB('HB1', h['freq'], [2, 2]); hb.block_thomas = True
        print(case, 'ladder block-Thomas', hb.run(y)[0], hb._plan.engine())
    elif case == 'osc':             # yhb_solve_oscillator: sensitivity solves, k_osc_* kernels
        y = yalrf_b200.Netlist('vdp'); h = circuits.vanderpol(y)
        hb = MTHB('HB1')
        print(case, 'van der pol', hb.run_oscillator(y, h['f0'], h['K'], h['V0'], h['node'])[0], hb.freq)
    elif case == 'stages':          # stage entry points + yhb_to_time
        import subprocess
        print(case, 'pytest rc', subprocess.call([sys.executable, '-m', 'pytest', os.path.join(ROOT, 'tests', 'test_gpu_stages.py'), '-x', '-q', '-m', 'gpu']))
    torch.cuda.synchronize()
    print(case, 'done', flush=True)
