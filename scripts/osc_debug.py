"""Oscillator solve diagnostics: analytic 2x2 Jacobian (sensitivities of the factored HB Jacobian) against finite
differences of the KCL objective, then the full solve.  python scripts/osc_debug.py [peltz|vanderpol]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import circuits, yalrf_b200
from yalrf_b200.Analyses import MultiToneHarmonicBalance
name = sys.argv[1] if len(sys.argv) > 1 else 'peltz'
y = yalrf_b200.Netlist('o'); h = getattr(circuits, name)(y)
hb = MultiToneHarmonicBalance('HB1')
f0, A0 = (71000.0, 0.7) if name == 'peltz' else (h['f0'], h['V0'])
for rel in (1e-4, 1e-6):
    f = np.array([f0, f0 * (1 + rel), f0]); A = np.array([A0, A0, A0 * (1 + rel)])
    r = hb.run_oscillator_batch(y, f, h['K'], A, h['node'], maxiter=1)
    Y = r.Yosc
    fd = np.array([[(Y[1].real - Y[0].real) / (f0 * rel), (Y[2].real - Y[0].real) / (A0 * rel)],
                   [(Y[1].imag - Y[0].imag) / (f0 * rel), (Y[2].imag - Y[0].imag) / (A0 * rel)]])
    print('rel step %g: Y0 = %s, hb_converged %s' % (rel, Y[0], r.hb_converged))
    print(' finite differences:\n', fd)
    print(' analytic Jacobian :\n', r.jacobian[0])
    print(' next step', r.next_step[0], 'gave up', r.gave_up)
    print(' dbg xa[n], xa[nosc2], xf[n], I:', ['%.17g' % v for v in r.dbg[0]])
r = hb.run_oscillator_batch(y, h['f0'], h['K'], h['V0'], h['node'], maxiter=40)
print('solve: fosc %r Vosc %r |Yosc| %g converged %s outer %s newton %d gave_up %s next_step %s' % (
    r.fosc, r.Vosc, r.yosc[0], r.converged, r.outer_iterations, r.newton_iterations, r.gave_up, r.next_step))
