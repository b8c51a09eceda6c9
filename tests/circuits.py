"""Circuit builders shared by the golden generator, the oracle tests and the GPU parity tests.

Each builder takes any object with the reference's ``Netlist.add_*`` API (the reference's
``yalrf.Netlist``, the oracle's ``FlatNetlist`` or the product's ``yalrf_b200.Netlist``) and
returns a dict of the device handles a sweep mutates.  Netlists follow the reference scripts
cited in each docstring (paths relative to the reference checkout).
"""
import numpy as np


def hb_diode(y, ac=5.0):
    """tests/HB_Diode.py:10-21 -- BASELINE config 1."""
    i1 = y.add_iac('I1', 'nx', 'gnd', ac=ac, freq=1e6)
    y.add_gyrator('G1', 'nx', 'n1', 'gnd', 'gnd', 1)
    y.add_resistor('R1', 'n1', 'n2', 100)
    d1 = y.add_diode('D1', 'gnd', 'n2')
    d1.options['Is'] = 1e-15
    d1.options['N'] = 1
    d1.options['Area'] = 1
    return dict(i1=i1, d1=d1, freq=1e6, K=10)


def hb_diode_two_tones(y, ac=1.0):
    """tests/HB_DiodeTwoTones.py:11-22 -- BASELINE config 3 (run with [3,3] or [8,8])."""
    i1 = y.add_iac('I1', 'nx', 'gnd', ac=ac, freq=1.1e6)
    y.add_gyrator('G1', 'nx', 'nz', 'gnd', 'gnd', 1)
    i2 = y.add_iac('I2', 'ny', 'gnd', ac=ac, freq=0.9e6)
    y.add_gyrator('G2', 'ny', 'n1', 'nz', 'gnd', 1)
    y.add_resistor('R1', 'n1', 'n2', 100)
    d1 = y.add_diode('D1', 'n2', 'gnd')
    d1.options['Is'] = 1e-15
    d1.options['N'] = 1
    d1.options['Area'] = 1
    return dict(i1=i1, i2=i2, d1=d1, freq=[1.1e6, 0.9e6])


def peltz(y, c=10e-9):
    """tests/HBOsc_Peltz.py:13-45 / README example -- BASELINE config 2."""
    vcc, r, l, re = 10, 200e3, 0.5e-3, 50e3
    y.add_idc('I1', 'nx', 'gnd', dc=vcc)
    y.add_gyrator('G1', 'nx', 'nvcc', 'gnd', 'gnd', 1)
    y.add_resistor('R1', 'nvcc', 'nb', r)
    y.add_inductor('L1', 'nvcc', 'nb', l)
    c1 = y.add_capacitor('C1', 'nvcc', 'nb', c)
    y.add_resistor('RE', 'ne', 'gnd', re)
    q1 = y.add_bjt('Q1', 'nb', 'nvcc', 'ne')
    q2 = y.add_bjt('Q2', 'nvcc', 'nb', 'ne')
    q1.options['Is'] = 1e-16
    q1.options['Bf'] = 200
    q1.options['Br'] = 1
    q2.options = q1.options.copy()
    return dict(c1=c1, q1=q1, q2=q2, f0=80e3, V0=0.1, K=10, node='nb')


def diffamp(y, vin=60e-3, vbias=1.2):
    """tests/HB_DiffAmp.py:13-69 -- BASELINE config 4, reference-pinned (resistor loads).

    The sweep of tests/HB_DiffAmp.py:113-117 sets I2.ac = I4.ac = vin / 2.
    """
    ibias, vcc, rload = 50e-3, 5, 50
    y.add_idc('I1', 'nx', 'gnd', dc=vcc)
    y.add_gyrator('G1', 'nx', 'nvcc', 'gnd', 'gnd', 1)
    i2 = y.add_iac('I2', 'ny', 'gnd', ac=vin, phase=0, freq=10e6)
    i3 = y.add_idc('I3', 'ny', 'gnd', dc=vbias)
    y.add_gyrator('G2', 'ny', 'nb1', 'gnd', 'gnd', 1)
    i4 = y.add_iac('I4', 'nz', 'gnd', ac=vin, phase=+180, freq=10e6)
    i5 = y.add_idc('I5', 'nz', 'gnd', dc=vbias)
    y.add_gyrator('G3', 'nz', 'nb2', 'gnd', 'gnd', 1)
    y.add_resistor('R1', 'nvcc', 'nc1', rload)
    y.add_resistor('R2', 'nvcc', 'nc2', rload)
    q1 = y.add_bjt('Q1', 'nb1', 'nc1', 'ne')
    q2 = y.add_bjt('Q2', 'nb2', 'nc2', 'ne')
    q3 = y.add_bjt('Q3', 'nbx', 'ne', 'gnd')
    q4 = y.add_bjt('Q4', 'nbx', 'nbx', 'gnd')
    y.add_idc('I6', 'nbx', 'gnd', dc=ibias)
    o = q1.options
    o['Is'] = 8.11e-14
    o['Nf'] = 1
    o['Nr'] = 1
    o['Ikf'] = 0.5
    o['Ikr'] = 0.226
    o['Vaf'] = 113
    o['Var'] = 24
    o['Ise'] = 1.06e-11
    o['Ne'] = 2
    o['Isc'] = 0
    o['Nc'] = 2
    o['Bf'] = 205
    o['Br'] = 4
    o['Cje'] = 2.95e-11
    o['Cjc'] = 1.52e-11
    o['Cjs'] = 0.
    q2.options = q1.options
    q3.options = q1.options
    q4.options = q1.options
    return dict(i2=i2, i3=i3, i4=i4, i5=i5, q=[q1, q2, q3, q4], freq=10e6, K=10)


def diffamp_activeload(y, vin=60e-3, vbias=1.5):
    """tests/HB_DiffAmp_ActiveLoad.py:17-54 -- BASELINE config 4 netlist (PNP active loads).

    The reference runs only DC on it; its HB treats PNP as NPN and diverges.
    Usable for HB only with the PNP polarity extension.
    """
    ibias, vcc = 10e-3, 5
    y.add_idc('I1', 'nx', 'gnd', dc=vcc)
    y.add_gyrator('G1', 'nx', 'nvcc', 'gnd', 'gnd', 1)
    i2 = y.add_iac('I2', 'ny', 'gnd', ac=vin, phase=0, freq=10e6)
    i3 = y.add_idc('I3', 'ny', 'gnd', dc=vbias)
    y.add_gyrator('G2', 'ny', 'nb1', 'gnd', 'gnd', 1)
    i4 = y.add_iac('I4', 'nz', 'gnd', ac=vin, phase=+180, freq=10e6)
    i5 = y.add_idc('I5', 'nz', 'gnd', dc=vbias)
    y.add_gyrator('G3', 'nz', 'nb2', 'gnd', 'gnd', 1)
    q1 = y.add_bjt('Q1', 'nb1', 'nc1', 'ne')
    q2 = y.add_bjt('Q2', 'nb2', 'nc2', 'ne')
    q5 = y.add_bjt('Q5', 'nc1', 'nc1', 'nvcc', ispnp=True)
    q6 = y.add_bjt('Q6', 'nc1', 'nc2', 'nvcc', ispnp=True)
    q3 = y.add_bjt('Q3', 'nbx', 'ne', 'gnd')
    q4 = y.add_bjt('Q4', 'nbx', 'nbx', 'gnd')
    y.add_idc('I6', 'nbx', 'gnd', dc=ibias)
    q1.options['Is'] = 1e-15
    q1.options['Bf'] = 100
    q1.options['Vaf'] = 60
    for q in (q2, q3, q4, q5, q6):
        q.options = q1.options
    return dict(i2=i2, i3=i3, i4=i4, i5=i5, q=[q1, q2, q3, q4, q5, q6], freq=10e6, K=10)


def hb_bjt(y):
    """tests/HB_BJT.py:10-30 -- single common-emitter stage (junction capacitances off)."""
    f = 10e6
    y.add_iac('I1', 'n1', 'gnd', ac=10e-3, freq=f)
    y.add_idc('I2', 'n1', 'gnd', dc=10e-3)
    y.add_idc('I3', 'n2', 'gnd', dc=20e-3)
    y.add_resistor('R1', 'n1', 'gnd', 1e3)
    y.add_resistor('R2', 'n1', 'nb', 1e3)
    y.add_resistor('R3', 'n2', 'gnd', 100)
    y.add_resistor('R4', 'n2', 'nc', 1e3)
    y.add_resistor('R5', 'ne', 'gnd', 1e3)
    q1 = y.add_bjt('Q1', 'nb', 'nc', 'ne')
    q1.options['Is'] = 1e-15
    q1.options['Bf'] = 100
    q1.options['Br'] = 1
    return dict(q1=q1, freq=f, K=10)


def vanderpol(y, alpha=1):
    """tests/HBOsc_VanDerPol.py:10-21 -- cubic nonlinearity oscillator."""
    y.add_inductor('L', 'np', 'gnd', 1)
    y.add_capacitor('C', 'np', 'gnd', 1)
    r = y.add_resistor('R', 'np', 'gnd', -1 / alpha)
    x = y.add_cubicnl('X1', 'np', 'gnd', alpha)
    return dict(r=r, x=x, f0=0.13, V0=1.2, K=20, node='np')


def ladder(y, stages=8, cje=0.0, seed=0, ac=1.0):
    """BASELINE config 5: two-tone source stack into a diode / diode-connected-BJT ladder."""
    rng = np.random.default_rng(seed)
    y.add_iac('I1', 'nx', 'gnd', ac=ac, freq=1.1e6)
    y.add_gyrator('G1', 'nx', 'nz', 'gnd', 'gnd', 1)
    y.add_iac('I2', 'ny', 'gnd', ac=ac, freq=0.9e6)
    y.add_gyrator('G2', 'ny', 'n0', 'nz', 'gnd', 1)
    devs = []
    for i in range(1, stages + 1):
        y.add_resistor('R%d' % i, 'n%d' % (i - 1), 'n%d' % i, 100)
        jitter = float(10.0 ** rng.uniform(np.log10(0.9), np.log10(1.1)))
        if i % 2 == 1:
            d = y.add_diode('D%d' % i, 'n%d' % i, 'gnd')
            d.options['Is'] = 1e-15 * jitter
        else:
            d = y.add_bjt('Q%d' % i, 'n%d' % i, 'n%d' % i, 'gnd')
            d.options['Is'] = 1e-15 * jitter
            d.options['Bf'] = 100
            d.options['Cje'] = cje
            # Itf = 0 (default) makes If / (If + Itf) = 0/0 = NaN at the zero-bias start (BJT.py:556): the
            # reference itself cannot start this ladder without it.  Tf = 0, so Itf changes nothing else.
            d.options['Itf'] = 1e-3
        devs.append(d)
    return dict(devs=devs, freq=[1.1e6, 0.9e6])


def current_mirror(y):
    """tests/HB_CurrentMirror.py:10-24 (K = 5): two BJTs sharing one options dict?  No -- only Q1 is configured."""
    y.add_iac('I1', 'n1', 'gnd', ac=3e-3, freq=1e6, phase=+90)
    y.add_idc('I2', 'n1', 'gnd', dc=20e-3)
    y.add_resistor('R1', 'n1', 'gnd', 200)
    y.add_resistor('R2', 'n1', 'nc', 200)
    y.add_idc('I3', 'nb', 'gnd', dc=10e-3)
    q1 = y.add_bjt('Q1', 'nb', 'nc', 'gnd')
    q2 = y.add_bjt('Q2', 'nb', 'nb', 'gnd')
    q1.options['Is'] = 1e-15
    q1.options['Bf'] = 100
    q1.options['Vaf'] = 20
    # Q2 keeps the defaults, whose Itf = 0 makes If / (If + Itf) = 0/0 when Vbe == 0 exactly; the DC point is not 0
    return dict(q1=q1, q2=q2, freq=1e6, K=5)


def fullwave_rectifier(y):
    """tests/HB_FullWaveRectifier.py:22-48: four diodes, floating source (both source nodes non-ground)."""
    y.add_iac('I1', 'vinp', 'vinn', ac=15e-3, phase=+90, freq=1e6)
    y.add_idc('I2', 'vinp', 'vinn', dc=5e-3)
    y.add_resistor('R1', 'vinp', 'vinn', 100)
    y.add_resistor('R2', 'voutp', 'gnd', 100)
    ds = [y.add_diode('D1', 'vinp', 'voutp'), y.add_diode('D2', 'vinn', 'voutp'),
          y.add_diode('D3', 'gnd', 'vinp'), y.add_diode('D4', 'gnd', 'vinn')]
    for d in ds:
        d.options['Is'] = 1e-15
        d.options['N'] = 1
        d.options['Area'] = 1
    return dict(d=ds, freq=1e6, K=10)


def am_demod(y):
    """tests/HB_AMDemod.py:10-26: two-tone envelope detector with a capacitor (run at [3,3] here)."""
    y.add_iac('I1', 'nx', 'gnd', ac=5, freq=1e6)
    y.add_gyrator('G1', 'nx', 'nz', 'gnd', 'gnd', 1)
    y.add_iac('I2', 'ny', 'gnd', ac=0.5, freq=10e3)
    y.add_gyrator('G2', 'ny', 'n1', 'nz', 'gnd', 1)
    y.add_resistor('R1', 'n1', 'nd', 50)
    y.add_resistor('R2', 'n2', 'gnd', 5e3)
    y.add_capacitor('C1', 'n2', 'gnd', 2.2e-9)
    d1 = y.add_diode('D1', 'nd', 'n2')
    d1.options['Is'] = 1e-15
    d1.options['N'] = 1
    d1.options['Area'] = 1
    return dict(d1=d1, freq=[1e6, 10e3])


def colpitts_cb(y):
    """tests/HBOsc_ColpittsCB.py:13-66: common-base Colpitts with the full charge model (Tf, Tr, Mje = 0.5 ...)."""
    vcc, vee, re, l, c1, c2 = 3, -3, 2.2e3, 5e-6, 200e-12, 50e-12
    y.add_idc('I1', 'nx1', 'gnd', dc=vcc)
    y.add_gyrator('G1', 'nx1', 'nvcc', 'gnd', 'gnd', 1)
    y.add_idc('I2', 'nx2', 'gnd', dc=vee)
    y.add_gyrator('G2', 'nx2', 'nvee', 'gnd', 'gnd', 1)
    y.add_resistor('Re', 'ne', 'nvee', re)
    y.add_resistor('Rx', 'nvcc', 'nc', 8.73e3)
    y.add_inductor('L1', 'nvcc', 'nc', l)
    y.add_capacitor('C1', 'ne', 'nc', c1)
    y.add_capacitor('C2', 'nvcc', 'ne', c2)
    q1 = y.add_bjt('Q1', 'gnd', 'nc', 'ne')
    q1.options.update({'Is': 10.2e-15, 'Bf': 301, 'Br': 4, 'Ne': 2, 'Nc': 2, 'Ise': 5.82e-12, 'Vaf': 121, 'Var': 24,
                       'Ikf': 60.7e-3, 'Ikr': 0.15, 'Cje': 26.8e-12, 'Vje': 1.1, 'Mje': 0.5, 'Cjc': 8.67e-12,
                       'Vjc': 0.3, 'Mjc': 0.3, 'Xcjc': 1, 'Cjs': 0, 'Vjs': 0.75, 'Mjs': 0, 'Fc': 0.5,
                       'Tf': 427e-12, 'Tr': 50.3e-9})
    return dict(q1=q1, f0=9e6, V0=1, K=20, node='nc')


def colpitts_nvec(y, n=0.3):
    """tests/HBOscillators.ipynb cell 8 ('Oscillator Parameter Sweep'): Colpitts oscillator whose capacitive divider ratio n
    the notebook sweeps over np.geomspace(0.05, 0.6, 20) in a loop around run_oscillator (its heaviest workload)."""
    vcc, vbias, ibias, rl, fosc, l = 5, 1, 1e-3, 850, 50e6, 50e-9
    ct = 1 / (l * (2 * np.pi * fosc) ** 2)
    y.add_idc('I1', 'nx1', 'gnd', dc=vcc)
    y.add_gyrator('G1', 'nx1', 'nvcc', 'gnd', 'gnd', 1)
    y.add_idc('I2', 'ny1', 'gnd', dc=vbias)
    y.add_gyrator('G2', 'ny1', 'nb', 'gnd', 'gnd', 1)
    y.add_idc('I3', 'gnd', 'ne', dc=ibias)
    y.add_resistor('Rl', 'nvcc', 'nc', rl)
    y.add_inductor('L1', 'nvcc', 'nc', l)
    c1 = y.add_capacitor('C1', 'nc', 'ne', ct / (1 - n))
    c2 = y.add_capacitor('C2', 'ne', 'gnd', ct / n)
    q1 = y.add_bjt('Q1', 'nb', 'nc', 'ne')
    q1.options['Is'] = 1e-15
    q1.options['Bf'] = 100
    return dict(q1=q1, c1=c1, c2=c2, ct=ct, f0=1 / (2 * np.pi * np.sqrt(l * ct)), V0=lambda n_: 2 * ibias * rl * (1 - n_),
                K=20, node='nc', rl=rl, ibias=ibias)


def add_probe(y, name, node, f, amp):
    """The oscillator probe of run_oscillator (MultiToneHarmonicBalance.py:146-150) at a fixed (f, amplitude)."""
    y.add_iac(name + '.I.Oscprobe', name + '.nosc1', 'gnd', ac=amp, freq=f)
    y.add_gyrator(name + '.G.Oscprobe', name + '.nosc1', name + '.nosc2', 'gnd', 'gnd', 1)
    return y.add_idealharmonicfilter(name + 'IHF.Oscprobe', name + '.nosc2', node, f)


def ring(y, n=4, ac=0.5, seed=1):
    """Not a reference script: a ring of n nodes (resistors around, a diode to ground at every node) behind a source
    resistor.  Eliminating a ring node couples its two neighbours: the smallest core whose block-sparse LU has FILL."""
    rng = np.random.default_rng(seed)
    y.add_iac('I1', 'nx', 'gnd', ac=ac, freq=1e6)
    y.add_gyrator('G1', 'nx', 'ns', 'gnd', 'gnd', 1)
    y.add_resistor('Rs', 'ns', 'n1', 50)
    for i in range(1, n + 1):
        y.add_resistor('R%d' % i, 'n%d' % i, 'n%d' % (i % n + 1), 100 + 10 * i)
        d = y.add_diode('D%d' % i, 'n%d' % i, 'gnd')
        d.options['Is'] = 1e-15 * float(10.0 ** rng.uniform(-0.05, 0.05))
    return dict(freq=1e6, K=4)


def stiff_pair(y, ac=3.0, rc=1e-3, rleak=1e9):
    """Not a reference script: an adversarial core for pivoting INSIDE node blocks.  Two diode nodes tied together by a
    1 mOhm resistor (1e3 S on both diagonal blocks and between them) and to everything else only through leakage: each
    diagonal block alone is well conditioned, but the pair's difference mode is 12 decades below its common mode, so
    a factorisation that eliminates node a inside its own block has to carry the cancellation through the Schur
    complement of node b -- LAPACK on the dense matrix would interleave the rows of both nodes.  Strong drive: the
    diode conductance waveforms span many decades."""
    y.add_iac('I1', 'nx', 'gnd', ac=ac, freq=1e6)
    y.add_gyrator('G1', 'nx', 'ns', 'gnd', 'gnd', 1)
    y.add_resistor('Rs', 'ns', 'na', 100)
    y.add_resistor('Rc', 'na', 'nb', rc)
    y.add_resistor('Ra', 'na', 'gnd', rleak)
    y.add_resistor('Rb', 'nb', 'nc', 50)
    y.add_resistor('Rl', 'nc', 'gnd', rleak)
    da = y.add_diode('Da', 'na', 'gnd')
    db = y.add_diode('Db', 'gnd', 'nb')
    dc = y.add_diode('Dc', 'nc', 'gnd')
    da.options['Is'] = 1e-15
    db.options['Is'] = 3e-15
    dc.options['Is'] = 1e-14
    return dict(freq=1e6, K=6)
