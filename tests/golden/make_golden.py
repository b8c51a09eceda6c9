"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    OMP_NUM_THREADS=1 python tests/golden/make_golden.py

The reference imports matplotlib at module top (MultiToneHarmonicBalance.py:16) and it is
not installed, so a no-op stub package is put first on sys.path.  Nothing here is used at
test time except the .npz files it writes.
"""
import contextlib
import io
import os
import re
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))          # tests/
REF = os.environ.get('YALRF_REFERENCE', '/root/reference')


def import_reference():
    stub = tempfile.mkdtemp(prefix='mplstub')
    os.makedirs(os.path.join(stub, 'matplotlib'))
    for f in ('__init__.py', 'pyplot.py'):
        with open(os.path.join(stub, 'matplotlib', f), 'w') as fh:
            fh.write('def __getattr__(name):\n    return lambda *a, **k: None\n')
    sys.path.insert(0, stub)
    sys.path.insert(0, REF)
    warnings.filterwarnings('ignore')
    import logging
    import yalrf  # noqa
    from yalrf import Netlist
    from yalrf.Analyses import MultiToneHarmonicBalance
    logging.getLogger('YalRF.Analyses.DC').setLevel(logging.ERROR)
    import yalrf.Analyses.DC as _dc  # noqa
    return Netlist, MultiToneHarmonicBalance


def parse_levels(text):
    """[(alpha, iterations, total_error)] from the reference's prints (:354, :597-598)."""
    alphas = [float(x) for x in re.findall(r'Alpha level: (\S+)', text)]
    its = [int(x) for x in re.findall(r'NR number of iterations: (\d+)', text)]
    errs = [float(x) for x in re.findall(r'HB total error: (\S+)', text)]
    if len(alphas) < len(its):          # a direct (V0 given) attempt has no alpha line
        alphas = [1.0] * (len(its) - len(alphas)) + alphas
    return np.array(alphas), np.array(its), np.array(errs)


def quiet_run(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), np.errstate(all='ignore'):
        out = fn(*a, **k)
    return out, buf.getvalue()


class Recorder:
    """Records (J, F, dV) of every LU the reference performs (scipy.linalg.lu_factor / lu_solve, :620-621)."""

    def __init__(self, W=None):
        import scipy.linalg
        self.sl = scipy.linalg
        self.J, self.F, self.dV = [], [], []
        self.on = False
        self.W = W      # only W x W systems are HB Jacobians (the DC analysis uses the same LAPACK entry points)

    def __enter__(self):
        self._f, self._s = self.sl.lu_factor, self.sl.lu_solve
        rec = self

        def lu_factor(J, *a, **k):
            if rec.on and J.shape[0] == rec.W:
                rec.J.append(np.array(J, copy=True))
            return rec._f(J, *a, **k)

        def lu_solve(lu, F, *a, **k):
            x = rec._s(lu, F, *a, **k)
            if rec.on and F.shape[0] == rec.W:
                rec.F.append(np.array(F, copy=True))
                rec.dV.append(np.array(x, copy=True))
            return x
        self.sl.lu_factor, self.sl.lu_solve = lu_factor, lu_solve
        return self

    def __exit__(self, *a):
        self.sl.lu_factor, self.sl.lu_solve = self._f, self._s


def hb_W(hb, y):
    if isinstance(hb.freq, float):
        K = hb.numharmonics
    else:
        K = (2 * hb.numharmonics[1] + 1) * hb.numharmonics[0] + hb.numharmonics[1]
    return (2 * K + 1) * (y.get_num_nodes() - 1)


def extra_fixtures(Netlist, MTHB, save):
    """Second batch of fixtures (more reference scripts): python tests/golden/make_golden.py --extra"""
    import circuits
    for tag, build, nh in (('current_mirror', circuits.current_mirror, None),
                           ('fullwave_rectifier', circuits.fullwave_rectifier, None),
                           ('am_demod_3x3', circuits.am_demod, [3, 3])):
        y = Netlist(tag)
        h = build(y)
        hb = MTHB('HB1', h['freq'], nh if nh is not None else h['K'])
        (conv, freqs, Vf, _, _), text = quiet_run(hb.run, y)
        al, its, errs = parse_levels(text)
        save(tag, converged=conv, freqs=freqs if conv else np.zeros(1), Vf=Vf if conv else np.zeros(1),
             V=hb.V if conv else np.zeros(1), alphas=al, iters=its, errs=errs)
    # Colpitts CB with the probe fixed at (9.5 MHz, 0.8 V): full BJT charge model (Tf, Tr, Cje, Cjc) in hb_loop
    y = Netlist('colpitts')
    h = circuits.colpitts_cb(y)
    circuits.add_probe(y, 'HB1', 'nc', 9.5e6, 0.8)
    hb = MTHB('HB1', 9.5e6, 8)
    with Recorder(hb_W(hb, y)) as rec:
        rec.on = True
        (conv, freqs, Vf, _, _), text = quiet_run(hb.run, y)
    al, its, errs = parse_levels(text)
    save('colpitts_probe_fixed', converged=conv, freqs=freqs if conv else np.zeros(1), Vf=Vf if conv else np.zeros(1),
         V=hb.V if conv else np.zeros(1), alphas=al, iters=its, errs=errs, J0=rec.J[0], F0=rec.F[0], J3=rec.J[3],
         F3=rec.F[3])


# the oscillation point the GPU engine's 2 x 2 Newton iteration finds for the Peltz oscillator (deterministic; the test
# checks that the engine still lands there before it uses the fixture)
PELTZ_GPU_ROOT = (71092.6637049824, 0.7501097624413049)


def osc_root_fixture(Netlist, MTHB, save):
    """the REFERENCE's oscillator objective evaluated at the GPU engine's answer.  The reference's own HB
    solves the Peltz oscillator with its probe fixed at the GPU root (cold start, like the first objective evaluation
    of run_oscillator, :171), and Yosc is formed exactly as objFunc does (:183-186).
        python tests/golden/make_golden.py --osc-root"""
    import circuits
    f, v = PELTZ_GPU_ROOT
    y = Netlist('Oscillator')
    h = circuits.peltz(y)
    circuits.add_probe(y, 'HB1', h['node'], f, v)
    hb = MTHB('HB1', f, h['K'])
    (conv, freqs, Vf, _, _), text = quiet_run(hb.run, y)
    al, its, errs = parse_levels(text)
    n1 = hb.get_node_idx(h['node'])
    n2 = hb.get_node_idx('HB1.nosc2')
    Iosc = (Vf[n1, 1] - Vf[n2, 1]) * 1e9           # Zoscprobe.g (IdealHarmonicFilter.py:12)
    Yosc = Iosc / Vf[n1, 1]
    save('peltz_ref_at_gpu_root', converged=conv, fosc=f, Vosc=v, freqs=freqs, Vf=Vf, V=hb.V, iters=its, yosc=Yosc,
         ulp_floor=1e9 * np.spacing(abs(Vf[n1, 1])) / abs(Vf[n1, 1]))
    print('reference |Yosc| at the GPU root: %.3e S (one ulp of the probe voltage = %.3e S)' % (
        abs(Yosc), 1e9 * np.spacing(abs(Vf[n1, 1])) / abs(Vf[n1, 1])))


def save_npz(name, **arrs):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrs)
    print('wrote', path, {k: np.asarray(v).shape for k, v in arrs.items()})


def colpitts_nvec_fixture(Netlist, MTHB, save):
    """tests/HBOscillators.ipynb cell 8: the 'single iteration' run_oscillator of the Colpitts divider-ratio sweep (n = 0.3)
    and two more points of its nvec loop, each a cold run_oscillator of the unmodified reference.
        python tests/golden/make_golden.py --colpitts-nvec"""
    import circuits
    nvec = np.array([0.05, 0.3, 0.6])
    fosc, vosc, vtank, Vfs = [], [], [], []
    for n in nvec:
        y = Netlist('Oscillator Parameter Sweep')
        h = circuits.colpitts_nvec(y, n=float(n))
        hb = MTHB('HB1')
        hb.options['maxiter'] = 100
        (res, text) = quiet_run(hb.run_oscillator, y, h['f0'], h['K'], h['V0'](float(n)), h['node'])
        conv, freqs, Vf, _, _ = res
        assert conv
        fosc.append(float(re.findall(r'Frequency of oscillation = (\S+) Hz', text)[-1]))
        vosc.append(float(re.findall(r'Oscillation amplitude = (\S+) V', text)[-1]))
        vtank.append(abs(Vf[hb.get_node_idx('nc'), 1]))
        Vfs.append(Vf)
        print(n, fosc[-1], vosc[-1], vtank[-1])
    save('colpitts_nvec', nvec=nvec, fosc=np.array(fosc), Vosc=np.array(vosc), vtank=np.array(vtank), Vf=np.array(Vfs))


def diode_charge_fixture(Netlist, MTHB, save):
    """The reference's own diode charge model (Diode.calc_oppoint, Diode.py:207-235: what its AC / transient analyses use)
    at 96 junction voltages on both sides of Fc Vj: pins the model the hb.diode_charge extension evaluates in HB.
        python tests/golden/make_golden.py --diode-charge"""
    y = Netlist('d')
    d = y.add_diode('D1', 'n1', 'gnd')
    opt = {'Is': 2e-14, 'N': 1.1, 'Cj0': 2e-9, 'Vj': 0.8, 'M': 0.4, 'Fc': 0.5, 'Tt': 1e-7, 'Cp': 1e-10, 'Area': 2.0}
    d.options.update(opt)
    d.init()
    vd = np.linspace(-3.0, 0.75, 96)
    out = np.zeros((96, 5))
    for i, v in enumerate(vd):
        d.It, d.gt = 0., 0.          # Rs = 0: the junction voltage is the terminal voltage (calc_dc, Diode.py:251-255)
        d.calc_oppoint(np.array([v]), usevlimit=False)
        out[i] = [d.oppoint['Vd'], d.oppoint['gd'], d.oppoint['Id'], d.oppoint['Cd'], d.oppoint['Qd']]
    names = sorted(d.options.keys())
    save('diode_charge_model', vd=vd, out=out, opt_names=np.array(names), opt_vals=np.array([float(d.options[k]) for k in names]))


def main():
    import circuits
    Netlist, MTHB = import_reference()
    out = {}
    if '--diode-charge' in sys.argv:
        return diode_charge_fixture(Netlist, MTHB, save_npz)
    if '--osc-root' in sys.argv:
        return osc_root_fixture(Netlist, MTHB, save_npz)
    if '--colpitts-nvec' in sys.argv:
        return colpitts_nvec_fixture(Netlist, MTHB, save_npz)

    save = save_npz
    if '--extra' in sys.argv:
        return extra_fixtures(Netlist, MTHB, save)
    if '--ladder' in sys.argv:   # config 5 reductions: minutes of reference time
        for tag, stages, nh in (('ladder32_4x4', 32, [4, 4]), ('ladder64_2x2', 64, [2, 2])):
            y = Netlist('ladder')
            h = circuits.ladder(y, stages=stages, cje=0.0)
            hb = MTHB('HB1', h['freq'], nh)
            (conv, freqs, Vf, _, _), text = quiet_run(hb.run, y)
            al, its, errs = parse_levels(text)
            save(tag, converged=conv, freqs=freqs if conv else np.zeros(1), Vf=Vf if conv else np.zeros(1),
                 V=hb.V if conv else np.zeros(1), alphas=al, iters=its, errs=errs)
        return

    # ---- config 1: HB_Diode, cold start, plus tight-tolerance variant and two more drive levels
    for tag, ac, opts in (('hb_diode', 5.0, {}), ('hb_diode_tight', 5.0, {'reltol': 1e-9, 'abstol': 1e-12}),
                          ('hb_diode_ac0p5', 0.5, {}), ('hb_diode_ac2', 2.0, {})):
        y = Netlist('Diode Testbench')
        h = circuits.hb_diode(y, ac=ac)
        hb = MTHB('HB1', h['freq'], h['K'])
        hb.options.update(opts)
        with Recorder(hb_W(hb, y)) as rec:
            rec.on = (tag == 'hb_diode')
            (conv, freqs, Vf, _, _), text = quiet_run(hb.run, y)
        al, its, errs = parse_levels(text)
        extra = {}
        if tag == 'hb_diode':
            extra = dict(J0=rec.J[0], F0=rec.F[0], dV0=rec.dV[0], J5=rec.J[5], F5=rec.F[5], dV5=rec.dV[5],
                         DFT=hb.DFT, IDFT=hb.IDFT)
        save(tag, converged=conv, freqs=freqs, Vf=Vf, V=hb.V, alphas=al, iters=its, errs=errs, **extra)

    # ---- config 3: two tones [3,3] and [8,8]
    for tag, nh in (('two_tones_3x3', [3, 3]), ('two_tones_8x8', [8, 8])):
        y = Netlist('Diode Testbench')
        h = circuits.hb_diode_two_tones(y)
        hb = MTHB('HB1', h['freq'], nh)
        with Recorder(hb_W(hb, y)) as rec:
            rec.on = (tag == 'two_tones_3x3')
            (conv, freqs, Vf, _, _), text = quiet_run(hb.run, y)
        al, its, errs = parse_levels(text)
        extra = {}
        if tag == 'two_tones_3x3':
            extra = dict(J0=rec.J[0], F0=rec.F[0], dV0=rec.dV[0], J9=rec.J[9], F9=rec.F[9], dV9=rec.dV[9])
        save(tag, converged=conv, freqs=freqs, Vf=Vf, V=hb.V, alphas=al, iters=its, errs=errs, **extra)

    # ---- HB_BJT (no junction capacitance)
    y = Netlist('bjt')
    h = circuits.hb_bjt(y)
    hb = MTHB('HB1', h['freq'], h['K'])
    (conv, freqs, Vf, _, _), text = quiet_run(hb.run, y)
    al, its, errs = parse_levels(text)
    save('hb_bjt', converged=conv, freqs=freqs, Vf=Vf, V=hb.V, alphas=al, iters=its, errs=errs)

    # ---- config 4 (reference-pinned): DiffAmp cold starts on a small vin x vbias grid + the script's warm sweep
    vins = np.array([100e-6, 5e-3, 20e-3, 48.1e-3])
    vbs = np.array([1.0, 1.2, 1.4])
    Vs, Vfs, itl, convs = [], [], [], []
    for vb in vbs:
        for vin in vins:
            y = Netlist('Differential Amplifier')
            h = circuits.diffamp(y, vin=vin / 2, vbias=vb)
            hb = MTHB('HB1', h['freq'], h['K'])
            with Recorder(hb_W(hb, y)) as rec:
                rec.on = (vb == 1.2 and vin == 20e-3)
                (conv, freqs, Vf, _, _), text = quiet_run(hb.run, y)
            if rec.on:
                recd = dict(J0=rec.J[0], F0=rec.F[0], dV0=rec.dV[0], J1=rec.J[1], F1=rec.F[1], dV1=rec.dV[1],
                            J7=rec.J[7], F7=rec.F[7], dV7=rec.dV[7])
            al, its, errs = parse_levels(text)
            convs.append(conv)
            Vs.append(hb.V if conv else np.full((hb.W, 1), np.nan))
            Vfs.append(Vf if conv else np.full((hb.N, hb.K + 1), np.nan, dtype=complex))
            it16 = np.zeros(64, dtype=int)
            it16[:len(its)] = its
            itl.append(it16)
    save('diffamp_cold_grid', vins=vins, vbs=vbs, converged=np.array(convs), V=np.array(Vs), Vf=np.array(Vfs),
         iters=np.array(itl), freqs=freqs, **recd)

    # warm-start chain exactly as tests/HB_DiffAmp.py:110-121
    y = Netlist('Differential Amplifier')
    h = circuits.diffamp(y)
    hb = MTHB('HB1', h['freq'], h['K'])
    V0 = None
    Vs, Vfs, itl = [], [], []
    sweep = np.arange(100e-6, 50.1e-3, 2e-3)
    for vin in sweep:
        h['i2'].ac = vin / 2
        h['i4'].ac = vin / 2
        (conv, freqs, Vf, _, _), text = quiet_run(hb.run, y, V0)
        V0 = hb.V
        al, its, errs = parse_levels(text)
        it16 = np.zeros(16, dtype=int)
        it16[:len(its)] = its
        itl.append(it16)
        Vs.append(hb.V.copy())
        Vfs.append(Vf.copy())
    save('diffamp_warm_sweep', vins=sweep, V=np.array(Vs), Vf=np.array(Vfs), iters=np.array(itl))

    # ---- device-model pins: BJT.get_hb_params / Diode.get_mthb_params on seeded inputs
    rng = np.random.default_rng(1234)
    y = Netlist('x')
    h = circuits.diffamp(y)
    q = h['q'][0]
    q.options = dict(q.options)
    q.options.update(Tf=4e-10, Tr=2.1e-8, Xtf=2.0, Vtf=3.0, Itf=0.05, Cjs=1e-12, Mjs=0.3)
    q.init()
    n = 512
    vin_ = np.column_stack([rng.uniform(-1.0, 1.0, n), rng.uniform(-2, 5, n), rng.uniform(-1.0, 0.6, n)])
    vold = vin_ + rng.normal(0, 0.2, vin_.shape)
    outs = np.zeros((n, 14))
    with np.errstate(all='ignore'):
        for i in range(n):
            outs[i] = [float(np.asarray(v)) for v in
                       q.get_hb_params(vin_[i, 0], vin_[i, 1], vin_[i, 2], 0, 0, vold[i, 0], vold[i, 1], vold[i, 2])]
    names = sorted(q.options.keys())
    save('bjt_model', vin=vin_, vold=vold, out=outs, opt_names=np.array(names),
         opt_vals=np.array([float(q.options[k]) for k in names]))

    y = Netlist('x')
    d = y.add_diode('D', 'a', 'gnd')
    d.options.update(Is=2e-14, N=1.1, Isr=1e-12, Nr=2.0, Ikf=0.02, Bv=6.0, Ibv=1e-3)
    vd = rng.uniform(-8, 1.2, n)
    vdold = vd + rng.normal(0, 0.3, n)
    outs = np.zeros((n, 2))
    with np.errstate(all='ignore'):
        for i in range(n):
            outs[i] = [float(np.asarray(v)) for v in d.get_mthb_params(vd[i], vdold[i])]
    names = sorted(d.options.keys())
    save('diode_model', vd=vd, vdold=vdold, out=outs, opt_names=np.array(names),
         opt_vals=np.array([float(d.options[k]) for k in names]))

    # ---- config 2: Peltz oscillator (README example), ~15 s
    y = Netlist('Oscillator')
    h = circuits.peltz(y)
    hb = MTHB('HB1')
    hb.options['maxiter'] = 100
    (res, text) = quiet_run(hb.run_oscillator, y, h['f0'], h['K'], h['V0'], h['node'])
    conv, freqs, Vf, _, _ = res
    fosc = float(re.findall(r'Frequency of oscillation = (\S+) Hz', text)[-1])
    vosc = float(re.findall(r'Oscillation amplitude = (\S+) V', text)[-1])
    evals = np.array([[float(a), float(b), float(c)] for a, b, c in
                      re.findall(r'\n\d+\t(\S+)\t(\S+)\t(\S+)\n', text)])
    save('peltz_osc', converged=conv, freqs=freqs, Vf=Vf, V=hb.V, fosc=fosc, Vosc=vosc, evals=evals)

    # one plain (non-oscillator) Peltz-with-probe HB solve at a fixed (f, V): pins the IHF / inductor / capacitor stamps
    y = Netlist('Oscillator')
    h = circuits.peltz(y)
    y.add_iac('P.I', 'P.n1', 'gnd', ac=0.7, freq=71e3)
    y.add_gyrator('P.G', 'P.n1', 'P.n2', 'gnd', 'gnd', 1)
    y.add_idealharmonicfilter('P.IHF', 'P.n2', 'nb', 71e3)
    hb = MTHB('HB1', 71e3, 10)
    with Recorder(hb_W(hb, y)) as rec:
        rec.on = True
        (conv, freqs, Vf, _, _), text = quiet_run(hb.run, y)
    al, its, errs = parse_levels(text)
    save('peltz_probe_fixed', converged=conv, freqs=freqs, Vf=Vf, V=hb.V, alphas=al, iters=its, errs=errs,
         J0=rec.J[0], F0=rec.F[0], dV0=rec.dV[0])

    # ---- Van der Pol oscillator (cubic nonlinearity)
    y = Netlist('vdp')
    h = circuits.vanderpol(y)
    hb = MTHB('HB1')
    (res, text) = quiet_run(hb.run_oscillator, y, h['f0'], h['K'], h['V0'], h['node'])
    conv, freqs, Vf, _, _ = res
    fosc = float(re.findall(r'Frequency of oscillation = (\S+) Hz', text)[-1])
    vosc = float(re.findall(r'Oscillation amplitude = (\S+) V', text)[-1])
    save('vanderpol_osc', converged=conv, freqs=freqs, Vf=Vf, V=hb.V, fosc=fosc, Vosc=vosc)

    # ---- config 5 reduction: 8-stage ladder, [2,2], Cje = 0 (oracle-checkable)
    y = Netlist('ladder')
    h = circuits.ladder(y, stages=8, cje=0.0)
    hb = MTHB('HB1', h['freq'], [2, 2])
    (conv, freqs, Vf, _, _), text = quiet_run(hb.run, y)
    al, its, errs = parse_levels(text)
    save('ladder8_2x2', converged=conv, freqs=freqs, Vf=Vf, V=hb.V, alphas=al, iters=its, errs=errs)


if __name__ == '__main__':
    main()
