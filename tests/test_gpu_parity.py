"""GPU parity tests: the CUDA path (through the C ABI) against the golden fixtures written by the unmodified
reference (tests/golden/) and against the numpy oracle on the same inputs.  Run on the B200 box:
    python -m pytest tests -m gpu -q
Tolerances: 1e-9 relative on converged phasors (BASELINE.json north_star); per-level Newton iteration
counts must be identical."""
import ctypes as C
import os

import numpy as np
import pytest

import circuits

pytestmark = pytest.mark.gpu


def relmax(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.fixture(scope='module')
def yb():
    import torch
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    import yalrf_b200
    return yalrf_b200


def mthb(*a, **k):
    from yalrf_b200.Analyses import MultiToneHarmonicBalance
    return MultiToneHarmonicBalance(*a, **k)


# ----------------------------------------------------------------------------- DC operating point
@pytest.mark.parametrize('name', ['hb_diode', 'diffamp', 'peltz', 'hb_bjt', 'diffamp_activeload', 'vanderpol'])
def test_dc_vs_oracle(yb, name):
    from oracle import hb_oracle as ho
    from yalrf_b200.Analyses import DC
    build = getattr(circuits, name)
    y = yb.Netlist('x')
    build(y)
    x = DC('dc').run(y)
    yo = ho.FlatNetlist()
    build(yo)
    xo = ho.DCOracle(yo).run()
    n = y.get_num_nodes() - 1
    assert np.max(np.abs(x[:n, 0] - xo[:n, 0])) <= 1e-10 * max(1.0, np.max(np.abs(xo[:n, 0])))


def test_dc_engines_bit_identical(yb, monkeypatch):
    """The one-warp DC engine (k_dc_warp: junction-per-lane evaluation, entry-parallel ordered stamping, register
    Gaussian elimination) against the general engine (k_dc) on a bias sweep whose points need plain Newton, source
    stepping and many limiter-bound iterations: bit-identical operating points, identical Newton iteration counts,
    and the oracle's counts (DC.py:125-279)."""
    from oracle import hb_oracle as ho
    from yalrf_b200 import _engine as E
    from yalrf_b200._flatten import ParamBlock
    from yalrf_b200.Analyses.DC import dc_solve_batch
    for build, dev_keys, lo, hi in [(circuits.diffamp, ('i3', 'i5'), 1.0, 1.4), (circuits.diffamp_activeload, ('i3', 'i5'), 1.3, 1.7)]:
        y = yb.Netlist('x')
        h = build(y)
        n = 48
        vb = np.linspace(lo, hi, n)
        plan, topo = E.get_plan(y, h['freq'], h['K'])
        pb = ParamBlock(topo, h['freq'], n)
        for k in dev_keys:
            pb.set(h[k], 'dc', vb)
        monkeypatch.delenv('YHB_DC_ENGINE', raising=False)
        v_warp, i_warp = dc_solve_batch(plan, pb)
        monkeypatch.setenv('YHB_DC_ENGINE', 'block')
        v_blk, i_blk = dc_solve_batch(plan, pb)
        monkeypatch.delenv('YHB_DC_ENGINE', raising=False)
        assert np.array_equal(i_warp, i_blk)
        assert np.array_equal(v_warp.view(np.int64), v_blk.view(np.int64))
        assert i_warp[:, 0].all()
        for j in (0, 17, 47):
            yo = ho.FlatNetlist()
            build(yo, vbias=vb[j])
            o = ho.DCOracle(yo)
            o.total_newton = 0
            xo = o.run()
            assert o.total_newton == i_warp[j, 1]
            assert np.max(np.abs(v_warp[j] - xo[:plan.N, 0])) <= 1e-10 * max(1.0, np.max(np.abs(xo[:plan.N, 0])))


# ----------------------------------------------------------------------------- config 1
@pytest.mark.parametrize('tag,ac,opts', [('hb_diode', 5.0, {}), ('hb_diode_tight', 5.0, {'reltol': 1e-9, 'abstol': 1e-12}),
                                         ('hb_diode_ac0p5', 0.5, {}), ('hb_diode_ac2', 2.0, {})])
def test_hb_diode(yb, golden, tag, ac, opts):
    g = golden(tag)
    y = yb.YalRF('Diode Testbench')
    h = circuits.hb_diode(y, ac=ac)
    hb = mthb('HB1', h['freq'], h['K'])
    hb.options.update(opts)
    conv, freqs, Vf, _, _ = hb.run(y)
    assert conv
    assert hb.last.levels() == g['iters'].tolist()
    np.testing.assert_allclose(hb.last.level_alpha[0][:len(g['alphas'])], g['alphas'], rtol=1e-15)
    assert relmax(hb.V, g['V']) < 1e-9
    assert relmax(Vf, g['Vf']) < 1e-9
    np.testing.assert_array_equal(freqs, g['freqs'])
    assert hb.V.shape == (hb.W, 1) and Vf.shape == (hb.N, hb.K + 1)


# ----------------------------------------------------------------------------- config 3
@pytest.mark.parametrize('tag,nh', [('two_tones_3x3', [3, 3]), ('two_tones_8x8', [8, 8])])
def test_two_tones(yb, golden, tag, nh):
    g = golden(tag)
    y = yb.YalRF('Diode Testbench')
    h = circuits.hb_diode_two_tones(y)
    hb = mthb('HB1', h['freq'], nh)
    conv, freqs, Vf, _, _ = hb.run(y)
    assert conv
    assert hb.last.levels() == g['iters'].tolist()
    np.testing.assert_allclose(freqs, g['freqs'], rtol=1e-15)
    assert relmax(hb.V, g['V']) < 1e-9
    assert relmax(Vf, g['Vf']) < 1e-9


def test_hb_bjt(yb, golden):
    g = golden('hb_bjt')
    y = yb.YalRF('bjt')
    h = circuits.hb_bjt(y)
    hb = mthb('HB1', h['freq'], h['K'])
    conv, freqs, Vf, _, _ = hb.run(y)
    assert conv
    assert hb.last.levels() == g['iters'].tolist()
    assert relmax(hb.V, g['V']) < 1e-9


# ----------------------------------------------------------------------------- config 4 (reference-pinned)
def test_diffamp_cold_grid_batched(yb, golden):
    """The 12 cold starts of the fixture solved as ONE batch (run_sweep) == 12 reference hb.run(y) calls."""
    g = golden('diffamp_cold_grid')
    y = yb.YalRF('Differential Amplifier')
    h = circuits.diffamp(y)
    vb, vin = np.meshgrid(g['vbs'], g['vins'], indexing='ij')
    vb, vin = vb.ravel(), vin.ravel()
    hb = mthb('HB1', h['freq'], h['K'])
    res = hb.run_sweep(y, [(h['i2'], 'ac', vin / 2), (h['i4'], 'ac', vin / 2), (h['i3'], 'dc', vb), (h['i5'], 'dc', vb)])
    for k in range(len(vb)):
        assert bool(res.converged[k]) == bool(g['converged'][k])
        its = g['iters'][k]
        assert res.levels(k) == its[its > 0].tolist(), (k, vb[k], vin[k])
        assert relmax(res.V[k], g['V'][k][:, 0]) < 1e-9, (k, vb[k], vin[k])
        for n in range(res.Vf.shape[1]):
            assert relmax(res.Vf[k, n], g['Vf'][k][n]) < 1e-9


def test_diffamp_warm_sweep(yb, golden):
    """tests/HB_DiffAmp.py:110-121 verbatim: sequential warm-start chain through hb.run(y, V0)."""
    g = golden('diffamp_warm_sweep')
    y = yb.YalRF('Differential Amplifier')
    h = circuits.diffamp(y)
    hb = mthb('HB1', h['freq'], h['K'])
    V0 = None
    for k, vin in enumerate(g['vins']):
        h['i2'].ac = vin / 2
        h['i4'].ac = vin / 2
        conv, freqs, Vf, _, _ = hb.run(y, V0)
        V0 = hb.V
        assert conv
        its = g['iters'][k]
        assert hb.last.levels() == its[its > 0].tolist()
        assert relmax(hb.V, g['V'][k]) < 1e-9
        assert relmax(Vf, g['Vf'][k]) < 1e-9


def test_peltz_probe_fixed(yb, golden):
    g = golden('peltz_probe_fixed')
    y = yb.Netlist('Oscillator')
    circuits.peltz(y)
    y.add_iac('P.I', 'P.n1', 'gnd', ac=0.7, freq=71e3)
    y.add_gyrator('P.G', 'P.n1', 'P.n2', 'gnd', 'gnd', 1)
    y.add_idealharmonicfilter('P.IHF', 'P.n2', 'nb', 71e3)
    hb = mthb('HB1', 71e3, 10)
    conv, freqs, Vf, _, _ = hb.run(y)
    assert conv
    assert hb.last.levels() == g['iters'].tolist()
    V, G = hb.V.reshape(hb.N, hb.S), g['V'].reshape(hb.N, hb.S)
    probe = y.get_node_idx('P.n1') - 1
    for n in range(hb.N):
        if n != probe:
            assert np.max(np.abs(V[n] - G[n])) < 1e-9 * np.max(np.abs(G)), n
    # V('P.n1') is the probe CURRENT (gyrator), i.e. 1e9 S * (V(P.n2) - V(nb)): one ulp of the 0.7 V phasors is
    # 1.1e-16 V = 1.1e-7 A.  The reference's own value carries that rounding noise, so it is
    # pinned to a few ulps of the filter voltage, not to 1e-9 relative.  Measured on B200: 1.9e-7.
    assert np.max(np.abs(V[probe] - G[probe])) < 1e9 * 8 * np.spacing(0.7)


def test_ladder(yb, golden):
    g = golden('ladder8_2x2')
    y = yb.Netlist('ladder')
    h = circuits.ladder(y, stages=8, cje=0.0)
    hb = mthb('HB1', h['freq'], [2, 2])
    conv, freqs, Vf, _, _ = hb.run(y)
    assert conv == bool(g['converged'])
    assert hb.last.levels() == g['iters'].tolist()
    assert relmax(hb.V, g['V']) < 1e-9


def test_pnp_extension_vs_oracle(yb):
    """ActiveLoad netlist with the PNP polarity extension (parity unpinned by reference HB): oracle with the same
    extension, vin = 100 uV and 1 mV."""
    from oracle import hb_oracle as ho
    for vin in (100e-6, 1e-3):
        y = yb.YalRF('x')
        h = circuits.diffamp_activeload(y, vin=vin / 2)
        hb = mthb('HB1', h['freq'], h['K'])
        hb.pnp_extension = True
        conv, freqs, Vf, _, _ = hb.run(y)
        yo = ho.FlatNetlist()
        ho_h = circuits.diffamp_activeload(yo, vin=vin / 2)
        o = ho.HBOracle('HB1', ho_h['freq'], ho_h['K'], pnp_extension=True)
        conv_o, _, Vf_o, _, _ = o.run(yo)
        assert conv == conv_o
        assert hb.last.levels() == [l[1] for l in o.level_log]
        if conv:
            assert relmax(hb.V, o.V) < 1e-9


def test_nonconvergence_is_data(yb):
    """Unmodified-reference behaviour on the ActiveLoad netlist (PNP treated as NPN): no exception, converged False."""
    y = yb.YalRF('x')
    h = circuits.diffamp_activeload(y)
    hb = mthb('HB1', h['freq'], 2)
    hb.options['maxiter'] = 20
    out = hb.run(y)
    assert out == (False, None, None, None, None)


# ----------------------------------------------------------------------------- engines: fused / staged / unpeeled
@pytest.mark.parametrize('engine,peel', [('auto', True), ('staged', True), ('staged', False), ('auto', False)])
def test_engines_agree_with_golden(yb, golden, engine, peel):
    """The fused on-chip engine, the staged engine, and the unpeeled (full W x W) factorisation all reproduce the
    reference: DiffAmp cold grid (subset), HB_Diode, two tones [3,3], Peltz with probe."""
    g = golden('diffamp_cold_grid')
    y = yb.YalRF('Differential Amplifier')
    h = circuits.diffamp(y)
    vb, vin = np.meshgrid(g['vbs'], g['vins'], indexing='ij')
    vb, vin = vb.ravel(), vin.ravel()
    hb = mthb('HB1', h['freq'], h['K'])
    hb.engine, hb.peel = engine, peel
    res = hb.run_sweep(y, [(h['i2'], 'ac', vin / 2), (h['i4'], 'ac', vin / 2), (h['i3'], 'dc', vb), (h['i5'], 'dc', vb)])
    Nc, Wc, fused = hb._plan.core()
    assert fused == (engine == 'auto' and peel)      # the unpeeled 210 x 210 system does not fit one SM
    assert Wc == (84 if peel else 210)
    for k in range(len(vb)):
        its = g['iters'][k]
        assert res.levels(k) == its[its > 0].tolist(), (k, engine, peel)
        assert relmax(res.V[k], g['V'][k][:, 0]) < 1e-9, (k, engine, peel)

    g = golden('hb_diode')
    y = yb.YalRF('d')
    h = circuits.hb_diode(y)
    hb = mthb('HB1', h['freq'], h['K'])
    hb.engine, hb.peel = engine, peel
    conv, freqs, Vf, _, _ = hb.run(y)
    assert hb._plan.core()[1] == (21 if peel else 63)
    assert conv and hb.last.levels() == g['iters'].tolist()
    assert relmax(hb.V, g['V']) < 1e-9

    g = golden('two_tones_3x3')
    y = yb.YalRF('d')
    h = circuits.hb_diode_two_tones(y)
    hb = mthb('HB1', h['freq'], [3, 3])
    hb.engine, hb.peel = engine, peel
    conv, freqs, Vf, _, _ = hb.run(y)
    assert hb._plan.core()[1] == (49 if peel else 245)
    assert conv and hb.last.levels() == g['iters'].tolist()
    assert relmax(hb.V, g['V']) < 1e-9


def test_warm_start_failure_falls_back_to_ladder(yb):
    """run(y, V0) with a useless V0: the direct attempt fails, the engine restarts from the DC point (:333-345) and
    lands on the cold-start answer; the oracle does the same."""
    from oracle import hb_oracle as ho
    y = yb.YalRF('d')
    h = circuits.hb_diode(y)
    hb = mthb('HB1', h['freq'], h['K'])
    hb.options['maxiter'] = 8
    rng = np.random.default_rng(3)
    V0 = rng.normal(scale=3.0, size=(63, 1))
    conv, freqs, Vf, _, _ = hb.run(y, V0)
    yo = ho.FlatNetlist()
    circuits.hb_diode(yo)
    o = ho.HBOracle('HB1', h['freq'], h['K'])
    o.options['maxiter'] = 8
    conv_o = o.run(yo, V0.copy())[0]
    assert conv == conv_o
    assert hb.last.levels() == [l[1] for l in o.level_log]
    if conv:
        assert relmax(hb.V, o.V) < 1e-9


# ----------------------------------------------------------------------------- config 2: oscillators
def test_peltz_oscillator(yb, golden, capsys):
    """README / tests/HBOsc_Peltz.py: run_oscillator on the Peltz oscillator, K = 10.  fosc, Vosc and the phasors of
    the unmodified reference (tests/golden/peltz_osc.npz); print_v text format (:112-118)."""
    g = golden('peltz_osc')
    y = yb.Netlist('Oscillator')
    h = circuits.peltz(y)
    hb = mthb('HB1')
    hb.options['maxiter'] = 100
    conv, freqs, Vf, _, _ = hb.run_oscillator(y, h['f0'], h['K'], h['V0'], h['node'])
    assert conv
    # Resolution of the reference's own answer: its objective |Yosc| = |V(nb) - V(nosc2)| * 1e9 S / |V(nb)| is
    # quantised -- one ulp of the 0.75 V phasors is 1.1e-16 V = 1.5e-7 S, and dYosc/df = 4 pi C = 1.26e-7 S/Hz, so
    # Re(Yosc) is exactly 0 on a plateau ~1 Hz (1.4e-5 relative) wide whose position depends on the last bit of the
    # converged V(nb), i.e. on the summation order of the residual (BLAS dgemv in the reference).  Measured on
    # B200: |df|/f = 4.1e-6, |dV|/V = 1.2e-5.  Everything that is well conditioned is pinned at 1e-9 elsewhere
    # (test_peltz_probe_fixed: same circuit, probe frequency and amplitude fixed).
    assert abs(hb.fosc - float(g['fosc'])) / float(g['fosc']) < 2e-5
    assert abs(hb.Vosc - float(g['Vosc'])) / float(g['Vosc']) < 5e-5
    assert hb.freq == hb.fosc                                               # result also in hb.freq (:226)
    nb = y.get_node_idx('nb') - 1
    assert relmax(Vf[nb], g['Vf'][nb]) < 5e-5
    assert abs(Vf[nb][0] - g['Vf'][nb][0]) / abs(g['Vf'][nb][0]) < 1e-9       # DC level does not depend on (f, V)
    # the GPU answer evaluated by the reference's own objective definition is at the quantisation floor
    n1, n2 = hb.get_node_idx('nb'), hb.get_node_idx('HB1.nosc2')
    assert abs((Vf[n1, 1] - Vf[n2, 1]) * 1e9 / Vf[n1, 1]) < 1e-6
    hb.print_v('nb')
    out = capsys.readouterr().out
    assert 'Voltage at node: nb' in out and 'Freq [Hz]\tVmag [V]\tPhase [°]' in out
    assert 'Frequency of oscillation = ' in out and 'Oscillation amplitude = ' in out


def test_vanderpol_oscillator(yb, golden):
    """tests/HBOsc_VanDerPol.py: cubic nonlinearity, inductor / capacitor / negative resistor, K = 20."""
    g = golden('vanderpol_osc')
    y = yb.Netlist('vdp')
    h = circuits.vanderpol(y)
    hb = mthb('HB1')
    conv, freqs, Vf, _, _ = hb.run_oscillator(y, h['f0'], h['K'], h['V0'], h['node'])
    assert conv
    # the fixture is the reference's Nelder-Mead answer (xtol 1e-12 on a simplex that stalls at the 1e-8 level: measured
    # distance from the Newton root 2.8e-8 in Vosc); the device Newton iteration itself ends at |Yosc| < 1e-12
    assert abs(hb.fosc - float(g['fosc'])) / float(g['fosc']) < 1e-7
    assert abs(hb.Vosc - float(g['Vosc'])) / float(g['Vosc']) < 1e-7
    assert hb.osc_yosc < 1e-12
    # the reference's own outer loop replayed on the GPU solves lands on the fixture to 1e-8
    hb2 = mthb('HB1')
    hb2.osc_method = 'fmin'
    conv2, _, _, _, _ = hb2.run_oscillator(y, h['f0'], h['K'], h['V0'], h['node'])
    assert conv2
    assert abs(hb2.fosc - float(g['fosc'])) / float(g['fosc']) < 1e-8
    assert abs(hb2.Vosc - float(g['Vosc'])) / float(g['Vosc']) < 1e-8


def test_oscillator_newton_batch(yb, golden):
    """run_oscillator_batch: the scipy-free oscillator solve (2x2 Newton on (f, V), probe current from KCL).  Eight
    start points land on one root to 1e-9; it lies within the resolution of the reference's Nelder-Mead answer
    (see test_peltz_oscillator); the oracle, solving the same circuit with the probe fixed at that (f, V), finds
    the probe carrying no current.  A capacitor sweep follows f ~ 1 / sqrt(L C)."""
    from oracle import hb_oracle as ho
    from yalrf_b200 import _osc
    from yalrf_b200._flatten import ParamBlock, Topology
    g = golden('peltz_osc')
    y = yb.Netlist('Oscillator')
    h = circuits.peltz(y)
    hb = mthb('HB1')
    f0 = np.linspace(70e3, 90e3, 8)
    V0 = np.linspace(0.05, 0.8, 8)
    r = hb.run_oscillator_batch(y, f0, h['K'], V0, h['node'])
    assert r.converged.all(), (r.converged, r.yosc)
    assert np.max(np.abs(r.fosc - r.fosc[0])) / r.fosc[0] < 1e-9
    assert np.max(np.abs(r.Vosc - r.Vosc[0])) / r.Vosc[0] < 1e-9
    assert np.max(r.yosc) < 1e-12
    assert abs(r.fosc[0] - float(g['fosc'])) / float(g['fosc']) < 2e-5
    assert abs(r.Vosc[0] - float(g['Vosc'])) / float(g['Vosc']) < 5e-5
    # oracle check at the GPU answer: fixed probe, tight inner tolerance, KCL objective from the oracle's own V and Inl
    yo = ho.FlatNetlist()
    circuits.peltz(yo)
    yo.add_iac('HB1.I.Oscprobe', 'HB1.nosc1', 'gnd', ac=float(r.Vosc[0]), freq=float(r.fosc[0]))
    yo.add_gyrator('HB1.G.Oscprobe', 'HB1.nosc1', 'HB1.nosc2', 'gnd', 'gnd', 1)
    yo.add_idealharmonicfilter('HB1IHF.Oscprobe', 'HB1.nosc2', 'nb', float(r.fosc[0]))
    o = ho.HBOracle('HB1', float(r.fosc[0]), h['K'])
    assert o.run(yo)[0]
    for _ in range(3):                      # warm restarts: >= 3 more Newton iterations each (tolerances cannot be
        assert o.run(yo, o.V)[0]            # tightened: the 1e9 S probe branch leaves 1e-7 A of rounding noise in F)
    topo = Topology(r.netlist, 1, (h['K'],))
    pb = ParamBlock(topo, float(r.fosc[0]), 1).finalize()
    pz = [d for _, _, d in topo.lin if d.name == 'HB1IHF.Oscprobe'][0]
    Y = _osc.kcl_yosc(topo, pb, pz, r.netlist.get_node_idx('nb'), o.V.reshape(1, -1), o.last_Inl.reshape(1, -1), 21)
    assert abs(Y[0]) < 1e-10, abs(Y[0])   # vs |dYosc/df| = 1.3e-7 S/Hz: the root is located to < 1e-3 Hz
    # parameter sweep: tank capacitance 5 ... 20 nF from one start point
    c = np.linspace(5e-9, 20e-9, 6)
    r2 = hb.run_oscillator_batch(y, 80e3, h['K'], 0.1, h['node'], sweeps=[(h['c1'], 'C', c)], maxiter=40)
    assert r2.converged.all(), (r2.converged, r2.yosc)
    f_lc = 1 / (2 * np.pi * np.sqrt(0.5e-3 * c))
    assert np.all(np.abs(r2.fosc / f_lc - 1) < 0.01)
    assert abs(r2.fosc[np.argmin(np.abs(c - 10e-9))] - r.fosc[0]) / r.fosc[0] < 0.35   # 10 nF is not on this grid


def test_vanderpol_newton(yb, golden):
    g = golden('vanderpol_osc')
    y = yb.Netlist('vdp')
    h = circuits.vanderpol(y)
    hb = mthb('HB1')
    r = hb.run_oscillator_batch(y, [h['f0'], 0.15], h['K'], [h['V0'], 1.0], h['node'])
    assert r.converged.all(), (r.converged, r.yosc)
    assert abs(r.fosc[0] - float(g['fosc'])) / float(g['fosc']) < 1e-7
    assert abs(r.Vosc[0] - float(g['Vosc'])) / float(g['Vosc']) < 1e-7
    assert abs(r.fosc[1] - r.fosc[0]) / r.fosc[0] < 1e-9


def test_oscillator_root_under_reference_objective(yb, golden):
    """the GPU engine's oscillation point fed back through the REFERENCE.  tests/golden/
    peltz_ref_at_gpu_root.npz (make_golden.py --osc-root) holds the unmodified reference's HB solution of the Peltz
    oscillator with its probe fixed at the GPU root and the reference's own objective Yosc (:183-186) there: the
    objective sits below its own resolution (one ulp of the probe voltage x 1e9 S), and the reference's phasors at that
    point are the GPU's."""
    g = golden('peltz_ref_at_gpu_root')
    assert bool(g['converged'])
    assert abs(complex(g['yosc'])) < float(g['ulp_floor'])                   # 8.1e-10 S against 1.5e-7 S
    y = yb.Netlist('Oscillator')
    h = circuits.peltz(y)
    hb = mthb('HB1')
    conv, freqs, Vf, _, _ = hb.run_oscillator(y, h['f0'], h['K'], h['V0'], h['node'])
    assert conv
    # the engine still lands on the point the fixture was written for
    assert abs(hb.fosc - float(g['fosc'])) / float(g['fosc']) < 1e-10
    assert abs(hb.Vosc - float(g['Vosc'])) / float(g['Vosc']) < 1e-10
    # the same solve as the fixture's -- cold start, probe fixed at the root -- on the GPU: identical iteration counts,
    # phasors within 1e-9 (the oscillator run's own last iterate is a warm start: another end of the iteration)
    y2 = yb.Netlist('Oscillator')
    circuits.peltz(y2)
    circuits.add_probe(y2, 'HB1', h['node'], float(g['fosc']), float(g['Vosc']))
    hb2 = mthb('HB1', float(g['fosc']), h['K'])
    assert hb2.run(y2)[0]
    assert hb2.last.levels() == g['iters'].tolist()
    V, G = hb2.V.reshape(hb2.N, hb2.S), g['V'].reshape(hb2.N, hb2.S)
    probe = hb2.get_node_idx('HB1.nosc1')
    for n in range(hb2.N):
        if n != probe:
            assert np.max(np.abs(V[n] - G[n])) < 1e-9 * np.max(np.abs(G)), n
    # the probe-current node carries 1e9 S x rounding noise in the reference (see test_peltz_probe_fixed)
    assert np.max(np.abs(V[probe] - G[probe])) < 1e9 * 8 * np.spacing(0.75)


def test_peltz_oscillator_fmin_replay(yb, golden):
    """hb.osc_method = 'fmin': the reference's own Nelder-Mead outer loop (:204-212), every evaluation a GPU solve."""
    g = golden('peltz_osc')
    y = yb.Netlist('Oscillator')
    h = circuits.peltz(y)
    hb = mthb('HB1')
    hb.osc_method = 'fmin'
    conv, freqs, Vf, _, _ = hb.run_oscillator(y, h['f0'], h['K'], h['V0'], h['node'])
    assert conv and len(hb.osc_evals) <= 300
    assert abs(hb.fosc - float(g['fosc'])) / float(g['fosc']) < 2e-5
    assert abs(hb.Vosc - float(g['Vosc'])) / float(g['Vosc']) < 5e-5


def test_oscillator_sweep_vs_oracle(yb):
    """a run_oscillator parameter sweep (the reference loops run_oscillator over a component value,
    tests/HBOscillators.ipynb:3052-3115) as ONE device call, every point checked against the oracle: with the probe
    fixed at the GPU's (fosc, Vosc) the oracle's own HB solution carries no probe current (KCL objective), and its
    phasors agree with the GPU's."""
    from oracle import hb_oracle as ho
    from yalrf_b200 import _osc
    from yalrf_b200._flatten import ParamBlock, Topology
    y = yb.Netlist('Oscillator')
    h = circuits.peltz(y)
    hb = mthb('HB1')
    c = np.array([6e-9, 10e-9, 16e-9])
    r = hb.run_oscillator_batch(y, 80e3, h['K'], 0.1, h['node'], sweeps=[(h['c1'], 'C', c)], maxiter=40)
    assert r.converged.all(), (r.converged, r.yosc)
    assert np.max(r.yosc) < 1e-12
    assert np.all(r.outer_iterations <= 25)
    f_lc = 1 / (2 * np.pi * np.sqrt(0.5e-3 * c))
    assert np.all(np.abs(r.fosc / f_lc - 1) < 0.01)
    nb = r.netlist.get_node_idx('nb')
    for k in range(len(c)):
        yo = ho.FlatNetlist()
        ho_h = circuits.peltz(yo)
        ho_h['c1'].C = float(c[k])
        yo.add_iac('HB1.I.Oscprobe', 'HB1.nosc1', 'gnd', ac=float(r.Vosc[k]), freq=float(r.fosc[k]))
        yo.add_gyrator('HB1.G.Oscprobe', 'HB1.nosc1', 'HB1.nosc2', 'gnd', 'gnd', 1)
        yo.add_idealharmonicfilter('HB1IHF.Oscprobe', 'HB1.nosc2', 'nb', float(r.fosc[k]))
        o = ho.HBOracle('HB1', float(r.fosc[k]), h['K'])
        assert o.run(yo)[0]
        for _ in range(3):
            assert o.run(yo, o.V)[0]
        topo = Topology(r.netlist, 1, (h['K'],))
        pb = ParamBlock(topo, float(r.fosc[k]), 1)
        pb.set(h['c1'], 'C', [float(c[k])])
        pb.finalize()
        pz = [d for _, _, d in topo.lin if d.name == 'HB1IHF.Oscprobe'][0]
        Y = _osc.kcl_yosc(topo, pb, pz, nb, o.V.reshape(1, -1), o.last_Inl.reshape(1, -1), 21)
        assert abs(Y[0]) < 1e-9, (k, abs(Y[0]))
        assert relmax(r.V[k], o.V[:, 0]) < 1e-6, k      # same circuit, same probe: the HB solutions agree


def test_colpitts_divider_sweep(yb, golden):
    """tests/HBOscillators.ipynb cell 8 (the reference's heaviest workload: a loop of 20 run_oscillator calls over the
    capacitive divider ratio n of a Colpitts oscillator, K = 20, 110 s there) as ONE batched device call.  Three of its
    points against cold run_oscillator runs of the unmodified reference (colpitts_nvec.npz; its Nelder-Mead stops at 300
    evaluations, so its answer is good to ~1e-4 in Vosc), and every point against the oracle: with the probe fixed at the
    GPU's (fosc, Vosc) the oracle's own HB solution carries no probe current, and its phasors are the GPU's."""
    from oracle import hb_oracle as ho
    from yalrf_b200 import _osc
    from yalrf_b200._flatten import ParamBlock, Topology
    g = golden('colpitts_nvec')
    nvec = g['nvec']
    y = yb.Netlist('Oscillator Parameter Sweep')
    h = circuits.colpitts_nvec(y)
    hb = mthb('HB1')
    hb.options['maxiter'] = 100
    sweeps = [(h['c1'], 'C', h['ct'] / (1 - nvec)), (h['c2'], 'C', h['ct'] / nvec)]
    r = hb.run_oscillator_batch(y, np.full(len(nvec), h['f0']), h['K'], h['V0'](nvec), h['node'], sweeps=sweeps, maxiter=60)
    assert r.converged.all(), (r.converged, r.yosc)
    assert np.max(r.yosc) < 1e-12 and np.all(r.outer_iterations <= 10)
    assert np.max(np.abs(r.fosc / g['fosc'] - 1)) < 1e-5
    assert np.max(np.abs(r.Vosc / g['Vosc'] - 1)) < 1e-3
    nc = r.netlist.get_node_idx('nc')
    assert np.max(np.abs(np.abs(r.Vf[:, nc - 1, 1]) / g['vtank'] - 1)) < 1e-3
    S = 2 * h['K'] + 1
    for k in range(len(nvec)):
        yo = ho.FlatNetlist()
        circuits.colpitts_nvec(yo, n=float(nvec[k]))
        circuits.add_probe(yo, 'HB1', 'nc', float(r.fosc[k]), float(r.Vosc[k]))
        o = ho.HBOracle('HB1', float(r.fosc[k]), h['K'])
        o.options['maxiter'] = 100
        assert o.run(yo)[0]
        for _ in range(3):
            assert o.run(yo, o.V)[0]
        topo = Topology(r.netlist, 1, (h['K'],))
        pb = ParamBlock(topo, float(r.fosc[k]), 1)
        pb.set(h['c1'], 'C', [float(h['ct'] / (1 - nvec[k]))])
        pb.set(h['c2'], 'C', [float(h['ct'] / nvec[k])])
        pb.finalize()
        pz = [d for _, _, d in topo.lin if d.name == 'HB1IHF.Oscprobe'][0]
        Y = _osc.kcl_yosc(topo, pb, pz, nc, o.V.reshape(1, -1), o.last_Inl.reshape(1, -1), S)
        assert abs(Y[0]) < 1e-8, (k, abs(Y[0]))
        assert relmax(r.V[k], o.V[:, 0]) < 1e-6, k


# ----------------------------------------------------------------------------- fused engine: block-sparse core LU
@pytest.mark.parametrize('mode', [None, True, False])
def test_block_sparse_core_lu(yb, golden, mode):
    """The fused engine factors the core node block by node block when its block structure is sparse (DiffAmp: the
    collectors and the mirror node couple to the emitter node only, 2 levels), hb.block_sparse = True forces it on
    dense block structures too, False keeps the dense Gauss-Jordan: all three reproduce the reference."""
    g = golden('diffamp_cold_grid')
    y = yb.YalRF('Differential Amplifier')
    h = circuits.diffamp(y)
    vb, vin = np.meshgrid(g['vbs'], g['vins'], indexing='ij')
    vb, vin = vb.ravel(), vin.ravel()
    hb = mthb('HB1', h['freq'], h['K'])
    hb.block_sparse = mode
    res = hb.run_sweep(y, [(h['i2'], 'ac', vin / 2), (h['i4'], 'ac', vin / 2), (h['i3'], 'dc', vb), (h['i5'], 'dc', vb)])
    assert hb._plan.engine() == 'fused'
    assert hb._plan.block_sparse() == (0 if mode is False else 2)
    for k in range(len(vb)):
        its = g['iters'][k]
        assert bool(res.converged[k]) == bool(g['converged'][k])
        assert res.levels(k) == its[its > 0].tolist(), (k, mode)
        assert relmax(res.V[k], g['V'][k][:, 0]) < 1e-9, (k, mode)

    # warm-start chain of tests/HB_DiffAmp.py
    g = golden('diffamp_warm_sweep')
    V0 = None
    for k, v in enumerate(g['vins'][:8]):
        h['i2'].ac = v / 2
        h['i4'].ac = v / 2
        conv, freqs, Vf, _, _ = hb.run(y, V0)
        V0 = hb.V
        its = g['iters'][k]
        assert conv and hb.last.levels() == its[its > 0].tolist()
        assert relmax(hb.V, g['V'][k]) < 1e-9

    # dense block structures (forced) and single-node cores (never): Peltz with the probe fixed, current mirror, hb_bjt
    for tag, builder in [('current_mirror', 'current_mirror'), ('hb_bjt', 'hb_bjt')]:
        g = golden(tag)
        y = yb.YalRF(tag)
        h = getattr(circuits, builder)(y)
        hb = mthb('HB1', h['freq'], h['K'])
        hb.block_sparse = mode
        conv, freqs, Vf, _, _ = hb.run(y)
        if mode is False:
            assert hb._plan.block_sparse() == 0
        assert conv == bool(g['converged'])
        assert hb.last.levels() == g['iters'].tolist(), (tag, mode)
        assert relmax(hb.V, g['V']) < 1e-9, (tag, mode)


@pytest.mark.parametrize('stages,nh,ac', [(4, [1, 1], 1.0), (6, [1, 1], 1.0), (5, [2, 1], 0.5)])
def test_block_sparse_chain_vs_oracle(yb, stages, nh, ac):
    """Short diode / BJT ladders on the fused engine: the core is a chain of node blocks, eliminated from both ends
    towards the middle (three and more levels, blocks of earlier AND later nodes in one block row).  Block-sparse
    and dense core solves against the oracle on the same inputs."""
    from oracle import hb_oracle as ho
    yo = ho.FlatNetlist()
    circuits.ladder(yo, stages=stages, cje=0.0, ac=ac)
    o = ho.HBOracle('HB1', [1.1e6, 0.9e6], nh)
    conv_o = o.run(yo)[0]
    for mode in (None, False):
        y = yb.Netlist('ladder')
        h = circuits.ladder(y, stages=stages, cje=0.0, ac=ac)
        hb = mthb('HB1', h['freq'], nh)
        hb.block_sparse = mode
        conv, freqs, Vf, _, _ = hb.run(y)
        assert hb._plan.engine() == 'fused'
        assert hb._plan.block_sparse() >= 3 if mode is None else hb._plan.block_sparse() == 0
        assert conv == conv_o
        assert hb.last.levels() == [l[1] for l in o.level_log], mode
        assert relmax(hb.V, o.V) < 1e-9, mode


@pytest.mark.parametrize('n,ac', [(4, 0.5), (6, 1.0)])
def test_block_sparse_fill_vs_oracle(yb, n, ac):
    """A ring of diode nodes: eliminating a ring node couples its two neighbours, so the block-row storage holds FILL
    blocks (zeroed before the assembly, created by the Schur updates).  Block-sparse and dense core solves against
    the oracle on the same inputs."""
    from oracle import hb_oracle as ho
    yo = ho.FlatNetlist()
    h = circuits.ring(yo, n=n, ac=ac)
    o = ho.HBOracle('HB1', h['freq'], h['K'])
    conv_o = o.run(yo)[0]
    for mode in (None, False):
        y = yb.Netlist('ring')
        circuits.ring(y, n=n, ac=ac)
        hb = mthb('HB1', h['freq'], h['K'])
        hb.block_sparse = mode
        conv, freqs, Vf, _, _ = hb.run(y)
        assert hb._plan.engine() == 'fused' and hb._plan.core()[0] == n
        assert (hb._plan.block_sparse() >= 2) if mode is None else hb._plan.block_sparse() == 0
        assert conv == conv_o
        assert hb.last.levels() == [l[1] for l in o.level_log], mode
        assert relmax(hb.V, o.V) < 1e-9, mode


# ----------------------------------------------------------------------------- more reference scripts
@pytest.mark.parametrize('tag,builder,nh', [('current_mirror', 'current_mirror', None),
                                            ('fullwave_rectifier', 'fullwave_rectifier', None),
                                            ('am_demod_3x3', 'am_demod', [3, 3])])
def test_more_reference_scripts(yb, golden, tag, builder, nh):
    """tests/HB_CurrentMirror.py (phase = +90 source, diode-connected BJT), HB_FullWaveRectifier.py (floating
    source between two non-ground nodes, four diodes), HB_AMDemod.py at [3,3] (two tones 1 MHz / 10 kHz, capacitor)."""
    g = golden(tag)
    y = yb.YalRF(tag)
    h = getattr(circuits, builder)(y)
    hb = mthb('HB1', h['freq'], nh if nh is not None else h['K'])
    conv, freqs, Vf, _, _ = hb.run(y)
    assert conv == bool(g['converged'])
    assert hb.last.levels() == g['iters'].tolist()
    from test_oracle_golden import _check_extra
    _check_extra(tag, hb.V, g['V'], hb.S)
    if tag != 'fullwave_rectifier':
        assert relmax(Vf, g['Vf']) < 1e-9
    np.testing.assert_allclose(freqs, g['freqs'], rtol=1e-15)


def test_colpitts_probe_fixed(yb, golden):
    """Colpitts CB (tests/HBOsc_ColpittsCB.py netlist) with the probe fixed: BJT transit-time and depletion charges,
    inductor, capacitors, two supply gyrators, K = 8."""
    g = golden('colpitts_probe_fixed')
    y = yb.Netlist('colpitts')
    circuits.colpitts_cb(y)
    circuits.add_probe(y, 'HB1', 'nc', 9.5e6, 0.8)
    hb = mthb('HB1', 9.5e6, 8)
    conv, freqs, Vf, _, _ = hb.run(y)
    assert conv == bool(g['converged'])
    assert hb.last.levels() == g['iters'].tolist()
    from test_oracle_golden import colpitts_tolerance
    tol = colpitts_tolerance()      # 20 x the reference algorithm's measured self-sensitivity to 1 ulp (about 1e-6)
    V, G = hb.V.reshape(hb.N, hb.S), g['V'].reshape(hb.N, hb.S)
    probe = y.get_node_idx('HB1.nosc1') - 1
    for n in range(hb.N):
        if n != probe:
            assert np.max(np.abs(V[n] - G[n])) < tol * np.max(np.abs(G)), n
    assert np.max(np.abs(V[probe] - G[probe])) < 1e9 * 16 * np.spacing(0.8) + tol * 1e-2   # probe current node


# ----------------------------------------------------------------------------- config 5 reductions: block-banded cores
@pytest.mark.parametrize('stages,nh', [(16, [2, 2]), (24, [3, 3])])
def test_ladder_banded_vs_oracle(yb, stages, nh):
    """Diode / diode-connected-BJT ladder at oracle-checkable sizes: the core is block
    tridiagonal over the ladder nodes, the staged engine stores and eliminates it in band form."""
    from oracle import hb_oracle as ho
    y = yb.Netlist('ladder')
    h = circuits.ladder(y, stages=stages, cje=0.0)
    hb = mthb('HB1', h['freq'], nh)
    conv, freqs, Vf, _, _ = hb.run(y)
    Nc, Wc, fused = hb._plan.core()
    assert not fused and hb._plan.band() > 0 and 3 * hb._plan.band() + 1 < Wc
    yo = ho.FlatNetlist()
    circuits.ladder(yo, stages=stages, cje=0.0)
    o = ho.HBOracle('HB1', h['freq'], nh)
    conv_o = o.run(yo)[0]
    assert conv == conv_o
    assert hb.last.levels() == [l[1] for l in o.level_log]
    assert relmax(hb.V, o.V) < 1e-9
    # dense storage of the same core gives the same answer
    hb2 = mthb('HB1', h['freq'], nh)
    hb2.band = False
    conv2, _, _, _, _ = hb2.run(y)
    assert hb2._plan.band() == 0 and conv2 == conv
    assert relmax(hb2.V, hb.V) < 1e-11


@pytest.mark.parametrize('stages,nh', [(16, [2, 2]), (24, [3, 3])])
def test_ladder_block_thomas_vs_oracle(yb, stages, nh):
    """The same reductions through the block-Thomas engine (block-tridiagonal core solved block by block: the
    engine of config 5 at full size, forced here at oracle-checkable sizes)."""
    from oracle import hb_oracle as ho
    y = yb.Netlist('ladder')
    h = circuits.ladder(y, stages=stages, cje=0.0)
    hb = mthb('HB1', h['freq'], nh)
    hb.block_thomas = True
    conv, freqs, Vf, _, _ = hb.run(y)
    assert hb._plan.engine() == 'staged-block-thomas'
    yo = ho.FlatNetlist()
    circuits.ladder(yo, stages=stages, cje=0.0)
    o = ho.HBOracle('HB1', h['freq'], nh)
    conv_o = o.run(yo)[0]
    assert conv == conv_o
    assert hb.last.levels() == [l[1] for l in o.level_log]
    assert relmax(hb.V, o.V) < 1e-9


@pytest.mark.parametrize('tag,stages,nh,bt', [('ladder32_4x4', 32, [4, 4], None), ('ladder64_2x2', 64, [2, 2], None),
                                              ('ladder32_4x4', 32, [4, 4], True), ('ladder64_2x2', 64, [2, 2], True)])
def test_ladder_golden(yb, golden, tag, stages, nh, bt):
    """SURVEY 8d's oracle-checkable reductions of config 5, against the unmodified reference."""
    import os
    from conftest import GOLDEN
    if not os.path.exists(os.path.join(GOLDEN, tag + '.npz')):
        pytest.skip('fixture not generated')
    g = golden(tag)
    y = yb.Netlist('ladder')
    h = circuits.ladder(y, stages=stages, cje=0.0)
    hb = mthb('HB1', h['freq'], nh)
    hb.block_thomas = bt
    conv, freqs, Vf, _, _ = hb.run(y)
    assert hb._plan.engine() == ('staged-block-thomas' if bt else 'staged-banded')
    assert conv == bool(g['converged'])
    assert hb.last.levels() == g['iters'].tolist()
    if conv:
        assert relmax(hb.V, g['V']) < 1e-9


# ----------------------------------------------------------------------------- boundary
def _integration_stub():
    """the python code block of INTEGRATION.md section 2, verbatim, bound to this repo's device classes and library"""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md = open(os.path.join(root, 'INTEGRATION.md')).read()
    sec = md[md.index('## 2. Bind the C ABI'):]
    code = re.search(r'```python\n(.*?)```', sec, re.S).group(1)
    assert 'def run(self, netlist, V0=None):' in code
    # the stub targets the reference package; the host mirror has the same class names
    code = code.replace('from yalrf.Devices import *', 'from yalrf_b200.Devices import *')
    from yalrf_b200 import _lib
    os.environ['YHB_LIB'] = _lib.LIB_PATH
    ns = {}
    exec(compile(code, 'INTEGRATION.md', 'exec'), ns)
    return ns


def test_integration_stub_verbatim(yb, golden):
    """A maintainer's binding written from include/yhb.h + INTEGRATION.md alone: HB_Diode through it reproduces the
    unmodified reference (iteration path included: the phasors agree to 1e-9)."""
    ns = _integration_stub()

    class RefLikeHB:                       # the attributes the reference class has before run() (:25-30)
        def __init__(self, name, freq, numharmonics):
            self.name, self.freq, self.numharmonics = name, freq, numharmonics
            self.options = {'reltol': 1e-3, 'abstol': 1e-6, 'maxiter': 100}
    RefLikeHB.run = ns['run']
    g = golden('hb_diode')
    y = yb.YalRF('Diode Testbench')
    h = circuits.hb_diode(y)
    hb = RefLikeHB('HB1', h['freq'], h['K'])
    conv, freqs, Vf, _, _ = hb.run(y)
    assert conv
    assert relmax(hb.V, g['V']) < 1e-9 and relmax(Vf, g['Vf']) < 1e-9
    np.testing.assert_array_equal(freqs, g['freqs'])
    assert np.array_equal(hb.IDFT, g['IDFT']) and np.array_equal(hb.DFT, g['DFT'])
    # warm restart through the stub (:330-335): one hb_loop from the converged point
    conv2, _, Vf2, _, _ = hb.run(y, hb.V)
    assert conv2 and relmax(Vf2, g['Vf']) < 1e-6
    # two tones and a BJT circuit through the same stub
    g2 = golden('two_tones_3x3')
    y2 = yb.YalRF('Diode Testbench')
    h2 = circuits.hb_diode_two_tones(y2)
    hb2 = RefLikeHB('HB1', h2['freq'], [3, 3])
    conv, freqs, Vf, _, _ = hb2.run(y2)
    assert conv and relmax(hb2.V, g2['V']) < 1e-9
    np.testing.assert_allclose(freqs, g2['freqs'], rtol=1e-15)


def test_result_struct_size_is_checked(yb):
    """yhb_result without struct_size is refused (no read past the caller's struct); a SHORTER struct from an older
    header works, its missing fields are treated as NULL."""
    from yalrf_b200 import _engine as E, _lib as L
    from yalrf_b200._flatten import ParamBlock
    y = yb.YalRF('Diode Testbench')
    h = circuits.hb_diode(y)
    plan, topo = E.get_plan(y, h['freq'], h['K'])
    pb = ParamBlock(topo, h['freq'], 1).finalize()
    p, o = E.make_params(pb), E.make_options({'reltol': 1e-3, 'abstol': 1e-6, 'maxiter': 100})
    V = np.zeros(plan.W)
    conv = np.zeros(1, dtype=np.int32)
    lib = L.lib()
    r = L.Result()
    r.struct_size = 0
    r.V, r.converged = V.ctypes.data, conv.ctypes.data
    assert lib.yhb_run_host(plan.handle, C.byref(p), C.byref(o), None, C.byref(r)) == -1
    assert b'struct_size' in lib.yhb_last_error()

    class OldResult(C.Structure):          # round-1 layout + struct_size: no Inl, no Ic
        _fields_ = [('struct_size', C.c_size_t)] + [(n, C.c_void_p) for n in
                    ('V', 'Vf', 'converged', 'newton_iters', 'lu_count', 'level_iters', 'level_alpha', 'err')]
    ro = OldResult()
    ro.struct_size = C.sizeof(OldResult)
    ro.V, ro.converged = V.ctypes.data, conv.ctypes.data
    rc = lib.yhb_run_host(plan.handle, C.byref(p), C.byref(o), None, C.cast(C.byref(ro), C.POINTER(L.Result)))
    assert rc == 0, lib.yhb_last_error()
    assert conv[0] == 1 and np.max(np.abs(V)) > 0.1


def test_bjt_ic_side_channel(yb):
    """dev.Ic after run() (:512): the collector-current samples of the last evaluation, against the oracle's."""
    from oracle import hb_oracle as ho
    y = yb.YalRF('x')
    h = circuits.diffamp(y, vin=5e-3)
    hb = mthb('HB1', h['freq'], h['K'])
    assert hb.run(y)[0]
    yo = ho.FlatNetlist()
    ho_h = circuits.diffamp(yo, vin=5e-3)
    o = ho.HBOracle('HB1', h['freq'], h['K'])
    assert o.run(yo)[0]
    assert relmax(hb.V, o.V) < 1e-9
    for q, qo in zip(h['q'], ho_h['q']):
        assert q.Ic.shape == (hb.S,)
        assert relmax(q.Ic, np.asarray(qo.Ic).ravel()) < 1e-8


def test_swept_tone_moves_harmonic_filters(yb):
    """run_sweep(freq=...): an IdealHarmonicFilter at the nominal tone follows the swept tone (its exact-equality test,
    IdealHarmonicFilter.py:34, would otherwise miss at every swept point and the filter would be 1e-12 S everywhere)."""
    fs = np.array([0.9e6, 1.0e6, 1.1e6])
    y = yb.YalRF('x')
    i1 = y.add_iac('I1', 'nx', 'gnd', ac=1.0, freq=1e6)
    y.add_gyrator('G1', 'nx', 'n1', 'gnd', 'gnd', 1)
    y.add_idealharmonicfilter('F1', 'n1', 'n2', 1e6)
    y.add_resistor('R1', 'n2', 'gnd', 50.0)
    d1 = y.add_diode('D1', 'n2', 'gnd')
    hb = mthb('HB1', 1e6, 5)
    res = hb.run_sweep(y, [(i1, 'ac', np.ones(3))], freq=fs)
    assert res.converged.all()
    ref = []
    for f in fs:                            # the same points one by one, every frequency-bearing attribute set by hand
        y2 = yb.YalRF('x')
        y2.add_iac('I1', 'nx', 'gnd', ac=1.0, freq=float(f))
        y2.add_gyrator('G1', 'nx', 'n1', 'gnd', 'gnd', 1)
        y2.add_idealharmonicfilter('F1', 'n1', 'n2', float(f))
        y2.add_resistor('R1', 'n2', 'gnd', 50.0)
        y2.add_diode('D1', 'n2', 'gnd')
        hb2 = mthb('HB1', float(f), 5)
        assert hb2.run(y2)[0]
        ref.append(hb2.V[:, 0])
    assert relmax(res.V, np.array(ref)) < 1e-12
    n2 = y.get_node_idx('n2') - 1
    assert np.all(np.abs(res.Vf[:, n2, 1]) > 0.3)      # the fundamental passes the filter at every swept point


# ----------------------------------------------------------------------------- extensions and adversarial cases
@pytest.mark.parametrize('circuit', ['diffamp', 'ladder', 'ladder_bt'])
def test_exact_jacobian_vs_oracle(yb, circuit):
    """hb.exact_jacobian = True (YHB_FLAG_EXACT_JACOBIAN): dQdV is re-zeroed every iteration instead of
    accumulating (:413 vs :438-440), which makes Newton quadratic for circuits with junction capacitances and lets
    the config-5 ladder with Cje = 1e-12 converge.  Against the oracle with the same switch: identical per-level
    iteration counts, phasors within 1e-9 -- on the fused engine (block-sparse core), the banded staged engine and the
    block-Thomas engine."""
    from oracle import hb_oracle as ho
    if circuit == 'diffamp':
        build, kw, nh = circuits.diffamp, dict(vin=20e-3), None
    else:
        build, kw, nh = circuits.ladder, dict(stages=10, cje=1e-12, ac=0.5), [2, 2]
    yo = ho.FlatNetlist()
    ho_h = build(yo, **kw)
    K = nh if nh is not None else ho_h['K']
    o = ho.HBOracle('HB1', ho_h['freq'], K, exact_jacobian=True)
    conv_o = o.run(yo)[0]
    o_acc = ho.HBOracle('HB1', ho_h['freq'], K)         # the reference's accumulating Jacobian, for contrast
    yo2 = ho.FlatNetlist()
    build(yo2, **kw)
    o_acc.run(yo2)
    y = yb.Netlist('x')
    h = build(y, **kw)
    hb = mthb('HB1', h['freq'], K)
    hb.exact_jacobian = True
    if circuit == 'ladder_bt':
        hb.engine = 'staged'
        hb.block_thomas = True
    elif circuit == 'ladder':
        hb.engine = 'staged'
    conv, freqs, Vf, _, _ = hb.run(y)
    assert conv and conv_o
    assert hb.last.levels() == [l[1] for l in o.level_log]
    assert relmax(hb.V, o.V) < 1e-9
    assert sum(hb.last.levels()) <= o_acc.total_newton          # never slower than the accumulating Jacobian
    if circuit == 'ladder_bt':
        assert hb._plan.engine() == 'staged-block-thomas'


@pytest.mark.parametrize('ac', [0.5, 3.0])
def test_block_pivoting_adversarial(yb, ac):
    """Pivoting restricted to the diagonal node blocks (bs_factor_row) on a core that punishes it: two diode nodes tied
    by 1 mOhm (circuits.stiff_pair).  The block-sparse LU (forced), the dense tensor-core Gauss-Jordan with pivoting
    over all rows, and the oracle (LAPACK dgetrf on the full matrix) must walk the same Newton path: identical
    per-level iteration counts, phasors within 1e-9."""
    from oracle import hb_oracle as ho
    yo = ho.FlatNetlist()
    h = circuits.stiff_pair(yo, ac=ac)
    o = ho.HBOracle('HB1', h['freq'], h['K'])
    conv_o = o.run(yo)[0]
    assert conv_o
    for mode in (True, False):
        y = yb.Netlist('stiff')
        circuits.stiff_pair(y, ac=ac)
        hb = mthb('HB1', h['freq'], h['K'])
        hb.block_sparse = mode
        conv, freqs, Vf, _, _ = hb.run(y)
        assert hb._plan.engine() == 'fused'
        assert (hb._plan.block_sparse() >= 2) if mode else hb._plan.block_sparse() == 0
        assert conv
        assert hb.last.levels() == [l[1] for l in o.level_log], mode
        assert relmax(hb.V, o.V) < 1e-9, mode


# ----------------------------------------------------------------------------- post-processing
def test_convert_to_time_and_two_tone_reconstruction(yb, golden):
    """convert_to_time (:125-135) and the two-tone reconstruction loop of tests/HB_TwoTones.ipynb cell 3 on the GPU
    (yhb_to_time) against the reference's formulas evaluated in numpy."""
    y = yb.YalRF('Diode Testbench')
    h = circuits.hb_diode(y)
    hb = mthb('HB1', h['freq'], h['K'])
    assert hb.run(y)[0]
    xf = hb.get_v('n2')
    t, xt = hb.convert_to_time(xf)
    S = 32 * hb.K
    s = np.arange(S)[:, None]
    k = np.arange(1, hb.K + 1)[None, :]
    ref = xf[0].real + (xf[1:].real[None, :] * np.cos(2. * np.pi * k * s / (S / 2)) -
                        xf[1:].imag[None, :] * np.sin(2. * np.pi * k * s / (S / 2))).sum(axis=1)
    np.testing.assert_array_equal(t, 2 / hb.freq / S * np.linspace(0, S, S))
    assert np.max(np.abs(xt - ref)) < 1e-12 * np.max(np.abs(ref))
    # two tones: quasi-periodic waveform at arbitrary times
    y2 = yb.YalRF('Diode Testbench')
    h2 = circuits.hb_diode_two_tones(y2)
    hb2 = mthb('HB1', h2['freq'], [3, 3])
    conv, freqs, Vf, _, _ = hb2.run(y2)
    assert conv
    tt = np.arange(500) * (1. / (8. * np.max(freqs)))
    x = hb2.to_time(['n1', 'n2'], tt)
    for row, node in zip(x, ['n1', 'n2']):
        v = hb2.get_v(node)
        ref = v[0].real + sum(v[i].real * np.cos(2. * np.pi * freqs[i] * tt) - v[i].imag * np.sin(2. * np.pi * freqs[i] * tt)
                              for i in range(1, len(freqs)))
        assert np.max(np.abs(row - ref)) < 1e-12 * np.max(np.abs(ref))
    assert hb2.to_time('n2', tt).shape == (500,)


# ----------------------------------------------------------------------------- cold solve: DC groups / DC beside HB
_COLD_SCRIPT = r'''
import sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + '/tests')
import circuits, yalrf_b200
from yalrf_b200.Analyses import MultiToneHarmonicBalance
y = yalrf_b200.YalRF('Differential Amplifier')
h = circuits.diffamp(y)
vb = np.repeat(np.linspace(1.0, 1.4, 12), 8)
vb[8] = np.nextafter(vb[8], 2.0)          # one ulp away from its neighbours: must NOT share their DC analysis
vin = np.tile(np.linspace(100e-6, 50e-3, 8), 12)
hb = MultiToneHarmonicBalance('HB1', h['freq'], h['K'])
r = hb.run_sweep(y, [(h['i2'], 'ac', vin / 2), (h['i4'], 'ac', vin / 2), (h['i3'], 'dc', vb), (h['i5'], 'dc', vb)])
np.savez(sys.argv[2], V=r.V, conv=r.converged, it=r.newton_iters, lev=r.level_iters)
'''


def test_cold_solve_modes_bit_identical(tmp_path):
    """The cold solve's three ways of getting the operating points -- one DC analysis per group of circuits with
    bit-identical DC inputs, run beside the HB kernel (default); the same in front of it (YHB_NO_DC_OVERLAP); one DC
    analysis per circuit in front of it (+ YHB_NO_DC_GROUPS) -- give bit-identical results: 96 DiffAmp points on 12
    bias values, one of them moved by one ulp (it must form a group of its own)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for i, env in enumerate(({}, {'YHB_NO_DC_OVERLAP': '1'}, {'YHB_NO_DC_OVERLAP': '1', 'YHB_NO_DC_GROUPS': '1'})):
        out = str(tmp_path / ('m%d.npz' % i))
        e = dict(os.environ)
        e.update(env)
        subprocess.run([sys.executable, '-c', _COLD_SCRIPT, root, out], check=True, env=e, timeout=300)
        outs.append(np.load(out))
    assert int((outs[0]['conv'] > 0).sum()) == 96
    for o in outs[1:]:
        for k in ('V', 'conv', 'it', 'lev'):
            np.testing.assert_array_equal(o[k], outs[0][k])


def test_vsource_extension_hb_diode(yb, golden):
    """Netlist.vsource_extension: tests/HB_Diode.py written with add_vac instead of the hand-made current source + gyrator
    pair gives the reference's golden result (same arrays, same solve)."""
    g = golden('hb_diode')
    y = yb.YalRF('Diode Testbench')
    y.vsource_extension = True
    y.add_vac('V1', 'n1', 'gnd', ac=5, freq=1e6)
    y.add_resistor('R1', 'n1', 'n2', 100)
    d1 = y.add_diode('D1', 'gnd', 'n2')
    d1.options.update({'Is': 1e-15, 'N': 1, 'Area': 1})
    hb = mthb('HB1', 1e6, 10)
    conv, freqs, Vf, _, _ = hb.run(y)
    assert conv
    assert hb.last.levels() == g['iters'].tolist()
    n2 = y.get_node_idx('n2') - 1
    assert relmax(Vf[n2], g['Vf'][2]) < 1e-9        # the golden netlist's nodes: nx, n1, n2


def test_linear_only_netlist_vs_oracle(yb):
    """Edge case: a netlist without any nonlinear device (R, C, L behind a gyrator source): hb_loop still runs its minimum
    of three iterations (:596) on a Jacobian that is the linear admittance alone; the DC analysis takes the linear branch
    (DC.py:90-95).  Single solve and a three-point amplitude sweep against the oracle."""
    from oracle import hb_oracle as ho

    def build(y, ac=2.0):
        i1 = y.add_iac('I1', 'nx', 'gnd', ac=ac, freq=1e6)
        y.add_gyrator('G1', 'nx', 'n1', 'gnd', 'gnd', 1)
        y.add_resistor('R1', 'n1', 'n2', 50)
        y.add_capacitor('C1', 'n2', 'gnd', 2e-9)
        y.add_inductor('L1', 'n2', 'n3', 1e-6)
        y.add_resistor('R2', 'n3', 'gnd', 75)
        return i1

    y = yb.Netlist('linear')
    i1 = build(y)
    hb = mthb('HB1', 1e6, 4)
    conv, freqs, Vf, _, _ = hb.run(y)
    yo = ho.FlatNetlist()
    build(yo)
    o = ho.HBOracle('HB1', 1e6, 4)
    conv_o = o.run(yo)[0]
    assert conv and conv_o
    assert hb.last.levels() == [l[1] for l in o.level_log]
    assert relmax(hb.V, o.V) < 1e-9
    res = hb.run_sweep(y, [(i1, 'ac', np.array([0.5, 1.0, 2.0]))])
    assert res.converged.tolist() == [1, 1, 1]
    assert relmax(res.V[2], o.V[:, 0]) < 1e-9
    assert relmax(res.V[0] * 4.0, res.V[2]) < 1e-12          # linear circuit: the response scales with the drive


def test_diode_charge_extension_vs_oracle(yb):
    """hb.diode_charge = True -- diodes carry the charge of Diode.calc_oppoint (Diode.py:207-235: Cp, Tt,
    Cj0 / Vj / M / Fc) in harmonic balance, which the reference's MTHB leaves out (:441-475).  Opt-in extension, parity
    unpinned by the reference: checked against the oracle with the same extension (single tone and two tones), identical
    iteration counts, phasors within 1e-9; off by default (the golden HB_Diode result is untouched)."""
    from oracle import hb_oracle as ho
    strong = {'Cj0': 2e-9, 'Vj': 0.8, 'M': 0.4, 'Fc': 0.5, 'Tt': 1e-7, 'Cp': 1e-10, 'Area': 2.0}
    mild = {'Cj0': 2e-10, 'Vj': 0.8, 'M': 0.4, 'Fc': 0.5, 'Tt': 1e-8, 'Cp': 1e-11, 'Area': 2.0}   # the two-tone ladder backs off twice
    for build, nh, charge in ((circuits.hb_diode, None, strong), (circuits.hb_diode_two_tones, [3, 3], mild)):
        y = yb.YalRF('Diode Testbench')
        h = build(y)
        h['d1'].options.update(charge)
        hb = mthb('HB1', h['freq'], nh if nh is not None else h['K'])
        hb.diode_charge = True
        conv, freqs, Vf, _, _ = hb.run(y)
        yo = ho.FlatNetlist()
        ho_h = build(yo)
        ho_h['d1'].options.update(charge)
        o = ho.HBOracle('HB1', h['freq'], nh if nh is not None else h['K'], diode_charge=True)
        conv_o = o.run(yo)[0]
        assert conv and conv_o
        assert hb.last.levels() == [l[1] for l in o.level_log]
        assert relmax(hb.V, o.V) < 1e-9
        hb0 = mthb('HB1', h['freq'], nh if nh is not None else h['K'])       # extension off: the charge parameters are ignored
        assert hb0.run(y)[0]
        assert relmax(hb0.V, hb.V) > 1e-4


def test_diode_rs_extension_vs_oracle(yb):
    """hb.diode_rs = True -- a diode's series resistance counts (internal node + resistor Rs / Area; the
    reference's get_mthb_params ignores Rs).  Against the oracle on the netlist written out by hand; the netlist's own
    node indices do not move; off by default."""
    from oracle import hb_oracle as ho
    y = yb.YalRF('Diode Testbench')
    h = circuits.hb_diode(y)
    h['d1'].options.update({'Rs': 5.0, 'Area': 2.0})
    hb = mthb('HB1', h['freq'], h['K'])
    hb.diode_rs = True
    conv, freqs, Vf, _, _ = hb.run(y)
    yo = ho.FlatNetlist()
    yo.add_iac('I1', 'nx', 'gnd', ac=5.0, freq=1e6)
    yo.add_gyrator('G1', 'nx', 'n1', 'gnd', 'gnd', 1)
    yo.add_resistor('R1', 'n1', 'n2', 100)
    yo.add_resistor('D1.Rs', 'gnd', 'D1.rs', 2.5)
    d = yo.add_diode('D1', 'D1.rs', 'n2')
    d.options.update({'Is': 1e-15, 'N': 1, 'Area': 2.0, 'Rs': 0.0})
    o = ho.HBOracle('HB1', h['freq'], h['K'])
    assert conv and o.run(yo)[0]
    assert hb.N == 4 and hb.get_node_idx('n2') == 2 and hb.get_node_idx('D1.rs') == 3
    assert hb.last.levels() == [l[1] for l in o.level_log]
    assert relmax(hb.V, o.V) < 1e-9
    hb0 = mthb('HB1', h['freq'], h['K'])
    assert hb0.run(y)[0] and hb0.N == 3
    assert relmax(hb0.Vf[2], Vf[2]) > 1e-3               # Rs = 2.5 ohm in series with the diode changes n2 visibly
    # an Is sweep of the original diode reaches its twin
    res = hb.run_sweep(y, [(h['d1'], 'Is', np.array([1e-15, 2e-15]))])
    assert res.converged.tolist() == [1, 1] and relmax(res.V[0], o.V[:, 0]) < 1e-9 and relmax(res.V[1], res.V[0]) > 1e-6
