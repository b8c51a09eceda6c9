"""Pins the numpy oracle (oracle/hb_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/*.npz, written by tests/golden/make_golden.py)."""
import numpy as np
import pytest

import circuits
from oracle import hb_oracle as ho


def relmax(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def levels(hb):
    return [l[1] for l in hb.level_log]


def test_bjt_model(golden):
    g = golden('bjt_model')
    opt = dict(zip(g['opt_names'].tolist(), g['opt_vals'].tolist()))
    vin, vold = g['vin'], g['vold']
    out = ho.bjt_hb_params(opt, vin[:, 0], vin[:, 1], vin[:, 2], vold[:, 0], vold[:, 1], vold[:, 2])
    out = np.column_stack(out)
    ref = g['out']
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    ok = ~np.isnan(ref)
    err = np.abs(out[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-300)
    assert err.max() < 1e-13


def test_diode_model(golden):
    g = golden('diode_model')
    opt = dict(zip(g['opt_names'].tolist(), g['opt_vals'].tolist()))
    gd, Id = ho.diode_mthb_params(opt, g['vd'], g['vdold'])
    ref = g['out']
    assert np.max(np.abs(gd - ref[:, 0]) / np.abs(ref[:, 0])) < 1e-13
    assert np.max(np.abs(Id - ref[:, 1]) / np.maximum(np.abs(ref[:, 1]), 1e-300)) < 1e-12


@pytest.mark.parametrize('tag,ac,opts', [('hb_diode', 5.0, {}), ('hb_diode_tight', 5.0, {'reltol': 1e-9, 'abstol': 1e-12}),
                                         ('hb_diode_ac0p5', 0.5, {}), ('hb_diode_ac2', 2.0, {})])
def test_hb_diode(golden, tag, ac, opts):
    g = golden(tag)
    y = ho.FlatNetlist()
    h = circuits.hb_diode(y, ac=ac)
    hb = ho.HBOracle('HB1', h['freq'], h['K'])
    hb.options.update(opts)
    if tag == 'hb_diode':
        hb.trace = []
    conv, freqs, Vf, _, _ = hb.run(y)
    assert conv and bool(g['converged'])
    assert levels(hb) == g['iters'].tolist()
    np.testing.assert_allclose([l[0] for l in hb.level_log], g['alphas'], rtol=1e-15)
    assert relmax(hb.V, g['V']) < 1e-11
    assert relmax(Vf, g['Vf']) < 1e-11
    np.testing.assert_allclose(freqs, g['freqs'], rtol=0, atol=0)
    if tag == 'hb_diode':
        assert np.max(np.abs(hb.DFT - g['DFT'])) < 1e-15
        assert np.array_equal(hb.IDFT, g['IDFT'])
        for i in (0, 5):
            t = hb.trace[i if i == 0 else 5 + 1]  # trace also holds the non-LU final iteration of level 0
            assert relmax(t['J'], g['J%d' % i]) < 1e-12
            assert relmax(t['F'], g['F%d' % i]) < 1e-9
            assert relmax(t['dV'], g['dV%d' % i]) < 1e-9


@pytest.mark.parametrize('tag,nh', [('two_tones_3x3', [3, 3]), ('two_tones_8x8', [8, 8])])
def test_two_tones(golden, tag, nh):
    g = golden(tag)
    y = ho.FlatNetlist()
    h = circuits.hb_diode_two_tones(y)
    hb = ho.HBOracle('HB1', h['freq'], nh)
    hb.trace = [] if tag == 'two_tones_3x3' else None
    conv, freqs, Vf, _, _ = hb.run(y)
    assert conv
    assert levels(hb) == g['iters'].tolist()
    # S = 289 amplifies BLAS summation-order differences to ~2e-10 (measured); 1e-9 is the parity bar
    assert relmax(hb.V, g['V']) < 1e-9
    assert relmax(Vf, g['Vf']) < 1e-9
    np.testing.assert_allclose(freqs, g['freqs'], rtol=1e-15)
    if hb.trace is not None:
        t = hb.trace[0]
        assert relmax(t['J'], g['J0']) < 1e-12
        assert relmax(t['F'], g['F0']) < 1e-12
        assert relmax(t['dV'], g['dV0']) < 1e-10


def test_hb_bjt(golden):
    g = golden('hb_bjt')
    y = ho.FlatNetlist()
    h = circuits.hb_bjt(y)
    hb = ho.HBOracle('HB1', h['freq'], h['K'])
    conv, freqs, Vf, _, _ = hb.run(y)
    assert conv
    assert levels(hb) == g['iters'].tolist()
    assert relmax(hb.V, g['V']) < 1e-10


def test_diffamp_cold_grid(golden):
    g = golden('diffamp_cold_grid')
    k = 0
    for vb in g['vbs']:
        for vin in g['vins']:
            y = ho.FlatNetlist()
            h = circuits.diffamp(y, vin=vin / 2, vbias=vb)
            hb = ho.HBOracle('HB1', h['freq'], h['K'])
            rec = (vb == 1.2 and vin == 20e-3)
            hb.trace = [] if rec else None
            conv, freqs, Vf, _, _ = hb.run(y)
            assert conv == bool(g['converged'][k])
            its = g['iters'][k]
            assert levels(hb) == its[its > 0].tolist()
            if conv:
                assert relmax(hb.V, g['V'][k]) < 1e-9, (vb, vin)
                for n in range(Vf.shape[0]):
                    assert relmax(Vf[n], g['Vf'][k][n]) < 1e-9
            if rec:
                lus = [t for t in hb.trace if t['dV'] is not None]
                for i in (0, 1, 7):
                    assert relmax(lus[i]['J'], g['J%d' % i]) < 1e-11
                    assert relmax(lus[i]['F'], g['F%d' % i]) < 1e-8
                    assert relmax(lus[i]['dV'], g['dV%d' % i]) < 1e-7
            k += 1


def test_diffamp_warm_sweep(golden):
    g = golden('diffamp_warm_sweep')
    y = ho.FlatNetlist()
    h = circuits.diffamp(y)
    hb = ho.HBOracle('HB1', h['freq'], h['K'])
    V0 = None
    for k, vin in enumerate(g['vins']):
        h['i2'].ac = vin / 2
        h['i4'].ac = vin / 2
        conv, freqs, Vf, _, _ = hb.run(y, V0)
        V0 = hb.V
        assert conv
        its = g['iters'][k]
        assert levels(hb) == its[its > 0].tolist()
        assert relmax(hb.V, g['V'][k]) < 1e-9
        assert relmax(Vf, g['Vf'][k]) < 1e-9


def test_peltz_probe_fixed(golden):
    g = golden('peltz_probe_fixed')
    y = ho.FlatNetlist()
    circuits.peltz(y)
    y.add_iac('P.I', 'P.n1', 'gnd', ac=0.7, freq=71e3)
    y.add_gyrator('P.G', 'P.n1', 'P.n2', 'gnd', 'gnd', 1)
    y.add_idealharmonicfilter('P.IHF', 'P.n2', 'nb', 71e3)
    hb = ho.HBOracle('HB1', 71e3, 10)
    hb.trace = []
    conv, freqs, Vf, _, _ = hb.run(y)
    assert conv
    assert levels(hb) == g['iters'].tolist()
    assert relmax(hb.trace[0]['J'], g['J0']) < 1e-12
    assert relmax(hb.V, g['V']) < 1e-9


def test_peltz_oscillator(golden):
    g = golden('peltz_osc')
    y = ho.FlatNetlist()
    h = circuits.peltz(y)
    hb = ho.HBOracle('HB1')
    conv, freqs, Vf, _, _ = hb.run_oscillator(y, h['f0'], h['K'], h['V0'], h['node'])
    assert conv
    # Nelder-Mead amplifies rounding differences over ~270 evaluations; the root itself is what is pinned.
    assert abs(hb.fosc - float(g['fosc'])) / float(g['fosc']) < 1e-9
    assert abs(hb.Vosc - float(g['Vosc'])) / float(g['Vosc']) < 1e-9
    assert relmax(Vf, g['Vf']) < 1e-8


def test_vanderpol_oscillator(golden):
    g = golden('vanderpol_osc')
    y = ho.FlatNetlist()
    h = circuits.vanderpol(y)
    hb = ho.HBOracle('HB1')
    conv, freqs, Vf, _, _ = hb.run_oscillator(y, h['f0'], h['K'], h['V0'], h['node'])
    assert conv
    assert abs(hb.fosc - float(g['fosc'])) / float(g['fosc']) < 1e-8
    assert abs(hb.Vosc - float(g['Vosc'])) / float(g['Vosc']) < 1e-8


def test_ladder(golden):
    g = golden('ladder8_2x2')
    y = ho.FlatNetlist()
    h = circuits.ladder(y, stages=8, cje=0.0)
    hb = ho.HBOracle('HB1', h['freq'], [2, 2])
    conv, freqs, Vf, _, _ = hb.run(y)
    assert conv == bool(g['converged'])
    assert levels(hb) == g['iters'].tolist()
    assert relmax(hb.V, g['V']) < 1e-9
