"""CPU-only tests: the C-ABI library loads and exports what include/yhb.h declares, the host mirror of the
reference interface behaves like the reference's classes, and nothing computes without a GPU."""
import os
import re

import numpy as np
import pytest

import circuits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    import ctypes
    from yalrf_b200 import _lib
    lib = _lib.lib()
    hdr = open(os.path.join(ROOT, 'include', 'yhb.h')).read()
    names = set(re.findall(r'\b(yhb_[a-z_0-9]+)\s*\(', hdr))
    assert names >= set(_lib.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.yhb_version() == 210
    assert _lib.DIODE_NPAR == 16 and _lib.BJT_NPAR == 32


def test_header_enums_match_binding():
    from yalrf_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'yhb.h')).read()
    body = re.search(r'YHB_BJT_TEMP = 0,(.*?)YHB_BJT_NPAR', hdr, re.S).group(1)
    names = ['Temp'] + [x.strip().split('_')[-1].capitalize() for x in re.findall(r'YHB_BJT_[A-Z]+', body)]
    assert [n.lower() for n in names] == [n.lower() for n in _lib.BJT_PARAMS]
    body = re.search(r'YHB_DIODE_TEMP = 0,(.*?)YHB_DIODE_NPAR', hdr, re.S).group(1)
    names = ['Temp'] + [x.strip().split('_')[-1].capitalize() for x in re.findall(r'YHB_DIODE_[A-Z0-9]+', body)]
    assert [n.lower() for n in names] == [n.lower() for n in _lib.DIODE_PARAMS]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import yalrf_b200
    from yalrf_b200.Analyses import MultiToneHarmonicBalance
    y = yalrf_b200.YalRF('x')
    h = circuits.hb_diode(y)
    hb = MultiToneHarmonicBalance('HB1', h['freq'], h['K'])
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        hb.run(y)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'yalrf_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert 'oracle' not in src.replace('oracle/', ''), f  # comments may cite the directory only
                assert 'import oracle' not in src and 'from oracle' not in src


def test_netlist_api_matches_reference_semantics():
    import yalrf_b200
    y = yalrf_b200.YalRF('Differential Amplifier')
    h = circuits.diffamp(y)
    assert y.get_num_nodes() == 11
    assert y.get_node_idx('gnd') == 0 and y.get_node_idx('nx') == 1 and y.get_node_idx('nvcc') == 2
    assert len(y.get_nonlinear_devices()) == 4 and len(y.get_linear_devices()) == 11
    q1, q2 = h['q'][0], h['q'][1]
    assert q1.options is q2.options                 # shared dict semantics (tests/HB_DiffAmp.py:67-69)
    c = y.copy()
    assert c.devices is not y.devices and c.devices[0] is y.devices[0]   # shallow copy (Netlist.py:889-897)
    assert h['i4'].phase == pytest.approx(np.pi)    # degrees -> radians (CurrentSource.py:12)
    assert y.get_num_vsources() == 0
    p = yalrf_b200.Netlist('p')
    circuits.peltz(p)
    assert p.get_num_vsources() == 1                # the inductor's DC branch row
    with pytest.raises(NotImplementedError):
        y.add_vdc('V1', 'a', 'gnd', 1.0)


def test_flatten_and_sweep_arrays():
    import yalrf_b200
    from yalrf_b200 import _lib as L
    from yalrf_b200._flatten import ParamBlock, Topology, dft_tables, source_kinds
    y = yalrf_b200.YalRF('x')
    h = circuits.diffamp(y)
    topo = Topology(y, 1, (10,))
    assert topo.N == 10 and len(topo.bjts) == 4 and len(topo.src) == 6 and len(topo.lin) == 5
    kinds = source_kinds(topo, 10e6)
    assert kinds == [L.SRC_DC, L.SRC_AC1, L.SRC_DC, L.SRC_AC1, L.SRC_DC, L.SRC_DC]
    with pytest.raises(ValueError):
        source_kinds(topo, 11e6)
    pb = ParamBlock(topo, 10e6, 4)
    pb.set(h['i2'], 'ac', [1., 2., 3., 4.])
    pb.set(h['q'][0], 'Is', [1e-15, 2e-15, 3e-15, 4e-15])   # shared options dict: applies to all four BJTs
    pb.finalize()
    assert pb.src_val.shape == (4, 6, 3)
    np.testing.assert_allclose(pb.src_val[:, 1, 1], [1, 2, 3, 4])
    np.testing.assert_allclose(pb.src_val[:, 3, 1], -60e-3)          # I4: phase 180 deg
    assert abs(pb.src_val[0, 3, 2]) < 1e-16
    np.testing.assert_allclose(pb.bjt_par[:, :, L.BJT_PARAMS.index('Is')], np.outer([1e-15, 2e-15, 3e-15, 4e-15], np.ones(4)))
    IDFT, DFT = dft_tables(10)
    assert np.max(np.abs(DFT @ IDFT - np.eye(21))) < 1e-14


def test_dft_tables_match_reference_fixture(golden):
    from yalrf_b200._flatten import dft_tables
    g = golden('hb_diode')
    IDFT, DFT = dft_tables(10)
    assert np.array_equal(IDFT, g['IDFT'])
    assert np.array_equal(DFT, g['DFT'])


def _struct_fields(hdr, name):
    end = hdr.index('} ' + name + ';')
    body = hdr[hdr.rindex('typedef struct {', 0, end) + len('typedef struct {'):end]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    out = []
    for decl in body.split(';'):
        decl = decl.strip()
        if decl:
            for nm in decl.split(','):
                out.append(re.sub(r'[^A-Za-z0-9_]', ' ', nm).split()[-1])
    return out


def test_binding_and_integration_stub_match_header_structs():
    """ctypes structs of the host mirror AND of the maintainer stub in INTEGRATION.md have the header's fields, in the
    header's order (a struct with a missing tail field is exactly the undefined behaviour round 1 shipped)."""
    import ctypes
    from yalrf_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'yhb.h')).read()
    md = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    code = re.search(r'```python\n(.*?)```', md[md.index('## 2. Bind the C ABI'):], re.S).group(1)
    code = code.replace('from yalrf.Devices import *', 'from yalrf_b200.Devices import *')
    os.environ['YHB_LIB'] = _lib.LIB_PATH
    ns = {}
    exec(compile(code, 'INTEGRATION.md', 'exec'), ns)
    for cname, pyname in [('yhb_result', 'Result'), ('yhb_params', 'Params'), ('yhb_options', 'Options'),
                          ('yhb_plan_desc', 'PlanDesc')]:
        want = _struct_fields(hdr, cname)
        for cls in (getattr(_lib, pyname), ns[pyname]):
            assert [f[0] for f in cls._fields_] == want, (cname, cls)
        assert ctypes.sizeof(getattr(_lib, pyname)) == ctypes.sizeof(ns[pyname])
    assert _lib.Result().struct_size == ctypes.sizeof(_lib.Result) == 8 * 11


def test_linear_mthb_stamps_match_oracle():
    """add_mthb_stamps of the host mirror's linear devices (the reference's calc_Y protocol, :686-695) against the
    oracle's calc_Y: identical dense Y for the Peltz oscillator (inductor, capacitor, gyrator), the DiffAmp and HB_Diode."""
    import yalrf_b200
    from oracle import hb_oracle as ho
    for build, f in ((circuits.peltz, 80e3), (circuits.diffamp, None), (circuits.hb_diode, None)):
        y = yalrf_b200.Netlist('x')
        h = build(y)
        y.add_idealharmonicfilter('F', 'nq', 'gnd', float(f or h['freq']))
        yo = ho.FlatNetlist()
        build(yo)
        yo.add_idealharmonicfilter('F', 'nq', 'gnd', float(f or h['freq']))
        o = ho.HBOracle('HB1', float(f or h['freq']), h['K'])
        o.netlist = yo
        o.lin_devs, o.nonlin_devs = yo.get_linear_devices(), yo.get_nonlinear_devices()
        o.setup_grid()
        Yo = o.calc_Y()
        S, N = 2 * h['K'] + 1, y.get_num_nodes() - 1
        Y = np.zeros((N * S, N * S))
        for dev in y.get_linear_devices():
            if hasattr(dev, 'add_mthb_stamps'):                       # :692
                for k in range(h['K'] + 1):
                    dev.add_mthb_stamps(Y, S, abs(o.freqs[k]), k)
        assert np.array_equal(Y, Yo)


def test_vsource_extension_is_the_gyrator_idiom():
    """Netlist.vsource_extension (off by default: add_vdc / add_vac raise like before): a voltage source is entered as the
    current source + unit gyrator pair the reference's HB scripts write by hand, so the flattened arrays equal those of
    the hand-written netlist; the returned handle forwards edits to the internal current source."""
    import pytest
    import yalrf_b200
    from yalrf_b200._flatten import ParamBlock, Topology
    y0 = yalrf_b200.Netlist('x')
    with pytest.raises(NotImplementedError):
        y0.add_vdc('V1', 'n1', 'gnd', 3.0)
    ya = yalrf_b200.Netlist('a')                       # hand-written (tests/HB_Diode.py:12-15 + a floating DC source)
    ya.add_iac('V1.I', 'V1.x', 'gnd', ac=5, freq=1e6)
    ya.add_gyrator('V1.G', 'V1.x', 'n1', 'gnd', 'gnd', 1)
    ya.add_resistor('R1', 'n1', 'n2', 100)
    ya.add_idc('V2.I', 'V2.x', 'gnd', dc=0.3)
    ya.add_gyrator('V2.G', 'V2.x', 'n2', 'n3', 'gnd', 1)
    ya.add_diode('D1', 'gnd', 'n3')
    yb = yalrf_b200.Netlist('b')
    yb.vsource_extension = True
    v1 = yb.add_vac('V1', 'n1', 'gnd', ac=5, freq=1e6)
    yb.add_resistor('R1', 'n1', 'n2', 100)
    v2 = yb.add_vdc('V2', 'n2', 'n3', 0.3)
    yb.add_diode('D1', 'gnd', 'n3')
    ta, tb = Topology(ya, 1, (10,)), Topology(yb, 1, (10,))
    from yalrf_b200._flatten import source_kinds
    assert ta.key(source_kinds(ta, 1e6, True), 0) == tb.key(source_kinds(tb, 1e6, True), 0)
    pa, pb = ParamBlock(ta, 1e6, 1).finalize(), ParamBlock(tb, 1e6, 1).finalize()
    for n in ('lin_val', 'src_val', 'diode_par'):
        assert np.array_equal(getattr(pa, n), getattr(pb, n)), n
    v2.dc = 0.5
    v1.ac = 2.0
    assert yb.get_device('V2.I').dc == 0.5 and yb.get_device('V1.I').ac == 2.0 and v2.dc == 0.5


def test_diode_rs_expansion_is_the_hand_written_netlist():
    """hb.diode_rs: the expanded netlist flattens to the arrays of the netlist written out by hand (resistor Rs / Area to an
    internal node behind the netlist's own nodes + a diode with Rs = 0); option sweeps of the original diode reach the twin."""
    import pytest
    import yalrf_b200
    from yalrf_b200.Analyses.MultiToneHarmonicBalance import _expand_diode_rs
    from yalrf_b200._flatten import ParamBlock, Topology, source_kinds
    ya = yalrf_b200.Netlist('a')
    h = circuits.hb_diode(ya)
    h['d1'].options.update({'Rs': 5.0, 'Area': 2.0})
    ye = _expand_diode_rs(ya)
    assert ya.get_num_nodes() == 4 and ye.get_num_nodes() == 5 and ye.get_node_idx('n2') == ya.get_node_idx('n2')
    yb = yalrf_b200.Netlist('b')
    yb.add_iac('I1', 'nx', 'gnd', ac=5.0, freq=1e6)
    yb.add_gyrator('G1', 'nx', 'n1', 'gnd', 'gnd', 1)
    yb.add_resistor('R1', 'n1', 'n2', 100)
    yb.add_resistor('D1.Rs', 'gnd', 'D1.rs', 2.5)
    d = yb.add_diode('D1', 'D1.rs', 'n2')
    d.options.update({'Is': 1e-15, 'N': 1, 'Area': 2.0, 'Rs': 0.0})
    te, tb = Topology(ye, 1, (10,)), Topology(yb, 1, (10,))
    assert te.key(source_kinds(te, 1e6, True), 0) == tb.key(source_kinds(tb, 1e6, True), 0)
    pe, pb = ParamBlock(te, 1e6, 2), ParamBlock(tb, 1e6, 2)
    pe.set(h['d1'], 'Is', [1e-15, 3e-15])
    pb.set(d, 'Is', [1e-15, 3e-15])
    for n in ('lin_val', 'src_val', 'diode_par'):
        assert np.array_equal(getattr(pe.finalize(), n), getattr(pb.finalize(), n)), n
    with pytest.raises(KeyError):
        pe.set(h['d1'], 'Rs', [1.0, 2.0])
