"""Stage-level parity (SURVEY 8c protocol i): identical inputs into the CUDA stage entry points and the
oracle / reference fixtures; compare vt, device outputs, J, F, dV."""
import ctypes as C

import numpy as np
import pytest

import circuits

pytestmark = pytest.mark.gpu


def relmax(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def _dev(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _params(pb):
    from yalrf_b200 import _lib as L
    pb.finalize()
    t = {n: _dev(getattr(pb, n)) for n in ('tones', 'lin_val', 'src_val', 'diode_par', 'bjt_par', 'cubic_par')}
    p = L.Params()
    p.batch = pb.B
    for n, v in t.items():
        setattr(p, n, v.data_ptr())
    return p, t


def test_bjt_model_golden(golden):
    """BJT.get_hb_params outputs of the unmodified reference on 512 seeded bias points."""
    import torch
    import yalrf_b200
    from yalrf_b200 import _engine as E, _lib as L
    from yalrf_b200._flatten import ParamBlock
    g = golden('bjt_model')
    opt = dict(zip(g['opt_names'].tolist(), g['opt_vals'].tolist()))
    y = yalrf_b200.Netlist('x')
    q = y.add_bjt('Q', 'b', 'c', 'e')
    q.options.update(opt)
    plan, topo = E.get_plan(y, 1e6, 10)
    S = plan.S
    n = len(g['vin'])
    B = (n + S - 1) // S
    vin = np.zeros((B * S, 3))
    vold = np.zeros((B * S, 3))
    vin[:n], vold[:n] = g['vin'], g['vold']
    vt = np.ascontiguousarray(vin.reshape(B, S, 3).transpose(0, 2, 1))      # [B][node][S]
    vp = np.ascontiguousarray(vold.reshape(B, S, 3).transpose(0, 2, 1))
    pb = ParamBlock(topo, 1e6, B)
    p, keep = _params(pb)
    out = torch.zeros((B, 1, 14, S), dtype=torch.float64, device='cuda')
    dvt, dvp = _dev(vt), _dev(vp)
    L.check(L.lib().yhb_stage_device_eval(plan.handle, C.byref(p), dvt.data_ptr(), dvp.data_ptr(), None,
                                          out.data_ptr(), None, None))
    torch.cuda.synchronize()
    o = out.cpu().numpy()[:, 0].transpose(0, 2, 1).reshape(B * S, 14)[:n]
    ref = g['out']
    assert np.array_equal(np.isnan(o), np.isnan(ref))
    ok = ~np.isnan(ref)
    err = np.abs(o[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-300)
    assert err.max() < 1e-12, err.max()


def test_diode_model_golden(golden):
    import torch
    import yalrf_b200
    from yalrf_b200 import _engine as E, _lib as L
    from yalrf_b200._flatten import ParamBlock
    g = golden('diode_model')
    opt = dict(zip(g['opt_names'].tolist(), g['opt_vals'].tolist()))
    y = yalrf_b200.Netlist('x')
    d = y.add_diode('D', 'a', 'gnd')
    d.options.update(opt)
    plan, topo = E.get_plan(y, 1e6, 10)
    S = plan.S
    n = len(g['vd'])
    B = (n + S - 1) // S
    vd = np.zeros(B * S)
    vo = np.zeros(B * S)
    vd[:n], vo[:n] = g['vd'], g['vdold']
    pb = ParamBlock(topo, 1e6, B)
    p, keep = _params(pb)
    out = torch.zeros((B, 1, 2, S), dtype=torch.float64, device='cuda')
    dvt, dvp = _dev(vd.reshape(B, 1, S)), _dev(vo.reshape(B, 1, S))
    L.check(L.lib().yhb_stage_device_eval(plan.handle, C.byref(p), dvt.data_ptr(), dvp.data_ptr(), out.data_ptr(),
                                          None, None, None))
    torch.cuda.synchronize()
    o = out.cpu().numpy()[:, 0].transpose(0, 2, 1).reshape(B * S, 2)[:n]
    ref = g['out']
    err = np.abs(o - ref) / np.maximum(np.abs(ref), 1e-300)
    assert err.max() < 1e-12, err.max()


def _oracle_trace(build, freq_key='freq', nh=None, **kw):
    from oracle import hb_oracle as ho
    yo = ho.FlatNetlist()
    h = build(yo, **kw)
    o = ho.HBOracle('HB1', h['freq'], nh if nh is not None else h['K'])
    o.trace = []
    o.run(yo)
    return o


@pytest.mark.parametrize('name,nh,kw', [('diffamp', None, dict(vin=10e-3)), ('hb_diode', None, {}),
                                        ('hb_diode_two_tones', [3, 3], {})])
def test_assemble_and_lu_vs_oracle(name, nh, kw):
    """J, F of Newton iterations of an oracle run (limiter reference and accumulated dQdV included) and the LU step."""
    import torch
    import yalrf_b200
    from yalrf_b200 import _engine as E, _lib as L
    from yalrf_b200._flatten import ParamBlock
    build = getattr(circuits, name)
    o = _oracle_trace(build, nh=nh, **kw)
    y = yalrf_b200.Netlist('x')
    h = build(y, **kw)
    plan, topo = E.get_plan(y, h['freq'], nh if nh is not None else h['K'])
    pb = ParamBlock(topo, h['freq'], 1)
    p, keep = _params(pb)
    lib = L.lib()
    nbytes = lib.yhb_workspace_bytes(plan.handle, 1)
    ws = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    W = plan.W
    nb = max(len(topo.bjts), 1)
    cacc = torch.zeros((1, nb, 4, plan.S), dtype=torch.float64, device='cuda')
    # first level of the ladder: alpha = 0.1; iterations until the level's hb_loop returned
    alpha = _dev(np.array([0.1]))
    J = torch.zeros((1, W, W), dtype=torch.float64, device='cuda')
    F = torch.zeros((1, W), dtype=torch.float64, device='cuda')
    x = torch.zeros((1, W), dtype=torch.float64, device='cuda')
    nlev0 = o.level_log[0][1]
    for it in range(nlev0):
        t = o.trace[it]
        V = _dev(t['V'].reshape(1, W))
        vtp = None if it == 0 else _dev(t['vtprev'].reshape(1, W))
        L.check(lib.yhb_stage_assemble(plan.handle, C.byref(p), ws.data_ptr(), nbytes, V.data_ptr(),
                                       None if vtp is None else vtp.data_ptr(), alpha.data_ptr(), cacc.data_ptr(),
                                       J.data_ptr(), F.data_ptr(), None))
        torch.cuda.synchronize()
        Jh, Fh = J.cpu().numpy()[0], F.cpu().numpy()[0]
        assert relmax(Jh, t['J']) < 1e-12, ('J', it)
        assert np.max(np.abs(Fh - t['F'][:, 0])) < 1e-12 * max(1.0, np.max(np.abs(t['F']))) + 1e-9 * np.max(np.abs(t['F'])), ('F', it)
        if t['dV'] is not None:
            L.check(lib.yhb_stage_lu(1, W, J.data_ptr(), F.data_ptr(), x.data_ptr(), None))
            torch.cuda.synchronize()
            xh = x.cpu().numpy()[0]
            assert relmax(xh, np.linalg.solve(Jh, Fh)) < 1e-9, ('solve', it)
            if np.max(np.abs(t['F'])) > 1e-6:   # below that F is rounding noise, and so is the reference's dV
                assert relmax(xh, t['dV'][:, 0]) < 1e-8, ('dV', it)


def test_stage_ifft_vs_numpy():
    import torch
    import yalrf_b200
    from yalrf_b200 import _engine as E, _lib as L
    y = yalrf_b200.Netlist('x')
    h = circuits.diffamp(y)
    plan, topo = E.get_plan(y, h['freq'], h['K'])
    rng = np.random.default_rng(0)
    B = 7
    V = rng.normal(size=(B, plan.N, plan.S))
    vt = torch.zeros((B, plan.W), dtype=torch.float64, device='cuda')
    dV = _dev(V)
    L.check(L.lib().yhb_stage_ifft(plan.handle, B, dV.data_ptr(), vt.data_ptr(), None))
    torch.cuda.synchronize()
    ref = np.einsum('sc,bnc->bns', plan.IDFT, V).reshape(B, plan.W)
    assert np.max(np.abs(vt.cpu().numpy() - ref)) < 1e-13


def test_stage_lu_random():
    import torch
    from yalrf_b200 import _lib as L
    rng = np.random.default_rng(1)
    for W in (5, 63, 210):
        B = 9
        A = rng.normal(size=(B, W, W))
        b = rng.normal(size=(B, W))
        x = torch.zeros((B, W), dtype=torch.float64, device='cuda')
        dA, db = _dev(A), _dev(b)
        L.check(L.lib().yhb_stage_lu(B, W, dA.data_ptr(), db.data_ptr(), x.data_ptr(), None))
        torch.cuda.synchronize()
        ref = np.linalg.solve(A, b[:, :, None])[:, :, 0]
        assert relmax(x.cpu().numpy(), ref) < 1e-9


def test_stage_gj_random():
    """The fused engine's register-tiled Gauss-Jordan solver against LAPACK on random and badly scaled systems."""
    import torch
    from yalrf_b200 import _lib as L
    rng = np.random.default_rng(2)
    for n in (1, 3, 4, 5, 21, 31, 32, 42, 49, 63, 64, 84, 85, 95, 96, 100, 127):
        B = 5
        A = rng.normal(size=(B, n, n))
        A[1] *= np.exp(rng.uniform(-20, 20, size=(n, 1)))      # wild row scaling
        if n > 2:
            A[2][np.arange(n), np.arange(n)] = 0.0              # zero diagonal: pivoting is mandatory
        b = rng.normal(size=(B, n))
        x = torch.zeros((B, n), dtype=torch.float64, device='cuda')
        dA, db = _dev(A), _dev(b)
        L.check(L.lib().yhb_stage_gj(B, n, dA.data_ptr(), db.data_ptr(), x.data_ptr(), None))
        torch.cuda.synchronize()
        ref = np.linalg.solve(A, b[:, :, None])[:, :, 0]
        for k in range(B):
            assert relmax(x.cpu().numpy()[k], ref[k]) < 1e-9 * max(1.0, np.linalg.cond(A[k]) * 1e-4), (n, k)


def test_stage_gjp_random():
    """The staged engines' panel Gauss-Jordan solver (k_gjp_panel + DMMA k_gjp_update, yhb_dense.cuh) against LAPACK:
    sizes around the panel width and the 8-row tiles, the config-3 core (289) and a config-5 block row (1089 pivots
    with carried columns), wild row scaling, zero diagonals, several right-hand sides."""
    import torch
    from yalrf_b200 import _lib as L
    rng = np.random.default_rng(7)
    for n, nrhs, B in [(1, 1, 2), (5, 1, 3), (15, 2, 3), (16, 1, 3), (17, 3, 3), (33, 1, 3), (100, 5, 3), (289, 1, 4),
                       (300, 40, 2), (1089, 1090, 1)]:
        A = rng.normal(size=(B, n, n))
        if B > 1:
            A[1] *= np.exp(rng.uniform(-20, 20, size=(n, 1)))      # wild row scaling
        if B > 2 and n > 2:
            A[2][np.arange(n), np.arange(n)] = 0.0                  # zero diagonal: pivoting is mandatory
        b = rng.normal(size=(B, n, nrhs))
        x = torch.zeros((B, n, nrhs), dtype=torch.float64, device='cuda')
        dA, db = _dev(A), _dev(b)
        L.check(L.lib().yhb_stage_gjp(B, n, nrhs, dA.data_ptr(), db.data_ptr(), x.data_ptr(), None))
        torch.cuda.synchronize()
        ref = np.linalg.solve(A, b)
        for k in range(B):
            assert relmax(x.cpu().numpy()[k], ref[k]) < 1e-9 * max(1.0, np.linalg.cond(A[k]) * 1e-4), (n, nrhs, k)


def test_stage_gjp_singular_reports_nan():
    """a column without an admissible pivot (all NaN) flags the system: the solution is NaN (the Newton loop's NaN
    guard takes it from there, :624-627), the other systems of the batch are unaffected"""
    import torch
    from yalrf_b200 import _lib as L
    rng = np.random.default_rng(8)
    n = 40
    A = rng.normal(size=(2, n, n))
    A[0][:, 7] = np.nan
    b = rng.normal(size=(2, n, 1))
    x = torch.zeros((2, n, 1), dtype=torch.float64, device='cuda')
    dA, db = _dev(A), _dev(b)
    L.check(L.lib().yhb_stage_gjp(2, n, 1, dA.data_ptr(), db.data_ptr(), x.data_ptr(), None))
    xs = x.cpu().numpy()
    assert np.isnan(xs[0]).all()
    assert relmax(xs[1], np.linalg.solve(A[1], b[1])) < 1e-9


def test_device_protocol_methods_golden(golden):
    """The reference's per-sample device protocol on the host mirror's device objects -- BJT.get_hb_params,
    Diode.get_mthb_params, CubicNonLinearity.get_mthb_params (SURVEY 8b "device protocol") -- evaluates on the GPU
    through the Newton loop's own model code: a few of the 512 golden evaluations of the unmodified reference each."""
    import yalrf_b200
    g = golden('bjt_model')
    y = yalrf_b200.Netlist('x')
    q = y.add_bjt('Q', 'b', 'c', 'e')
    q.options.update(dict(zip(g['opt_names'].tolist(), g['opt_vals'].tolist())))
    for i in (0, 7, 100, 311):
        vb, vc, ve = g['vin'][i]
        ob, oc, oe = g['vold'][i]
        out = np.array(q.get_hb_params(vb, vc, ve, 0., 0, ob, oc, oe))
        ref = g['out'][i]
        assert out.shape == (14,)
        assert np.array_equal(np.isnan(out), np.isnan(ref))
        ok = ~np.isnan(ref)
        assert np.max(np.abs(out[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-300)) < 1e-12
    gd = golden('diode_model')
    d = y.add_diode('D', 'b', 'e')
    d.options.update(dict(zip(gd['opt_names'].tolist(), gd['opt_vals'].tolist())))
    for i in (0, 5, 200):
        gdi, idi = d.get_mthb_params(float(gd['vd'][i]), float(gd['vdold'][i]))
        ref = gd['out'][i]
        assert abs(gdi - ref[0]) <= 1e-12 * abs(ref[0]) and abs(idi - ref[1]) <= 1e-12 * abs(ref[1])
    c = y.add_cubicnl('N', 'b', 'e', 1e-3)
    gc, ic = c.get_mthb_params(0.3, 0.2)
    assert gc == 2 * 1e-3 * 0.3 * 0.3 and ic == 1e-3 * 0.3 * 0.3 * 0.3       # CubicNonLinearity.py:36-41
