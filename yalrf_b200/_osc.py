"""Autonomous-oscillator solve: host side of yhb_oscillator_host (include/yhb.h).

The reference finds the oscillation point by Nelder-Mead on |Yosc| with Yosc = (V(node) - V(nosc2)) * 1e9 S / V(node)
(MultiToneHarmonicBalance.py:152-212), an objective quantised at 1 ulp * 1e9 S = 1.5e-7 S.  The GPU owns the
replacement (csrc/yhb_osc.cuh): a damped 2 x 2 Newton iteration on (fosc, Vosc) with the probe current from Kirchhoff's
law at the probed node and the exact Jacobian from two extra right-hand sides of the factored HB Jacobian -- ONE inner
HB solve per outer iteration and oscillator, the whole batch (start points, parameter sweeps) in one call.  This module
only builds the probe, flattens the netlist and formats the results.
"""
import ctypes as C

import numpy as np

from . import _lib as L


def kcl_yosc(topo, pb, probe_filter, node, V, Inl, S):
    """Yosc[B] (complex) = probe current / probe voltage at the fundamental, by KCL at `node` (reference numbering),
    from phasors V[B][W] and the nonlinear current spectrum Inl[B][W]: the host restatement of k_osc_update's
    objective (the tests evaluate it on phasors computed by an independent CPU solve)."""
    B = V.shape[0]
    n = node - 1
    w = 2 * np.pi * np.abs(pb.tones[:, 0])
    Vr = V.reshape(B, -1, S)
    Ir = Inl.reshape(B, -1, S)

    def ph(m):  # phasor of the fundamental at node m (reference numbering, 0 = gnd)
        return np.zeros(B, dtype=complex) if m == 0 else Vr[:, m - 1, 1] + 1j * Vr[:, m - 1, 2]

    Vn = ph(node)
    I = (Ir[:, n, 1] + 1j * Ir[:, n, 2]).astype(complex)
    f = np.abs(pb.tones[:, 0])
    for i, (kind, nodes, dev) in enumerate(topo.lin):
        if dev is probe_filter:
            continue
        val = pb.lin_val[:, i, 0]
        if kind == L.LIN_GYRATOR:  # Gyrator.py:58-73 (unit stamps)
            n1, n2, n3, n4 = nodes
            if node == n1:
                I = I + ph(n2) - ph(n3)
            if node == n2:
                I = I - ph(n1) + ph(n4)
            if node == n3:
                I = I + ph(n1) - ph(n4)
            if node == n4:
                I = I - ph(n2) + ph(n3)
            continue
        n1, n2 = nodes[0], nodes[1]
        if node not in (n1, n2):
            continue
        if kind == L.LIN_RESISTOR:
            y = 1.0 / val + 0j
        elif kind == L.LIN_CAPACITOR:
            y = 1j * w * val
        elif kind == L.LIN_INDUCTOR:
            y = 1.0 / (1j * w * val)
        else:  # another ideal harmonic filter
            y = np.where(f == val, pb.lin_val[:, i, 1], 1e-12) + 0j
        other = n2 if node == n1 else n1
        I = I + y * (Vn - ph(other))
    for i, (nodes, dev) in enumerate(topo.src):  # AC sources at the node (:665-673)
        if dev.itype != 'ac':
            continue
        iac = pb.src_val[:, i, 1] + 1j * pb.src_val[:, i, 2]
        if nodes[0] == node:
            I = I - iac
        if nodes[1] == node:
            I = I + iac
    # KCL: y_probe (Vn - V_nosc2) + I = 0, and the reference's Iosc = (Vn - V_nosc2) g  ->  Yosc = -I / Vn
    return -I / Vn


class OscillatorResult:
    pass


def add_probe(netlist, name, node, f0, V0):
    """the reference's oscillator probe (:146-150) on a copy of the netlist"""
    netlist = netlist.copy()
    probe_i = netlist.add_iac(name + '.I.Oscprobe', name + '.nosc1', 'gnd', ac=float(V0), freq=float(f0))
    netlist.add_gyrator(name + '.G.Oscprobe', name + '.nosc1', name + '.nosc2', 'gnd', 'gnd', 1)
    probe_z = netlist.add_idealharmonicfilter(name + 'IHF.Oscprobe', name + '.nosc2', node, float(f0))
    return netlist, probe_i, probe_z


def solve(hb, netlist, f0, numharmonics, V0, node, sweeps=None, tol=1e-12, maxiter=30):
    """Batched oscillator solve on the GPU; f0 and V0 are arrays of start points (broadcast against each other and the
    sweeps).  Returns an OscillatorResult: fosc[B], Vosc[B], converged[B], yosc[B] = |Yosc| reached, V, Vf, ..."""
    from . import _engine as E
    from ._flatten import ParamBlock
    f0 = np.atleast_1d(np.asarray(f0, dtype=float))
    V0 = np.atleast_1d(np.asarray(V0, dtype=float))
    B = max(len(f0), len(V0), *[len(np.atleast_1d(v)) for _, _, v in (sweeps or [])])
    f = np.ascontiguousarray(np.broadcast_to(f0, (B,)), dtype=np.float64).copy()
    A = np.ascontiguousarray(np.broadcast_to(V0, (B,)), dtype=np.float64).copy()
    netlist, probe_i, probe_z = add_probe(netlist, hb.name, node, f[0], A[0])
    hb.freq = float(f[0])
    hb.numharmonics = numharmonics
    plan, topo = hb._setup(netlist)
    pb = ParamBlock(topo, hb.freq, B)
    for dev, field, values in (sweeps or []):
        pb.set(dev, field, np.broadcast_to(np.asarray(values, dtype=float), (B,)))
    pb.set_tones(f)
    pb.set(probe_z, 'freq', f)
    pb.set(probe_i, 'ac', A)
    pb.finalize()
    oo = L.OscOptions()
    oo.probe_src = [i for i, (_, d) in enumerate(topo.src) if d is probe_i][0]
    oo.probe_filter = [i for i, (_, _, d) in enumerate(topo.lin) if d is probe_z][0]
    oo.node = netlist.get_node_idx(node)
    oo.max_outer = int(maxiter)
    oo.tol = float(tol)
    res = E.BatchResult(plan, B)
    r = L.Result()
    for name in ('V', 'Vf', 'converged', 'newton_iters', 'lu_count', 'level_iters', 'level_alpha', 'err'):
        setattr(r, name, L.ptr(getattr(res, name)))
    yosc = np.zeros((B, 2))
    conv = np.zeros(B, dtype=np.int32)
    outer = np.zeros(B, dtype=np.int32)
    newton = np.zeros(B, dtype=np.int32)
    diag = np.zeros((B, 16))
    orr = L.OscResult()
    orr.diag = L.ptr(diag)
    orr.fosc, orr.Vosc, orr.yosc = L.ptr(f), L.ptr(A), L.ptr(yosc)
    orr.converged, orr.outer_iters, orr.hb_newton_iters = L.ptr(conv), L.ptr(outer), L.ptr(newton)
    p = E.make_params(pb)
    o = E.make_options(hb.options)
    L.check(L.lib().yhb_oscillator_host(plan.handle, C.byref(p), C.byref(o), C.byref(oo), C.byref(r), C.byref(orr)))
    out = OscillatorResult()
    out.fosc, out.Vosc = f, A
    out.converged = conv.astype(bool)
    out.Yosc = yosc[:, 0] + 1j * yosc[:, 1]
    out.yosc = np.abs(out.Yosc)
    out.outer_iterations = outer
    out.jacobian = diag[:, :4].reshape(B, 2, 2)    # d(Re Yosc, Im Yosc) / d(f, V) at the returned point
    out.next_step = diag[:, 4:6]
    out.gave_up = diag[:, 7] > 0
    out.dbg = diag[:, 8:]
    out.newton_iterations = int(newton.sum())
    out.V = res.V
    out.Vf = res.Vf
    out.hb_converged = res.converged
    out.freqs = np.outer(f, np.arange(plan.K + 1))
    out.plan, out.netlist = plan, netlist
    out.probe = (probe_i, probe_z)
    return out
