"""Autonomous-oscillator solve without scipy.optimize: batched Newton on (fosc, Vosc).

The reference finds the oscillation point by Nelder-Mead on |Yosc| with Yosc = (V(node) - V(nosc2)) * 1e9 S / V(node)
(MultiToneHarmonicBalance.py:152-212), an objective quantised at 1 ulp * 1e9 S = 1.5e-7 S.  Here the probe current
is taken from Kirchhoff's law at the probed node instead -- every other current into the node, linear branches from
the converged phasors and the nonlinear part from the engine's Inl output -- which is free of the 1e9 S
cancellation, and (Re Yosc, Im Yosc) = 0 is solved by a 2x2 Newton iteration with a finite-difference Jacobian.
Each outer iteration is ONE batched GPU solve of 3 circuits per oscillator (base, f + df, V + dV), warm-started.
"""
import numpy as np

from . import _lib as L


def probe_admittance(topo, probe_filter, node, V, Inl, S, tones):
    """Yosc[B] (complex) = probe current / probe voltage at the fundamental, by KCL at `node` (reference numbering)."""
    n = node - 1
    w = 2 * np.pi * np.abs(tones[:, 0])
    Vr = V.reshape(V.shape[0], -1, S)

    def ph(m):  # phasor of the fundamental at node m (reference numbering, 0 = gnd)
        if m == 0:
            return np.zeros(V.shape[0], dtype=complex)
        return Vr[:, m - 1, 1] + 1j * Vr[:, m - 1, 2]

    Vn = ph(node)
    I = (Inl.reshape(V.shape[0], -1, S)[:, n, 1] + 1j * Inl.reshape(V.shape[0], -1, S)[:, n, 2]).astype(complex)
    return Vn, I, w


def kcl_yosc(topo, pb, probe_filter, node, V, Inl, S):
    Vn, I, w = probe_admittance(topo, probe_filter, node, V, Inl, S, pb.tones)
    B = V.shape[0]
    Vr = V.reshape(B, -1, S)

    def ph(m):
        return np.zeros(B, dtype=complex) if m == 0 else Vr[:, m - 1, 1] + 1j * Vr[:, m - 1, 2]

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


def solve(hb, netlist, f0, numharmonics, V0, node, sweeps=None, tol=1e-13, maxiter=30, rel_step=1e-5):
    """Batched oscillator solve; f0 and V0 are arrays of start points (broadcast against each other and the sweeps)."""
    from . import _engine as E
    from ._flatten import ParamBlock
    f0 = np.atleast_1d(np.asarray(f0, dtype=float))
    V0 = np.atleast_1d(np.asarray(V0, dtype=float))
    B = max(len(f0), len(V0), *[len(np.atleast_1d(v)) for _, _, v in (sweeps or [])])
    f = np.broadcast_to(f0, (B,)).copy()
    A = np.broadcast_to(V0, (B,)).copy()
    netlist = netlist.copy()
    name = hb.name
    probe_i = netlist.add_iac(name + '.I.Oscprobe', name + '.nosc1', 'gnd', ac=float(A[0]), freq=float(f[0]))
    netlist.add_gyrator(name + '.G.Oscprobe', name + '.nosc1', name + '.nosc2', 'gnd', 'gnd', 1)
    probe_z = netlist.add_idealharmonicfilter(name + 'IHF.Oscprobe', name + '.nosc2', node, float(f[0]))
    hb.freq = float(f[0])
    hb.numharmonics = numharmonics
    plan, topo = hb._setup(netlist)
    node_idx = netlist.get_node_idx(node)
    S = plan.S
    pb = ParamBlock(topo, hb.freq, 3 * B)
    for dev, field, values in (sweeps or []):
        pb.set(dev, field, np.tile(np.broadcast_to(np.asarray(values, dtype=float), (B,)), 3))
    Vwarm = None
    done = np.zeros(B, dtype=bool)
    failed = np.zeros(B, dtype=bool)
    have = np.zeros(B, dtype=bool)           # an accepted point exists
    Yc = np.full(B, np.inf, dtype=complex)   # objective at the accepted point
    sf = np.zeros(B)
    sa = np.zeros(B)                         # full Newton step from the accepted point
    lam = np.ones(B)                         # damping of the step being tried
    iters = np.zeros(B, dtype=int)
    total_newton = 0
    Vacc = None
    Vfacc = None
    for it in range(maxiter):
        ft = np.where(have, f + lam * sf, f)
        At = np.where(have, A + lam * sa, A)
        df = rel_step * ft
        dA = rel_step * At
        fall = np.concatenate([ft, ft + df, ft])
        Aall = np.concatenate([At, At, At + dA])
        pb.set_tones(fall)
        pb.set(probe_z, 'freq', fall)
        pb.set(probe_i, 'ac', Aall)
        res = E.run_host(plan, pb, hb.options, None if Vwarm is None else np.tile(Vwarm, (3, 1)), want_inl=True)
        total_newton += int(res.newton_iters.sum())
        with np.errstate(all='ignore'):
            Y = kcl_yosc(topo, pb, probe_z, node_idx, res.V, res.Inl, S)
        ok = (res.converged[:B] > 0) & (res.converged[B:2 * B] > 0) & (res.converged[2 * B:] > 0) & np.isfinite(Y[:B]) \
            & np.isfinite(Y[B:2 * B]) & np.isfinite(Y[2 * B:])
        Y0, Yf, Ya = Y[:B], Y[B:2 * B], Y[2 * B:]
        live = ~done & ~failed
        accept = live & ok & (~have | (np.abs(Y0) < np.abs(Yc)))
        reject = live & ~accept
        # rejected trials: halve the step from the accepted point (or give up)
        lam = np.where(reject, lam * 0.5, lam)
        failed |= reject & (~have | (lam < 1.0 / 256))
        if Vacc is None:
            Vacc, Vfacc = res.V[:B].copy(), res.Vf[:B].copy()
        if accept.any():
            f = np.where(accept, ft, f)
            A = np.where(accept, At, A)
            Yc = np.where(accept, Y0, Yc)
            Vacc[accept] = res.V[:B][accept]
            Vfacc[accept] = res.Vf[:B][accept]
            have |= accept
            iters[accept] += 1
            # 2x2 Newton step on (Re Y, Im Y) with the finite-difference Jacobian
            with np.errstate(all='ignore'):
                a11 = (Yf.real - Y0.real) / df
                a12 = (Ya.real - Y0.real) / dA
                a21 = (Yf.imag - Y0.imag) / df
                a22 = (Ya.imag - Y0.imag) / dA
                det = a11 * a22 - a12 * a21
                nf = -(a22 * Y0.real - a12 * Y0.imag) / det
                na = -(-a21 * Y0.real + a11 * Y0.imag) / det
            nf = np.clip(np.nan_to_num(nf), -0.02 * ft, 0.02 * ft)      # trust region: 2 % in frequency,
            na = np.clip(np.nan_to_num(na), -0.3 * At, 0.3 * At)        # 30 % in amplitude
            sf = np.where(accept, nf, sf)
            sa = np.where(accept, na, sa)
            lam = np.where(accept, 1.0, lam)
            small = accept & (np.abs(sf) <= 4 * np.spacing(f)) & (np.abs(sa) <= 4 * np.spacing(A))
            done |= accept & ((np.abs(Yc) < tol) | small)
        Vwarm = Vacc
        if not (~done & ~failed).any():
            break
    ynorm = np.abs(Yc)
    out = OscillatorResult()
    out.fosc, out.Vosc = f, A
    out.converged = done & ~failed
    out.yosc = ynorm
    out.outer_iterations = iters
    out.newton_iterations = total_newton
    out.V = Vacc
    out.Vf = Vfacc
    out.freqs = np.outer(f, np.arange(plan.K + 1))
    out.plan, out.netlist = plan, netlist
    return out
