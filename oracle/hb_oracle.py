"""CPU oracle: numpy restatement of YalRF's MultiToneHarmonicBalance path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``yalrf_b200/`` may import this file;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / reported baseline.

Parity status: PINNED.  The reference is pure Python, so it is imported in the
build container by ``tests/golden/make_golden.py`` and its outputs (converged
phasors, per-level Newton iteration counts, J/F/dV of individual iterations, the
14-tuple BJT and 2-tuple diode model outputs, oscillator frequency/amplitude) are
committed under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this
restatement against them.  The reference's own tests hold no HB assertions
(SURVEY.md section 4), so those reference runs are the pins.

Every function cites the reference file:line it follows (paths relative to the
reference checkout, ``yalrf/...``).  The arithmetic is kept operation-for-operation
where it matters for the 1e-9 parity bar (same iteration sequence, limiter,
accumulating dQdV, >=3 iteration rule, stop test); per-sample Python loops are
vectorised over the S time samples, and the dense ``Omega @ X`` products are
applied as the row-pair swap/scale they are (identical results: every other
term of those dot products is an exact zero).
"""
from __future__ import annotations

import copy
import math

import numpy as np
import scipy.linalg
from scipy import optimize

# scipy.constants.k / .e (CODATA exact values) -- Diode.py:2, BJT.py:1
K_BOLTZ = 1.380649e-23
Q_ELEM = 1.602176634e-19

INF = math.inf

# --------------------------------------------------------------------------- #
# device option tables (Diode.py:6-36, BJT.py:5-52)
# --------------------------------------------------------------------------- #
DIODE_OPTIONS = {
    'Temp': 300.0, 'Is': 1e-15, 'N': 1.0, 'Isr': 0.0, 'Nr': 2.0, 'Rs': 0.0, 'Cj0': 0.0,
    'M': 0.5, 'Vj': 0.7, 'Fc': 0.5, 'Cp': 0.0, 'Tt': 0.0, 'Bv': INF, 'Ibv': 0.001,
    'Ikf': INF, 'Kf': 0.0, 'Af': 1.0, 'Ffe': 1.0, 'Xti': 3.0, 'Eg': 1.11, 'Ega': 7.02e-4,
    'Egb': 1108.0, 'Tbv': 0.0, 'Trs': 0.0, 'Tt1': 0.0, 'Tt2': 0.0, 'Tm1': 0.0, 'Tm2': 0.0,
    'Tnom': 300.0, 'Area': 1.0,
}

BJT_OPTIONS = {
    'Temp': 300.0, 'Is': 1e-15, 'Nf': 1.0, 'Nr': 1.0, 'Ikf': 1e9, 'Ikr': 1e9, 'Vaf': 1e12,
    'Var': 1e12, 'Ise': 0.0, 'Ne': 1.5, 'Isc': 0.0, 'Nc': 2.0, 'Bf': 100.0, 'Br': 1.0,
    'Rbm': 0.0, 'Irb': 1e9, 'Rc': 0.0, 'Re': 0.0, 'Rb': 0.0, 'Cje': 0.0, 'Vje': 0.75,
    'Mje': 0.33, 'Cjc': 0.0, 'Vjc': 0.75, 'Mjc': 0.33, 'Xcjc': 1.0, 'Cjs': 0.0, 'Vjs': 0.75,
    'Mjs': 0.0, 'Fc': 0.5, 'Tf': 0.0, 'Xtf': 0.0, 'Vtf': 1e12, 'Itf': 0.0, 'Ptf': 0.0,
    'Tr': 0.0, 'Kf': 0.0, 'Af': 1.0, 'Ffe': 1.0, 'Kb': 0.0, 'Ab': 1.0, 'Fb': 1.0, 'Xti': 3.0,
    'Xtb': 0.0, 'Eg': 1.11, 'Tnom': 300.0, 'Area': 1.0,
}


class Dev:
    """Plain record of one netlist device (kind + node indices + parameters)."""

    def __init__(self, kind, name, nodes, **kw):
        self.kind = kind
        self.name = name
        self.nodes = list(nodes)
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def n1(self):
        return self.nodes[0]

    @property
    def n2(self):
        return self.nodes[1]

    @property
    def n3(self):
        return self.nodes[2]

    @property
    def n4(self):
        return self.nodes[3]


NONLINEAR = ('diode', 'bjt', 'cubicnl')


class FlatNetlist:
    """Recorder with the reference's ``Netlist.add_*`` signatures (Netlist.py:31-694).

    The same circuit-builder functions (tests/circuits.py) run against this class,
    the reference's ``yalrf.Netlist`` and the product's ``yalrf_b200.Netlist``.
    """

    def __init__(self, name=''):
        self.name = name
        self.devices = []
        self.node_name_to_idx = {'gnd': 0}
        self.node_idx_to_name = ['gnd']

    # Netlist.py:703-725
    def add_node(self, n):
        if n not in self.node_name_to_idx:
            self.node_name_to_idx[n] = len(self.node_name_to_idx)
            self.node_idx_to_name.append(n)
        return self.node_name_to_idx[n]

    def get_node_idx(self, name):
        return self.node_name_to_idx[name]

    def get_num_nodes(self):
        return len(self.node_idx_to_name)

    def _add(self, dev):
        self.devices.append(dev)
        return dev

    def add_resistor(self, name, n1, n2, value):
        return self._add(Dev('resistor', name, [self.add_node(n1), self.add_node(n2)], R=float(value)))

    def add_capacitor(self, name, n1, n2, value):
        return self._add(Dev('capacitor', name, [self.add_node(n1), self.add_node(n2)], C=float(value)))

    def add_inductor(self, name, n1, n2, value):
        return self._add(Dev('inductor', name, [self.add_node(n1), self.add_node(n2)], L=float(value)))

    # CurrentSource.py:5-12 (phase stored in radians)
    def add_idc(self, name, n1, n2, dc):
        return self._add(Dev('isource', name, [self.add_node(n1), self.add_node(n2)], itype='dc',
                             dc=float(dc), ac=0.0, phase=np.radians(0.0), freq=0.0))

    def add_iac(self, name, n1, n2, ac, phase=0, freq=0):
        return self._add(Dev('isource', name, [self.add_node(n1), self.add_node(n2)], itype='ac',
                             dc=0.0, ac=float(ac), phase=np.radians(float(phase)), freq=float(freq)))

    def add_gyrator(self, name, n1, n2, n3, n4, G):
        nodes = [self.add_node(n1), self.add_node(n2), self.add_node(n3), self.add_node(n4)]
        return self._add(Dev('gyrator', name, nodes, G=float(G)))

    # IdealHarmonicFilter.py:5-12
    def add_idealharmonicfilter(self, name, n1, n2, freq):
        return self._add(Dev('ihf', name, [self.add_node(n1), self.add_node(n2)], freq=float(freq), g=1e9))

    def add_diode(self, name, n1, n2):
        return self._add(Dev('diode', name, [self.add_node(n1), self.add_node(n2)],
                             options=DIODE_OPTIONS.copy()))

    def add_bjt(self, name, n1, n2, n3, n4='gnd', ispnp=False):
        nodes = [self.add_node(n1), self.add_node(n2), self.add_node(n3), self.add_node(n4)]
        return self._add(Dev('bjt', name, nodes, options=BJT_OPTIONS.copy(), type=-1 if ispnp else 1))

    def add_cubicnl(self, name, n1, n2, alpha):
        return self._add(Dev('cubicnl', name, [self.add_node(n1), self.add_node(n2)], alpha=alpha))

    def copy(self):
        # Netlist.py:889-897 (shallow: same device objects)
        c = FlatNetlist(self.name)
        c.devices = self.devices.copy()
        c.node_name_to_idx = self.node_name_to_idx.copy()
        c.node_idx_to_name = self.node_idx_to_name.copy()
        return c

    def get_linear_devices(self):
        return [d for d in self.devices if d.kind not in NONLINEAR]

    def get_nonlinear_devices(self):
        return [d for d in self.devices if d.kind in NONLINEAR]


# --------------------------------------------------------------------------- #
# device models
# --------------------------------------------------------------------------- #
def exp_lim(x):
    """Diode.py:434-435 / BJT.py:588-589 -- exponential linearised above x = 200."""
    x = np.asarray(x, dtype=float)
    with np.errstate(over='ignore', invalid='ignore'):
        return np.where(x < 200., np.exp(np.minimum(x, 200.)), np.exp(200.) + np.exp(200.) * (x - 200.))


def diode_mthb_params(opt, Vd, Vdold, with_vlim=False):
    """Diode.get_mthb_params, Diode.py:395-427 (raw options: no Area scaling, no Rs, no charge)."""
    Is, N, Isr, Nr = opt['Is'], opt['N'], opt['Isr'], opt['Nr']
    Ikf, Bv, Ibv, T = opt['Ikf'], opt['Bv'], opt['Ibv'], opt['Temp']
    Vt = K_BOLTZ * T / Q_ELEM

    Vd = Vdold + 10. * N * Vt * np.tanh((Vd - Vdold) / (10. * N * Vt))

    Idf = Is * (exp_lim(Vd / (N * Vt)) - 1.)
    Idr = Isr * (exp_lim(Vd / (Nr * Vt)) - 1.)
    gdf = Is / (N * Vt) * exp_lim(Vd / (N * Vt))
    gdr = Isr / (Nr * Vt) * exp_lim(Vd / (Nr * Vt))

    if math.isfinite(Ikf):
        with np.errstate(invalid='ignore', divide='ignore'):
            Idf = Idf * np.sqrt(Ikf / (Ikf + Idf))
            gdf = gdf * (1. - 0.5 * Idf / (Ikf + Idf)) * np.sqrt(Ikf / (Ikf + Idf))

    Ibr = 0.
    gbr = 0.
    if math.isfinite(Bv):
        Ibr = Ibv * exp_lim(- Bv / Vt) * (1. - exp_lim(- Vd / Vt))
        gbr = Ibv / Vt * exp_lim(- (Bv + Vd) / Vt)

    Id = Idf + Idr + Ibr
    gd = gdf + gdr + gbr
    if with_vlim:
        return gd, Id, Vd
    return gd, Id


def diode_charge_params(opt, Vd, gd, Id):
    """Extension (diode_charge=True; the reference's MTHB has no diode charge): Diode.calc_oppoint's charge model,
    Diode.py:207-235 -- Cd = Cp + Tt gd + Cj, Qd = Cp Vd + Tt Id + Qj, Cj0 scaled by Area (Diode.py:95); Vd is the
    LIMITED junction voltage get_mthb_params worked with."""
    Cj0 = opt['Cj0'] * opt['Area']
    Cj = pn_capacitance(Vd, Cj0, opt['Vj'], opt['M'], opt['Fc'])
    Qj = pn_charge(Vd, Cj0, opt['Vj'], opt['M'], opt['Fc'])
    if Cj0 == 0.:
        Cj, Qj = np.zeros_like(np.asarray(Vd, dtype=float)), np.zeros_like(np.asarray(Vd, dtype=float))
    Cd = opt['Cp'] + opt['Tt'] * gd + Cj
    Qd = opt['Cp'] * Vd + opt['Tt'] * Id + Qj
    return Cd, Qd


def cubic_mthb_params(alpha, V):
    """CubicNonLinearity.get_mthb_params, CubicNonLinearity.py:36-41 (g = 2 alpha V^2 as written)."""
    g = 2 * alpha * V * V
    I = alpha * V * V * V
    return g, I


def pn_capacitance(Vpn, Cj, Vj, Mj, Fc):
    """BJT.py:591-597"""
    with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
        a = Cj * np.power((1. - (Vpn / Vj)), -Mj)
        b = Cj / np.power((1. - Fc), Mj) * (1. + Mj * (Vpn - Fc * Vj) / (Vj * (1. - Fc)))
    return np.where(Vpn <= Fc * Vj, a, b)


def pn_charge(Vpn, Cj, Vj, Mj, Fc):
    """BJT.py:599-608"""
    with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
        a = Cj * Vj / (1. - Mj) * (1. - np.power((1. - Vpn / Vj), (1. - Mj)))
        X = (1. - np.power((1. - Fc), (1. - Mj))) / (1. - Mj) + \
            (1. - Fc * (1. + Mj)) / np.power((1. - Fc), (1. + Mj)) * (Vpn / Vj - Fc) + \
            Mj / (2. * np.power((1. - Fc), (1. + Mj))) * (np.square(Vpn / Vj) - np.square(Fc))
        b = Cj * Vj * X
    return np.where(Vpn <= Fc * Vj, a, b)


def bjt_adjusted_options(opt):
    """BJT.init, BJT.py:114-132: Area scaling."""
    A = opt['Area']
    adj = {}
    for k in ('Is', 'Ise', 'Isc', 'Ikf', 'Ikr', 'Irb', 'Itf', 'Cje', 'Cjs', 'Cjc'):
        adj[k] = opt[k] * A
    for k in ('Rb', 'Rbm', 'Rc', 'Re'):
        adj[k] = opt[k] / A
    return adj


def diode_adjusted_options(opt):
    """Diode.init, Diode.py:89-96."""
    A = opt['Area']
    adj = {}
    for k in ('Is', 'Isr', 'Ikf', 'Ibv', 'Cj0'):
        adj[k] = opt[k] * A
    adj['Rs'] = opt['Rs'] / A
    return adj


def bjt_hb_params(opt, Vb, Vc, Ve, Vbold, Vcold, Veold, pnp_sign=1.0):
    """BJT.get_hb_params, BJT.py:462-582, vectorised over the time samples.

    ``pnp_sign`` = -1 is the documented *extension* (negate terminal voltages in, negate
    currents and charges out); the reference ignores ``self.type`` here (BJT.py:495-503),
    i.e. pnp_sign = +1 always.
    """
    adj = bjt_adjusted_options(opt)
    Vt = K_BOLTZ * opt['Temp'] / Q_ELEM
    Is, Nf, Nr = adj['Is'], opt['Nf'], opt['Nr']
    Ikf, Ikr, Vaf, Var = adj['Ikf'], adj['Ikr'], opt['Vaf'], opt['Var']
    Ise, Ne, Isc, Nc = adj['Ise'], opt['Ne'], adj['Isc'], opt['Nc']
    Bf, Br = opt['Bf'], opt['Br']
    Cje, Vje, Mje = adj['Cje'], opt['Vje'], opt['Mje']
    Cjc, Vjc, Mjc = adj['Cjc'], opt['Vjc'], opt['Mjc']
    Cjs, Vjs, Mjs = adj['Cjs'], opt['Vjs'], opt['Mjs']
    Fc, Xcjc, Tf, Xtf, Vtf, Itf, Tr = opt['Fc'], opt['Xcjc'], opt['Tf'], opt['Xtf'], opt['Vtf'], adj['Itf'], opt['Tr']

    Vs = 0.  # forced by the caller, MultiToneHarmonicBalance.py:504
    if pnp_sign != 1.0:
        Vb, Vc, Ve, Vbold, Vcold, Veold = (pnp_sign * x for x in (Vb, Vc, Ve, Vbold, Vcold, Veold))

    Vbe = Vb - Ve
    Vbc = Vb - Vc
    Vsc = Vs - Vc
    Vbeold = Vbold - Veold
    Vbcold = Vbold - Vcold

    Vbe = Vbeold + 10. * Nf * Vt * np.tanh((Vbe - Vbeold) / (10. * Nf * Vt))
    Vbc = Vbcold + 10. * Nr * Vt * np.tanh((Vbc - Vbcold) / (10. * Nr * Vt))

    gmin = 1e-12
    cmin = 1e-18

    with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
        If = Is * (exp_lim(Vbe / (Nf * Vt)) - 1.)
        Ibei = If / Bf
        Iben = Ise * (exp_lim(Vbe / (Ne * Vt)) - 1.)
        Ibe = Ibei + Iben + gmin * Vbe

        Ir = Is * (exp_lim(Vbc / (Nr * Vt)) - 1.)
        Ibci = Ir / Br
        Ibcn = Isc * (exp_lim(Vbc / (Nc * Vt)) - 1.)
        Ibc = Ibci + Ibcn + gmin * Vbc

        Q1 = 1. / (1. - (Vbc / Vaf) - (Vbe / Var))
        Q2 = (If / Ikf) + (Ir / Ikr)
        Qb = (Q1 / 2.) * (1. + np.sqrt(1. + 4. * Q2))

        It = (If - Ir) / Qb

        gbei = Is / (Nf * Vt * Bf) * exp_lim(Vbe / (Nf * Vt))
        gben = Ise / (Ne * Vt) * exp_lim(Vbe / (Ne * Vt))
        gpi = gbei + gben + gmin

        gbci = Is / (Nr * Vt * Br) * exp_lim(Vbc / (Nr * Vt))
        gbcn = Isc / (Nc * Vt) * exp_lim(Vbc / (Nc * Vt))
        gmu = gbci + gbcn + gmin

        gif = gbei * Bf
        gir = gbci * Br

        dQb_dVbe = Q1 * ((Qb / Var) + (gif / (Ikf * np.sqrt(1. + 4. * Q2))))
        dQb_dVbc = Q1 * ((Qb / Vaf) + (gir / (Ikr * np.sqrt(1. + 4. * Q2))))

        gmf = (1. / Qb) * (+ gif - It * dQb_dVbe)
        gmr = (1. / Qb) * (- gir - It * dQb_dVbc)

        Cbedep = pn_capacitance(Vbe, Cje, Vje, Mje, Fc)
        Qbe = pn_charge(Vbe, Cje, Vje, Mje, Fc)
        Cbcdep = pn_capacitance(Vbc, Cjc, Vjc, Mjc, Fc)
        Qbc = pn_charge(Vbc, Cjc, Vjc, Mjc, Fc)

        # BJT.py:549-554
        Csc_a = Cjs * np.power((1. - (Vsc / Vjs)), -Mjs)
        Qsc_a = Cjs * Vjs / (1 - Mjs) * (1 - np.power(1. - Vsc / Vjs, 1. - Mjs))
        Csc_b = Cjs * (1. + Mjs * Vsc / Vjs)
        Qsc_b = Cjs * Vsc / (1 + (Mjs * Vsc) / (2. * Vjs))
        Cscdep = np.where(Vsc <= 0, Csc_a, Csc_b)
        Qsc = np.where(Vsc <= 0, Qsc_a, Qsc_b)

        e_tf = exp_lim(Vbc / (1.44 * Vtf))
        Tff = Tf * (1. + Xtf * np.square(If / (If + Itf)) * e_tf)
        dTff_dVbe = (Tf * Xtf * 2 * gif * If * Itf / (If + Itf) ** 3) * e_tf
        dTff_dVbc = (Tf * Xtf / (1.44 * Vtf)) * np.square(If / (If + Itf)) * e_tf

        Cbcidep = Xcjc * Cbcdep
        Cbcdiff = Tr * gir
        Cbediff = (1. / Qb) * (If * dTff_dVbe + Tff * (gif - (If / Qb) * dQb_dVbe))
        Cbebc = (If / Qb) * (dTff_dVbc - (Tff / Qb) * dQb_dVbc) + cmin

        Ib = Ibe + Ibc
        Ic = It - Ibc
        Ie = Ib + Ic

        Cbc = Cbcidep + Cbcdiff + cmin
        Qbc = Qbc + (Ir * Tr + cmin * Vbc)
        Cbe = Cbedep + Cbediff + cmin
        Qbe = Qbe + (If * Tff / Qb + cmin * Vbe)
        Csc = Cscdep + cmin
        Qsc = Qsc + cmin * Vsc

    if pnp_sign != 1.0:
        Ib, Ic, Ie, Qbe, Qbc, Qsc = (pnp_sign * x for x in (Ib, Ic, Ie, Qbe, Qbc, Qsc))
    return Ib, Ic, Ie, Qbe, Qbc, Qsc, gmu, gpi, gmf, gmr, Cbc, Cbe, Cbebc, Csc


# --------------------------------------------------------------------------- #
# DC operating point (initial guess of a cold HB start)
# --------------------------------------------------------------------------- #
DC_OPTIONS = {
    'max_iterations': 150, 'gmin': 1e-12, 'reltol': 1e-3, 'vabstol': 1e-6, 'iabstol': 1e-12,
    'use_gmin_stepping': True, 'use_source_stepping': True,
    'gmin_max': 0.01, 'gmin_min': 1e-12, 'gmin_rate': 0.05, 'gmin_maxiter': 50,
    'alpha_min': 1e-3, 'alpha_delta': 1e-3, 'source_maxiter': 50,
}


class DCOracle:
    """Restates DC.run and the DC stamps of the HB-reachable devices (DC.py:54-308)."""

    def __init__(self, netlist):
        self.options = DC_OPTIONS.copy()
        self.netlist = netlist
        self.n = netlist.get_num_nodes()
        self.devs = netlist.devices
        self.lin = netlist.get_linear_devices()
        self.nonlin = netlist.get_nonlinear_devices()
        # Netlist.get_mna_extra_rows_dict, Netlist.py:843-865 (inductor = 1 branch row)
        self.iidx = {}
        k = 0
        for d in self.devs:
            if d.kind == 'inductor':
                self.iidx[id(d)] = self.n + k
                k += 1
        self.m = k
        self.state = {}
        self.total_newton = 0

    # -- per-device state: Diode.init (Diode.py:77-96), BJT.init (BJT.py:105-132)
    def init_devices(self):
        for d in self.nonlin:
            if d.kind == 'diode':
                self.state[id(d)] = {'Vdold': 0., 'It': 0., 'gt': 0., 'adj': diode_adjusted_options(d.options)}
            elif d.kind == 'bjt':
                self.state[id(d)] = {'Vbeold': 0., 'Vbcold': 0., 'adj': bjt_adjusted_options(d.options)}

    # -- linear stamps
    def stamp_linear(self, A, z):
        for d in self.lin:
            k = d.kind
            if k == 'resistor':  # Resistor.py:30-35
                g = 1.0 / d.R
                A[d.n1][d.n1] += g
                A[d.n2][d.n2] += g
                A[d.n1][d.n2] -= g
                A[d.n2][d.n1] -= g
            elif k == 'inductor':  # Inductor.py:21-26
                i = self.iidx[id(d)]
                A[d.n1][i] = +1.0
                A[d.n2][i] = -1.0
                A[i][d.n1] = +1.0
                A[i][d.n2] = -1.0
                z[i] = 0.0
            elif k == 'isource':  # CurrentSource.py:23-26
                if d.itype != 'ac':
                    z[d.n1] = z[d.n1] + d.dc
                    z[d.n2] = z[d.n2] - d.dc
            elif k == 'gyrator':  # Gyrator.py:22-30
                A[d.n1, d.n2] += d.G
                A[d.n1, d.n3] -= d.G
                A[d.n2, d.n1] -= d.G
                A[d.n2, d.n4] += d.G
                A[d.n3, d.n1] += d.G
                A[d.n3, d.n4] -= d.G
                A[d.n4, d.n2] -= d.G
                A[d.n4, d.n3] += d.G
            # capacitor (Capacitor.py:30-31) and ihf (IdealHarmonicFilter.py:23-24): no DC stamp

    # -- Diode.limit_diode_voltage, Diode.py:320-332
    def _limit_diode(self, d, st, Vd, Vt):
        N = d.options['N']
        Vd = st['Vdold'] + 10. * N * Vt * np.tanh((Vd - st['Vdold']) / (10. * N * Vt))
        st['Vdold'] = Vd
        return Vd

    # -- BJT.limit_bjt_voltages, BJT.py:438-457
    def _limit_bjt(self, d, st, Vbe, Vbc, Vt):
        Nf, Nr = d.options['Nf'], d.options['Nr']
        Vbe = st['Vbeold'] + 10. * Nf * Vt * np.tanh((Vbe - st['Vbeold']) / (10. * Nf * Vt))
        Vbc = st['Vbcold'] + 10. * Nr * Vt * np.tanh((Vbc - st['Vbcold']) / (10. * Nr * Vt))
        st['Vbeold'] = Vbe
        st['Vbcold'] = Vbc
        return Vbe, Vbc

    @staticmethod
    def _v(x, n):
        return x[n - 1, 0] if n > 0 else 0.

    def stamp_nonlinear(self, A, z, x):
        for d in self.nonlin:
            if d.kind == 'diode':
                self._stamp_diode(d, A, z, x)
            elif d.kind == 'bjt':
                self._stamp_bjt(d, A, z, x)
            else:
                self._stamp_cubic(d, A, z, x)

    def _stamp_diode(self, d, A, z, x):
        """Diode.calc_dc + add_dc_stamps, Diode.py:98-116, 238-294."""
        st = self.state[id(d)]
        o, adj = d.options, st['adj']
        Is, N, Isr, Nr = adj['Is'], o['N'], adj['Isr'], o['Nr']
        Ikf, Bv, Ibv, Rs = adj['Ikf'], o['Bv'], adj['Ibv'], adj['Rs']
        Vt = K_BOLTZ * o['Temp'] / Q_ELEM
        Vtot = self._v(x, d.n1) - self._v(x, d.n2)
        Id = st['It'] + st['gt'] * Vtot
        Vd = Vtot - Id * Rs
        Vd = self._limit_diode(d, st, Vd, Vt)
        Idf = Is * (exp_lim(Vd / (N * Vt)) - 1.)
        gdf = Is / (N * Vt) * exp_lim(Vd / (N * Vt))
        Idr = Isr * (exp_lim(Vd / (Nr * Vt)) - 1.)
        gdr = Isr / (Nr * Vt) * exp_lim(Vd / (Nr * Vt))
        if math.isfinite(Ikf):
            Idf = Idf * np.sqrt(Ikf / (Ikf + Idf))
            gdf = gdf * (1. - 0.5 * Idf / (Ikf + Idf)) * np.sqrt(Ikf / (Ikf + Idf))
        Ibr = 0.
        gbr = 0.
        if math.isfinite(Bv):
            Ibr = Ibv * exp_lim(- Bv / Vt) * (1. - exp_lim(- Vd / Vt))
            gbr = Ibv / Vt * exp_lim(- (Bv + Vd) / Vt)
        Id = Idf + Idr + Ibr + (Vd * 1e-12)
        gd = gdf + gdr + gbr + (1e-12)
        st['It'] = (Id - gd * Vd) / (1. + gd * Rs)
        st['gt'] = gd / (1. + gd * Rs)
        gt, It = st['gt'], st['It']
        A[d.n1][d.n1] += gt
        A[d.n2][d.n2] += gt
        A[d.n1][d.n2] -= gt
        A[d.n2][d.n1] -= gt
        z[d.n1] = z[d.n1] - It
        z[d.n2] = z[d.n2] + It

    def _stamp_bjt(self, d, A, z, x):
        """BJT.calc_dc + add_dc_stamps, BJT.py:134-167, 319-416 (honours self.type)."""
        st = self.state[id(d)]
        o, adj = d.options, st['adj']
        B, C, E = d.n1, d.n2, d.n3
        Vt = K_BOLTZ * o['Temp'] / Q_ELEM
        Is, Nf, Nr = adj['Is'], o['Nf'], o['Nr']
        Ikf, Ikr, Vaf, Var = adj['Ikf'], adj['Ikr'], o['Vaf'], o['Var']
        Ise, Ne, Isc, Nc, Bf, Br = adj['Ise'], o['Ne'], adj['Isc'], o['Nc'], o['Bf'], o['Br']
        Vb, Vc, Ve = self._v(x, B), self._v(x, C), self._v(x, E)
        Vbe = (Vb - Ve) * d.type
        Vbc = (Vb - Vc) * d.type
        gmin = 1e-12
        Vbe, Vbc = self._limit_bjt(d, st, Vbe, Vbc, Vt)

        If = Is * (exp_lim(Vbe / (Nf * Vt)) - 1.)
        Ibei = If / Bf
        Iben = Ise * (exp_lim(Vbe / (Ne * Vt)) - 1.)
        Ibe = Ibei + Iben + gmin * Vbe
        gbei = Is / (Nf * Vt * Bf) * exp_lim(Vbe / (Nf * Vt))
        gben = Ise / (Ne * Vt) * exp_lim(Vbe / (Ne * Vt))
        gpi = gbei + gben + gmin
        Ir = Is * (exp_lim(Vbc / (Nr * Vt)) - 1.)
        Ibci = Ir / Br
        Ibcn = Isc * (exp_lim(Vbc / (Nc * Vt)) - 1.)
        Ibc = Ibci + Ibcn + gmin * Vbc
        gbci = Is / (Nr * Vt * Br) * exp_lim(Vbc / (Nr * Vt))
        gbcn = Isc / (Nc * Vt) * exp_lim(Vbc / (Nc * Vt))
        gmu = gbci + gbcn + gmin
        Q1 = 1. / (1. - (Vbc / Vaf) - (Vbe / Var))
        Q2 = (If / Ikf) + (Ir / Ikr)
        Qb = (Q1 / 2.) * (1. + np.sqrt(1. + 4. * Q2))
        It = (If - Ir) / Qb
        gif = gbei * Bf
        gir = gbci * Br
        dQb_dVbe = Q1 * ((Qb / Var) + (gif / (Ikf * np.sqrt(1. + 4. * Q2))))
        dQb_dVbc = Q1 * ((Qb / Vaf) + (gir / (Ikr * np.sqrt(1. + 4. * Q2))))
        gmf = (1. / Qb) * (+ gif - It * dQb_dVbe)
        gmr = (1. / Qb) * (- gir - It * dQb_dVbc)

        Ibeeq = Ibe - gpi * Vbe
        Ibceq = Ibc - gmu * Vbc
        Iceeq = It - gmf * Vbe - gmr * Vbc
        A[B][B] += gmu + gpi
        A[B][C] -= gmu
        A[B][E] -= gpi
        A[C][B] += - gmu + gmf + gmr
        A[C][C] += gmu - gmr
        A[C][E] -= gmf
        A[E][B] += - gpi - gmf - gmr
        A[E][C] += gmr
        A[E][E] += gpi + gmf
        z[B] = z[B] + (- Ibeeq - Ibceq) * d.type
        z[C] = z[C] + (+ Ibceq - Iceeq) * d.type
        z[E] = z[E] + (+ Ibeeq + Iceeq) * d.type

    def _stamp_cubic(self, d, A, z, x):
        """CubicNonLinearity.add_dc_stamps, CubicNonLinearity.py:20-34."""
        V = self._v(x, d.n1) - self._v(x, d.n2)
        g = 3 * d.alpha * V * V
        I = d.alpha * V * V * V
        Ieq = I - g * V
        A[d.n1][d.n1] += g
        A[d.n2][d.n2] += g
        A[d.n1][d.n2] -= g
        A[d.n2][d.n1] -= g
        z[d.n1] = z[d.n1] - Ieq
        z[d.n2] = z[d.n2] + Ieq

    # -- Diode.check_vlimit (Diode.py:296-313), BJT.check_vlimit (BJT.py:418-436): mutate the limiter state
    def check_vlimit(self, x, vabstol):
        ok = True
        for d in self.devs:
            if d.kind == 'diode':
                st = self.state[id(d)]
                Vt = K_BOLTZ * d.options['Temp'] / Q_ELEM
                Vtot = self._v(x, d.n1) - self._v(x, d.n2)
                Id = st['It'] + st['gt'] * Vtot
                Vd = Vtot - Id * st['adj']['Rs']
                Vdlim = self._limit_diode(d, st, Vd, Vt)
                if Vd - Vdlim > vabstol:
                    ok = False
            elif d.kind == 'bjt':
                st = self.state[id(d)]
                Vt = K_BOLTZ * d.options['Temp'] / Q_ELEM
                Vb, Vc, Ve = self._v(x, d.n1), self._v(x, d.n2), self._v(x, d.n3)
                Vbe = (Vb - Ve) * d.type
                Vbc = (Vb - Vc) * d.type
                Vbelim, Vbclim = self._limit_bjt(d, st, Vbe, Vbc, Vt)
                if (Vbe - Vbelim > vabstol) or (Vbc - Vbclim > vabstol):
                    ok = False
        return ok

    @staticmethod
    def solve_linear(A, z):
        """Solver.py:6-14"""
        with np.errstate(all='ignore'):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                lu, piv = scipy.linalg.lu_factor(A, check_finite=False)
                x = scipy.linalg.lu_solve((lu, piv), z, check_finite=False)
        return x, not np.isnan(np.sum(x))

    def solve_dc_nonlinear(self, A, z, x0):
        """DC.py:125-191"""
        o = self.options
        reltol, vabstol, iabstol, maxiter = o['reltol'], o['vabstol'], o['iabstol'], o['max_iterations']
        xk = x0.copy()
        Anl = np.zeros(A.shape)
        znl = np.zeros(z.shape)
        converged = False
        k = 0
        while (not converged) and (k < maxiter):
            Anl[:, :] = 0.0
            znl[:] = 0.0
            self.stamp_nonlinear(Anl, znl, xk)
            An = A[1:, 1:] + Anl[1:, 1:]
            zn = z[1:] + znl[1:]
            self.x, issolved = self.solve_linear(An, zn)
            self.total_newton += 1
            if not issolved:
                return False
            dx = self.x - xk
            vconverged = bool(np.all(np.abs(dx[:self.n - 1, 0]) <= reltol * np.abs(xk[:self.n - 1, 0]) + vabstol))
            iconverged = bool(np.all(np.abs(dx[self.n - 1:, 0]) <= reltol * np.abs(xk[self.n - 1:, 0]) + iabstol))
            vlimconverged = self.check_vlimit(self.x, vabstol)
            if vconverged and iconverged and vlimconverged:
                converged = True
            else:
                xk = self.x
                k = k + 1
        return converged

    def solve_source_stepping(self, A, z, x0):
        """DC.py:236-279"""
        o = self.options
        alpha_min, alpha_delta, maxiter = o['alpha_min'], o['alpha_delta'], o['source_maxiter']
        alpha = alpha_min
        xprev = x0.copy()
        self.x = x0.copy()
        converged = False
        k = 0
        while (not converged) and (k < maxiter):
            zn = alpha * z
            issolved = self.solve_dc_nonlinear(A, zn, self.x)
            if issolved:
                xprev = self.x
                if alpha >= 1.0:
                    converged = True
                else:
                    alpha = alpha + alpha_delta
                    alpha_delta = 2 * alpha_delta
                    if alpha >= 1.0:
                        alpha = 1.0
            else:
                self.x = xprev
                alpha = alpha - alpha_delta / 2
                alpha_delta = alpha_delta / 2
                if alpha < alpha_min:
                    converged = False
                    break
            k = k + 1
        return converged

    def solve_gmin_stepping(self, A, z, x0):
        """DC.py:193-234 (note: k is never incremented in the reference loop)."""
        o = self.options
        gmin_max, gmin_min, gmin_rate = o['gmin_max'], o['gmin_min'], o['gmin_rate']
        gmin = gmin_max
        xprev = x0.copy()
        self.x = x0.copy()
        converged = False
        guard = 0
        while not converged and guard < 10000:
            guard += 1
            An = A.copy()
            for i in range(self.n):
                An[i, i] = An[i, i] + gmin
            issolved = self.solve_dc_nonlinear(An, z, self.x)
            if issolved:
                xprev = self.x
                if gmin < gmin_min:
                    converged = True
                else:
                    gmin = gmin * gmin_rate
                    gmin_rate = gmin_rate / 4
            else:
                self.x = xprev
                gmin = gmin / gmin_rate
                gmin_rate = 2 * gmin_rate
                if gmin > gmin_max:
                    converged = False
                    break
        return converged

    def run(self):
        """DC.run, DC.py:54-123"""
        n, m = self.n, self.m
        A = np.zeros((n + m, n + m))
        z = np.zeros((n + m, 1))
        x0 = np.zeros((n + m - 1, 1))
        self.init_devices()
        self.stamp_linear(A, z)
        for i in range(n):
            A[i, i] = A[i, i] + self.options['gmin']
        if not self.nonlin:
            self.x, issolved = self.solve_linear(A[1:, 1:], z[1:])
            if issolved:
                return self.x
        if self.solve_dc_nonlinear(A, z, x0):
            return self.x
        if self.options['use_source_stepping'] and self.solve_source_stepping(A, z, x0):
            return self.x
        if self.options['use_gmin_stepping'] and self.solve_gmin_stepping(A, z, x0):
            return self.x
        return self.x


# --------------------------------------------------------------------------- #
# MultiToneHarmonicBalance
# --------------------------------------------------------------------------- #
HB_OPTIONS = {'reltol': 1e-3, 'abstol': 1e-6, 'maxiter': 100}


class HBOracle:
    """Restates MultiToneHarmonicBalance (MultiToneHarmonicBalance.py:23-728).

    ``pnp_extension=True`` enables the documented PNP polarity extension (the reference
    treats every BJT as NPN in HB); ``exact_jacobian=True`` re-zeroes dQdV each iteration
    (the reference accumulates it, :413 vs :438-440).  Both default to the reference
    behaviour.
    """

    def __init__(self, name, freq=1e9, numharmonics=10, pnp_extension=False, exact_jacobian=False, diode_charge=False):
        self.name = name
        self.freq = freq
        self.numharmonics = numharmonics
        self.options = HB_OPTIONS.copy()
        self.pnp_extension = pnp_extension
        self.exact_jacobian = exact_jacobian
        self.diode_charge = diode_charge     # extension: diodes carry charge in HB (see diode_charge_params)
        self.trace = None        # set to a list to record (J, F, dV, V) per Newton iteration
        self.level_log = []      # [(alpha, iterations, converged, total_error)] of the last run
        self.total_newton = 0
        self.total_lu = 0

    def get_node_idx(self, node):
        return self.netlist.get_node_idx(node) - 1

    def get_v(self, node):
        return self.Vf[self.get_node_idx(node)]

    # ---- setup, :245-302
    def setup_grid(self):
        if isinstance(self.freq, float):
            self.num_tones = 1
            self.f1 = self.freq
            self.K = self.numharmonics
            self.freqs = self.f1 * np.linspace(0, self.K, self.K + 1)
        elif isinstance(self.freq, list) and len(self.freq) == 2:
            self.num_tones = 2
            self.f1, self.f2 = self.freq
            self.K1, self.K2 = self.numharmonics
            self.K = (2 * self.K2 + 1) * self.K1 + self.K2
            self.freqs = np.zeros((self.K + 1))
            i = 0
            for k1 in range(self.K1 + 1):
                for k2 in range(-self.K2, self.K2 + 1):
                    if k1 == 0 and k2 < 0:
                        continue
                    self.freqs[i] = k1 * self.f1 + k2 * self.f2
                    i += 1
        else:
            raise ValueError('only one and two-tones are currently supported')
        self.S = 2 * self.K + 1
        self.N = self.netlist.get_num_nodes() - 1
        self.W = self.S * self.N
        S, K = self.S, self.K
        # :294-302
        s = np.arange(S)[:, None]
        k = np.arange(1, K + 1)[None, :]
        self.IDFT = np.ones((S, S))
        self.IDFT[:, 1::2] = np.cos(2 * np.pi * k * s / S)
        self.IDFT[:, 2::2] = np.sin(2 * np.pi * k * s / S)
        self.DFT = np.linalg.inv(self.IDFT)
        # rows of V that carry omega, and the fft sign flip rows (two-tone negative bins, :723-726)
        self.omega = 2 * np.pi * np.abs(self.freqs)
        self.negbin = (self.freqs < 0) if self.num_tones == 2 else np.zeros(K + 1, dtype=bool)

    # ---- :647-684
    def calc_Is(self):
        S = self.S
        Is = np.zeros((self.W, 1))
        for dev in self.lin_devs:
            if dev.kind != 'isource':
                continue
            if dev.itype == 'ac':
                iac = dev.ac * np.exp(1j * dev.phase)
                fidx = None
                if self.num_tones == 1 and dev.freq == self.f1:
                    fidx = 1
                elif self.num_tones == 2:
                    if dev.freq == self.f1:
                        fidx = 4 * self.K2 + 1
                    elif dev.freq == self.f2:
                        fidx = 1
                if fidx is None:
                    raise ValueError('ac source frequency matches no tone (reference leaves fidx unbound)')
                n1 = dev.n1 - 1
                if n1 >= 0:
                    Is[S * n1 + fidx] += iac.real
                    Is[S * n1 + fidx + 1] += iac.imag
                n2 = dev.n2 - 1
                if n2 >= 0:
                    Is[S * n2 + fidx] -= iac.real
                    Is[S * n2 + fidx + 1] -= iac.imag
            elif dev.itype == 'dc':
                n1 = dev.n1 - 1
                if n1 >= 0:
                    Is[S * n1] += dev.dc
                n2 = dev.n2 - 1
                if n2 >= 0:
                    Is[S * n2] -= dev.dc
        return Is

    # ---- :686-695 with the per-device add_mthb_stamps
    def calc_Y(self):
        S = self.S
        Y = np.zeros((self.W, self.W))

        def two_terminal(n1, n2, k, y):
            if k == 0:
                n, m = (n1 - 1) * S, (n2 - 1) * S
                if n1 > 0:
                    Y[n, n] += y
                if n2 > 0:
                    Y[m, m] += y
                if n1 > 0 and n2 > 0:
                    Y[n, m] -= y
                    Y[m, n] -= y
            else:
                n = (n1 - 1) * S + 2 * (k - 1) + 1
                m = (n2 - 1) * S + 2 * (k - 1) + 1
                y = complex(y)
                Ymnk = np.array([[y.real, -y.imag], [y.imag, y.real]])
                if n1 > 0:
                    Y[n:n + 2, n:n + 2] += Ymnk
                if n2 > 0:
                    Y[m:m + 2, m:m + 2] += Ymnk
                if n1 > 0 and n2 > 0:
                    Y[n:n + 2, m:m + 2] -= Ymnk
                    Y[m:m + 2, n:n + 2] -= Ymnk

        for dev in self.lin_devs:
            for k in range(0, self.K + 1):
                f = np.abs(self.freqs[k])
                kind = dev.kind
                if kind == 'resistor':       # Resistor.py:43-65
                    two_terminal(dev.n1, dev.n2, k, 1. / dev.R)
                elif kind == 'capacitor':    # Capacitor.py:62-85
                    two_terminal(dev.n1, dev.n2, k, 1e-12 if k == 0 else 1j * 2 * np.pi * f * dev.C)
                elif kind == 'inductor':     # Inductor.py:64-87
                    two_terminal(dev.n1, dev.n2, k, 1e6 if k == 0 else 1. / (1j * 2 * np.pi * f * dev.L))
                elif kind == 'ihf':          # IdealHarmonicFilter.py:33-57
                    two_terminal(dev.n1, dev.n2, k, 1e-12 if k == 0 else (dev.g if f == dev.freq else 1e-12))
                elif kind == 'gyrator':      # Gyrator.py:38-73 (hard-coded 1, ignores G)
                    if k == 0:
                        idx = [(n - 1) * S for n in dev.nodes]
                        w = 1
                    else:
                        idx = [(n - 1) * S + 2 * (k - 1) + 1 for n in dev.nodes]
                        w = 2
                    n1, n2, n3, n4 = idx
                    E = np.eye(w)
                    a1, a2, a3, a4 = dev.nodes

                    def add(r, c, sgn):
                        Y[r:r + w, c:c + w] += sgn * E
                    if a1 > 0 and a2 > 0:
                        add(n1, n2, +1)
                        add(n2, n1, -1)
                    if a1 > 0 and a3 > 0:
                        add(n1, n3, -1)
                        add(n3, n1, +1)
                    if a2 > 0 and a4 > 0:
                        add(n2, n4, +1)
                        add(n4, n2, -1)
                    if a3 > 0 and a4 > 0:
                        add(n3, n4, -1)
                        add(n4, n3, +1)
                # isource: consumed by calc_Is; anything else has no add_mthb_stamps (:692)
        return Y

    # ---- :697-704
    def calc_V0(self):
        dc = DCOracle(self.netlist)
        X = dc.run()
        self.dc_newton = dc.total_newton
        V = np.zeros((self.W, 1))
        for n in range(self.N):
            V[n * self.S] = X[n]
        return V

    # ---- :706-728
    def ifft(self, X):
        return (self.IDFT @ X.reshape(self.N, self.S).T).T.reshape(self.W, 1)

    def fft(self, xt):
        X = (self.DFT @ xt.reshape(self.N, self.S).T).T.copy()
        if self.num_tones == 2:
            X[:, 2::2] = np.where(self.negbin[1:][None, :], -X[:, 2::2], X[:, 2::2])
        return X.reshape(self.W, 1)

    def apply_omega(self, X):
        """Omega @ X for the 2x2-block-diagonal Omega of :416-421 (X is W x anything)."""
        S, N = self.S, self.N
        Xr = X.reshape(N, S, -1)
        out = np.zeros_like(Xr)
        w = self.omega[1:][None, :, None]
        # row i = n S + 2k-1 (Re):  -w * X[i+1];  row i+1 (Im):  +w * X[i]
        out[:, 1::2, :] = -w * Xr[:, 2::2, :]
        out[:, 2::2, :] = w * Xr[:, 1::2, :]
        return out.reshape(X.shape)

    def _block(self, g):
        # DFT @ diag(g) @ IDFT, :459 / :514-522
        return (self.DFT * g[None, :]) @ self.IDFT

    # ---- :633-645
    def hb_converged(self, Il, Inl):
        abstol, reltol = self.options['abstol'], self.options['reltol']
        with np.errstate(all='ignore'):
            F = Il + Inl
            Frel = 2 * abs(F / (Il - Inl + 1e-15))
            bad = (abs(F) > abstol) & (Frel > reltol)
        return not bool(np.any(bad))

    def eval_devices(self, vt, vtprev, dIdV, dQdV, it, qt):
        """Device sweep of hb_loop, :441-570.  dIdV/it/qt were zeroed by the caller; dQdV accumulates."""
        S = self.S

        def node_wave(x, n):
            return x[n * S:(n + 1) * S, 0] if n >= 0 else np.zeros(S)

        for dev in self.nonlin_devs:
            if dev.kind in ('diode', 'cubicnl'):
                n1, n2 = dev.n1 - 1, dev.n2 - 1
                vd = node_wave(vt, n1) - node_wave(vt, n2)
                vdold = node_wave(vtprev, n1) - node_wave(vtprev, n2)
                if dev.kind == 'diode':
                    gd, Id, vlim = diode_mthb_params(dev.options, vd, vdold, with_vlim=True)
                    if self.diode_charge:  # extension: fresh every iteration (dQd is zeroed by hb_loop), never accumulated
                        Cd, Qd = diode_charge_params(dev.options, vlim, gd, Id)
                        cf = self._block(np.asarray(Cd, dtype=float))
                        Qd = np.asarray(Qd, dtype=float).reshape(S, 1)
                        for na, sg in ((n1, +1.), (n2, -1.)):
                            if na >= 0:
                                qt[na * S:(na + 1) * S] += sg * Qd
                                self.dQd[na * S:(na + 1) * S, na * S:(na + 1) * S] += cf
                        if n1 >= 0 and n2 >= 0:
                            self.dQd[n1 * S:(n1 + 1) * S, n2 * S:(n2 + 1) * S] -= cf
                            self.dQd[n2 * S:(n2 + 1) * S, n1 * S:(n1 + 1) * S] -= cf
                else:
                    gd, Id = cubic_mthb_params(dev.alpha, vd)
                Id = np.asarray(Id, dtype=float).reshape(S, 1)
                gf = self._block(np.asarray(gd, dtype=float))
                if n1 >= 0:
                    i = n1 * S
                    dIdV[i:i + S, i:i + S] += gf
                    it[i:i + S] += Id
                if n2 >= 0:
                    i = n2 * S
                    dIdV[i:i + S, i:i + S] += gf
                    it[i:i + S] -= Id
                if n1 >= 0 and n2 >= 0:
                    i, j = n1 * S, n2 * S
                    dIdV[i:i + S, j:j + S] -= gf
                    dIdV[j:j + S, i:i + S] -= gf
            elif dev.kind == 'bjt':
                B, C, E = dev.n1 - 1, dev.n2 - 1, dev.n3 - 1
                sign = float(dev.type) if self.pnp_extension else 1.0
                out = bjt_hb_params(dev.options, node_wave(vt, B), node_wave(vt, C), node_wave(vt, E),
                                    node_wave(vtprev, B), node_wave(vtprev, C), node_wave(vtprev, E), sign)
                Ib, Ic, Ie, Qbe, Qbc, Qsc = (np.asarray(x, dtype=float).reshape(S, 1) for x in out[:6])
                gmu, gpi, gmf, gmr, Cbc, Cbe, Cbebc, Csc = (np.asarray(x, dtype=float) for x in out[6:])
                dev.Ic = Ic  # side channel, :512
                gmuf, gpif, gmff, gmrf = (self._block(x) for x in (gmu, gpi, gmf, gmr))
                Cbef, Cbcf, Cbebcf, Cscf = (self._block(x) for x in (Cbe, Cbc, Cbebc, Csc))
                sB, sC, sE = (slice(n * S, n * S + S) for n in (B, C, E))
                if B >= 0:
                    dIdV[sB, sB] += (gmuf + gpif)
                    dQdV[sB, sB] += (Cbcf + Cbef + Cbebcf)
                    it[sB] += Ib
                    qt[sB] += (Qbe + Qbc)
                    if C >= 0:
                        dIdV[sB, sC] += (- gmuf)
                        dIdV[sC, sB] += (- gmuf + gmff + gmrf)
                        dQdV[sB, sC] += (- Cbcf - Cbebcf)
                        dQdV[sC, sB] += (- Cbcf)
                    if E >= 0:
                        dIdV[sB, sE] += (- gpif)
                        dIdV[sE, sB] += (- gpif - gmff - gmrf)
                        dQdV[sB, sE] += (- Cbef)
                        dQdV[sE, sB] += (- Cbef - Cbebcf)
                if C >= 0:
                    dIdV[sC, sC] += (+ gmuf - gmrf)
                    dQdV[sC, sC] += (+ Cbcf + Cscf)
                    it[sC] += Ic
                    qt[sC] -= Qbc - Qsc
                    if E >= 0:
                        dIdV[sC, sE] += (- gmff)
                        dIdV[sE, sC] += (+ gmrf)
                        dQdV[sE, sC] += (+ Cbebcf)
                if E >= 0:
                    dIdV[sE, sE] += (+ gpif + gmff)
                    dQdV[sE, sE] += (+ Cbef)
                    it[sE] -= Ie
                    qt[sE] -= Qbe

    # ---- :407-631
    def hb_loop(self, Is, Y, V):
        W = self.W
        vt = self.ifft(V)
        it = np.zeros((W, 1))
        qt = np.zeros((W, 1))
        dIdV = np.zeros((W, W))
        dQdV = np.zeros((W, W))
        maxiter = self.options['maxiter']
        itercnt = 0
        dV = None
        while True:
            vtprev = vt.copy()
            vt = self.ifft(V)
            dIdV[:, :] = 0
            it[:] = 0
            qt[:] = 0
            if self.exact_jacobian:
                dQdV[:, :] = 0
            if self.diode_charge:
                self.dQd = np.zeros((W, W))
            self.eval_devices(vt, vtprev, dIdV, dQdV, it, qt)
            J = Y + dIdV + self.apply_omega(dQdV + self.dQd if self.diode_charge else dQdV)
            Il = Y @ V - Is
            Inl = self.fft(it) + self.apply_omega(self.fft(qt))
            F = Il + Inl
            self.last_Inl = Inl          # kept for checks of the KCL-based oscillator objective (not in the reference)
            converged = self.hb_converged(Il, Inl)
            itercnt += 1
            self.total_newton += 1
            if (converged and itercnt >= 3) or itercnt >= maxiter:
                self.last_error = float(np.sum(np.abs(F)))
                self.last_itercnt = itercnt
                if self.trace is not None:
                    self.trace.append(dict(J=J.copy(), F=F.copy(), V=V.copy(), dV=None, vt=vt.copy(), vtprev=vtprev))
                return V, converged
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                lu, piv = scipy.linalg.lu_factor(J, check_finite=False)
                dVnew = scipy.linalg.lu_solve((lu, piv), F, check_finite=False)
            self.total_lu += 1
            if np.any(np.isnan(dVnew)):
                if dV is None:
                    raise FloatingPointError('NaN in the first Newton step (reference raises NameError here)')
            else:
                dV = dVnew
            if self.trace is not None:
                self.trace.append(dict(J=J.copy(), F=F.copy(), V=V.copy(), dV=dV.copy(), vt=vt.copy(), vtprev=vtprev))
            V = V - dV

    # ---- :234-405
    def run(self, netlist, V0=None):
        self.netlist = netlist.copy()
        self.lin_devs = self.netlist.get_linear_devices()
        self.nonlin_devs = self.netlist.get_nonlinear_devices()
        self.setup_grid()
        Is = self.calc_Is()
        Y = self.calc_Y()
        self.level_log = []
        V = self.calc_V0() if V0 is None else V0
        if V0 is None:
            converged = False
        else:
            V, converged = self.hb_loop(Is, Y, V)
            self.level_log.append((1.0, self.last_itercnt, converged, self.last_error))
            if not converged:
                V = self.calc_V0()
        if not converged:
            Vprev = V.copy()
            maxiter, alpha, inc, dec = 50, 0.1, 1.4, 1.2
            converged = False
            itercnt = 0
            rows = np.ones((self.N, self.S), dtype=bool)
            rows[:, 0] = False
            rows = rows.reshape(self.W)
            while (not converged) and (itercnt < maxiter):
                Isalpha = Is.copy()
                Isalpha[rows] *= alpha
                V, iterconverged = self.hb_loop(Isalpha, Y, V)
                self.level_log.append((alpha, self.last_itercnt, iterconverged, self.last_error))
                if iterconverged:
                    if alpha >= 1.0:
                        converged = True
                        break
                    Vprev = V.copy()
                    alpha = inc * alpha
                    if alpha >= 1.0:
                        alpha = 1.0
                else:
                    V = Vprev.copy()
                    if alpha >= 1.0:
                        dec = 1 + 1.5 * (dec - 1)
                        inc = 1 + (inc - 1) / 1.5
                    alpha = alpha / dec
                    if alpha <= 0.01:
                        converged = False
                        break
                itercnt += 1
        if not converged:
            return False, None, None, None, None
        if self.num_tones == 2:
            self.freqs = np.abs(self.freqs)
        self.V = V
        # :395-403
        Vr = V.reshape(self.N, self.S)
        self.Vf = np.zeros((self.N, self.K + 1), dtype=complex)
        self.Vf[:, 0] = Vr[:, 0]
        a, b = Vr[:, 1::2], Vr[:, 2::2]
        self.Vf[:, 1:] = np.sqrt(a ** 2 + b ** 2) * np.exp(1j * np.arctan2(b, a))
        return converged, self.freqs, self.Vf, None, None

    # ---- :137-232
    def run_oscillator(self, netlist, f0, numharmonics, V0, node, useprev=False):
        netlist = netlist.copy()
        self.freq = f0
        self.numharmonics = numharmonics
        Voscprobe = netlist.add_iac(self.name + '.I.Oscprobe', self.name + '.nosc1', 'gnd', ac=V0, freq=f0)
        netlist.add_gyrator(self.name + '.G.Oscprobe', self.name + '.nosc1', self.name + '.nosc2', 'gnd', 'gnd', 1)
        Zoscprobe = netlist.add_idealharmonicfilter(self.name + 'IHF.Oscprobe', self.name + '.nosc2', node, f0)
        self.osc_evals = []

        class SmallEnough(Exception):
            pass

        def objFunc(x, info):
            fosc, Vosc = x[0], x[1]
            Voscprobe.ac = Vosc
            Voscprobe.freq = fosc
            Zoscprobe.freq = fosc
            self.freq = fosc
            if info['itercnt'] > 0 or info['useprev']:
                converged, freqs, Vf, _, _ = self.run(netlist, self.V)
            else:
                converged, freqs, Vf, _, _ = self.run(netlist)
            if not converged:
                self.osc_evals.append((fosc, Vosc, 1e2))
                return 1e2
            n1 = self.get_node_idx(info['osc_node'])
            n2 = self.get_node_idx(self.name + '.nosc2')
            Voscx = Vf[n1, 1]
            Iosc = (Vf[n1, 1] - Vf[n2, 1]) * Zoscprobe.g
            Yosc = Iosc / Voscx
            info['itercnt'] += 1
            self.osc_evals.append((fosc, Vosc, abs(Yosc)))
            if abs(Yosc) < 1e-12:
                info['result'] = x
                raise SmallEnough()
            return abs(Yosc)

        args = ({'itercnt': 0, 'osc_node': node, 'result': 0, 'useprev': useprev},)
        try:
            xopt = optimize.fmin(func=objFunc, x0=[f0, V0], args=args, xtol=1e-12, ftol=1e-12,
                                 maxfun=300, disp=False, retall=False, full_output=True)[0]
        except SmallEnough:
            xopt = args[0]['result']
        fosc, Vosc = xopt[0], xopt[1]
        Voscprobe.ac = Vosc
        Voscprobe.freq = fosc
        Zoscprobe.freq = fosc
        self.freq = fosc
        self.fosc, self.Vosc = fosc, Vosc
        return self.run(netlist, self.V)


def clone_netlist(netlist):
    """Deep copy (separate device records) for sweeps evaluated in worker processes."""
    return copy.deepcopy(netlist)
