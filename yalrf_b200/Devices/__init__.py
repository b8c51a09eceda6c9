"""Device records with the reference's class names, constructor signatures and mutable attributes.

The reference's device classes (yalrf/Devices/*.py) carry the numerics of every analysis as methods
(add_dc_stamps, add_mthb_stamps, get_hb_params ...).  Here they are parameter holders: the
harmonic-balance numerics of these devices run on the GPU (yalrf_b200/csrc), and the objects are
re-read at every ``run`` because users mutate them between runs (``i2.ac = vin / 2``,
``q1.options['Is'] = ...``, ``q2.options = q1.options``; tests/HB_DiffAmp.py:67-69, :116-117).

The harmonic-balance part of the reference's device protocol is kept for scripts that call it directly:
``add_mthb_stamps(Y, S, freq, freqidx)`` of the linear classes writes the element's 2 x 2 blocks into a caller-owned
dense matrix (pure formatting of one admittance value; the engine itself never builds that matrix), and
``get_mthb_params`` / ``get_hb_params`` of the nonlinear classes evaluate ONE sample of the model -- on the GPU,
through the same kernel code the Newton loop runs (yhb_stage_device_eval); there is no CPU model.
"""
from math import inf

import numpy as np


def _stamp_two_terminal(Y, S, n1, n2, freqidx, re, im=0.0):
    """+y on the diagonal blocks, -y on the off-diagonal blocks of nodes n1, n2 (0 = gnd) at harmonic freqidx:
    DC is one real entry, harmonic k the block [[re, -im], [im, re]] at rows / columns (n - 1) S + 2 k - 1."""
    if freqidx == 0:
        n, m = (n1 - 1) * S, (n2 - 1) * S
        blk = re
        sl = lambda a: a
    else:
        n, m = (n1 - 1) * S + 2 * (freqidx - 1) + 1, (n2 - 1) * S + 2 * (freqidx - 1) + 1
        blk = np.array([[re, -im], [im, re]])
        sl = lambda a: slice(a, a + 2)
    if n1 > 0:
        Y[sl(n), sl(n)] += blk
    if n2 > 0:
        Y[sl(m), sl(m)] += blk
    if n1 > 0 and n2 > 0:
        Y[sl(n), sl(m)] -= blk
        Y[sl(m), sl(n)] -= blk


_model_cache = {}


def _model_sample(kind, dev, v_now, v_old):
    """One sample of a nonlinear device model on the GPU: v_now / v_old = terminal voltages of this iterate and of the
    previous one (the limiter reference).  Returns the model outputs of yhb_stage_device_eval at that sample."""
    import ctypes as C
    import torch
    from .. import _engine as E, _lib as L
    from .._flatten import ParamBlock
    from ..Netlist import Netlist
    if not torch.cuda.is_available():
        raise RuntimeError('yalrf_b200 needs a CUDA device: there is no CPU fallback')
    nodes = ['a', 'b', 'c'][:len(v_now)]
    y = Netlist('sample')
    if kind == 'diode':
        twin = y.add_diode('D', 'a', 'b')
        twin.options = dev.options
    elif kind == 'bjt':
        twin = y.add_bjt('Q', 'a', 'b', 'c', ispnp=dev.type < 0)
        twin.options = dev.options
    else:
        twin = y.add_cubicnl('N', 'a', 'b', dev.alpha)
    plan, topo = E.get_plan(y, 1.0, 1)
    pb = ParamBlock(topo, 1.0, 1).finalize()
    d = torch.device('cuda', torch.cuda.current_device())
    L.check(L.lib().yhb_set_device(d.index))
    t = {n: torch.from_numpy(getattr(pb, n)).to(d) for n in ('tones', 'lin_val', 'src_val', 'diode_par', 'bjt_par', 'cubic_par')}
    p = L.Params()
    p.batch = 1
    for n, v in t.items():
        setattr(p, n, v.data_ptr())
    S = plan.S
    vt = torch.tensor(np.repeat(np.asarray(v_now, dtype=np.float64), S)[None], device=d)
    vo = torch.tensor(np.repeat(np.asarray(v_old, dtype=np.float64), S)[None], device=d)
    dout = torch.zeros((1, 1, 2, S), dtype=torch.float64, device=d)
    bout = torch.zeros((1, 1, 14, S), dtype=torch.float64, device=d)
    cout = torch.zeros((1, 1, 2, S), dtype=torch.float64, device=d)
    L.check(L.lib().yhb_stage_device_eval(plan.handle, C.byref(p), vt.data_ptr(), vo.data_ptr(), dout.data_ptr(),
                                          bout.data_ptr(), cout.data_ptr(), None))
    torch.cuda.synchronize()
    out = {'diode': dout, 'bjt': bout, 'cubicnl': cout}[kind]
    return out[0, 0, :, 0].cpu().numpy()


class _Device:
    kind = None

    def is_nonlinear(self):
        return False

    def get_num_vsources(self):
        return 0

    def init(self):
        pass


class Resistor(_Device):
    """yalrf/Devices/Resistor.py:5-9"""
    kind = 'resistor'

    def __init__(self, name, n1, n2, value):
        self.name = name
        self.n1 = n1
        self.n2 = n2
        self.R = float(value)

    def add_mthb_stamps(self, Y, S, freq, freqidx):
        """Resistor.py:43-65"""
        _stamp_two_terminal(Y, S, self.n1, self.n2, freqidx, 1. / self.R)

    def __str__(self):
        return 'Resistor: {}\nNodes = {} -> {}\nValue = {}\n'.format(self.name, self.n1, self.n2, self.R)


class Capacitor(_Device):
    """yalrf/Devices/Capacitor.py:5-11"""
    kind = 'capacitor'

    def __init__(self, name, n1, n2, value):
        self.name = name
        self.n1 = n1
        self.n2 = n2
        self.C = float(value)

    def add_mthb_stamps(self, Y, S, freq, freqidx):
        """Capacitor.py:62-85: 1e-12 S at DC, j w C at a harmonic"""
        if freqidx == 0:
            _stamp_two_terminal(Y, S, self.n1, self.n2, 0, 1e-12)
        else:
            _stamp_two_terminal(Y, S, self.n1, self.n2, freqidx, 0., 2. * np.pi * freq * self.C)

    def __str__(self):
        return 'Capacitor: {}\nNodes = {} -> {}\nValue = {}\n'.format(self.name, self.n1, self.n2, self.C)


class Inductor(_Device):
    """yalrf/Devices/Inductor.py:5-9 (one branch row in DC, none in HB)"""
    kind = 'inductor'

    def __init__(self, name, n1, n2, value):
        self.name = name
        self.n1 = n1
        self.n2 = n2
        self.L = float(value)

    def get_num_vsources(self):
        return 1

    def add_mthb_stamps(self, Y, S, freq, freqidx):
        """Inductor.py:64-87: 1e6 S at DC (no branch row in HB), 1 / (j w L) at a harmonic"""
        if freqidx == 0:
            _stamp_two_terminal(Y, S, self.n1, self.n2, 0, 1e6)
        else:
            _stamp_two_terminal(Y, S, self.n1, self.n2, freqidx, 0., -(1. / (2. * np.pi * freq * self.L)))

    def __str__(self):
        return 'Inductor: {}\nNodes = {} -> {}\nValue = {}\n'.format(self.name, self.n1, self.n2, self.L)


class CurrentSource(_Device):
    """yalrf/Devices/CurrentSource.py:5-12 (phase is stored in radians)"""
    kind = 'isource'

    def __init__(self, name, n1, n2, dc=0, ac=0, phase=0, freq=0, itype=None):
        self.name = name
        self.n1 = n1
        self.n2 = n2
        self.itype = itype
        self.dc = float(dc)
        self.ac = float(ac)
        self.phase = np.radians(float(phase))
        self.freq = float(freq)

    def __str__(self):
        return 'Current Source: {}\nNodes = {} -> {}\nIdc = {}\nIac = {}\nPhase = {}'.format(
            self.name, self.n1, self.n2, self.dc, self.ac, np.degrees(self.phase))


class Gyrator(_Device):
    """yalrf/Devices/Gyrator.py:5-11"""
    kind = 'gyrator'

    def __init__(self, name, n1, n2, n3, n4, G=1):
        self.name = name
        self.n1 = n1
        self.n2 = n2
        self.n3 = n3
        self.n4 = n4
        self.G = float(G)

    def add_mthb_stamps(self, Y, S, freq, freqidx):
        """Gyrator.py:38-73: unit stamps (self.G is ignored by the reference's HB)"""
        one = 1 if freqidx == 0 else np.eye(2)
        w = 1 if freqidx == 0 else 2

        def pos(n):
            return (n - 1) * S + (0 if freqidx == 0 else 2 * (freqidx - 1) + 1)
        for a, b, sg in ((self.n1, self.n2, +1), (self.n1, self.n3, -1), (self.n2, self.n4, +1), (self.n3, self.n4, -1)):
            if a > 0 and b > 0:
                ia, ib = pos(a), pos(b)
                if freqidx == 0:
                    Y[ia, ib] += sg * one
                    Y[ib, ia] -= sg * one
                else:
                    Y[ia:ia + w, ib:ib + w] += sg * one
                    Y[ib:ib + w, ia:ia + w] -= sg * one


class IdealHarmonicFilter(_Device):
    """yalrf/Devices/IdealHarmonicFilter.py:5-12"""
    kind = 'ihf'

    def __init__(self, name, n1, n2, value):
        self.name = name
        self.n1 = n1
        self.n2 = n2
        self.freq = float(value)
        self.g = 1e9  # conductance of 'shorted' harmonics

    def add_mthb_stamps(self, Y, S, freq, freqidx):
        """IdealHarmonicFilter.py:33-57: g where freq == self.freq exactly, 1e-12 S elsewhere (and at DC)"""
        _stamp_two_terminal(Y, S, self.n1, self.n2, freqidx, self.g if (freqidx > 0 and freq == self.freq) else 1e-12)


# Diode.py:6-36
diode_options = {
    'Temp': 300.0, 'Is': 1e-15, 'N': 1.0, 'Isr': 0.0, 'Nr': 2.0, 'Rs': 0.0, 'Cj0': 0.0, 'M': 0.5,
    'Vj': 0.7, 'Fc': 0.5, 'Cp': 0.0, 'Tt': 0.0, 'Bv': inf, 'Ibv': 0.001, 'Ikf': inf, 'Kf': 0.0,
    'Af': 1.0, 'Ffe': 1.0, 'Xti': 3.0, 'Eg': 1.11, 'Ega': 7.02e-4, 'Egb': 1108.0, 'Tbv': 0.0,
    'Trs': 0.0, 'Tt1': 0.0, 'Tt2': 0.0, 'Tm1': 0.0, 'Tm2': 0.0, 'Tnom': 300.0, 'Area': 1.0,
}


class Diode(_Device):
    """yalrf/Devices/Diode.py:41-63"""
    kind = 'diode'

    def __init__(self, name, n1, n2):
        self.name = name
        self.n1 = n1  # anode (+)
        self.n2 = n2  # cathode (-)
        self.options = diode_options.copy()

    def is_nonlinear(self):
        return True

    def get_mthb_params(self, Vd, Vdold):
        """Diode.py:395-427 -> (gd, Id) of one sample (Vdold = the previous iterate's voltage, the limiter reference)"""
        g, i = _model_sample('diode', self, [Vd, 0.], [Vdold, 0.])
        return float(g), float(i)

    def __str__(self):
        return 'Diode: {}\nNodes = {} -> {}\n'.format(self.name, self.n1, self.n2)


# BJT.py:5-52
bjt_options = {
    'Temp': 300.0, 'Is': 1e-15, 'Nf': 1.0, 'Nr': 1.0, 'Ikf': 1e9, 'Ikr': 1e9, 'Vaf': 1e12, 'Var': 1e12,
    'Ise': 0.0, 'Ne': 1.5, 'Isc': 0.0, 'Nc': 2.0, 'Bf': 100.0, 'Br': 1.0, 'Rbm': 0.0, 'Irb': 1e9,
    'Rc': 0.0, 'Re': 0.0, 'Rb': 0.0, 'Cje': 0.0, 'Vje': 0.75, 'Mje': 0.33, 'Cjc': 0.0, 'Vjc': 0.75,
    'Mjc': 0.33, 'Xcjc': 1.0, 'Cjs': 0.0, 'Vjs': 0.75, 'Mjs': 0.0, 'Fc': 0.5, 'Tf': 0.0, 'Xtf': 0.0,
    'Vtf': 1e12, 'Itf': 0.0, 'Ptf': 0.0, 'Tr': 0.0, 'Kf': 0.0, 'Af': 1.0, 'Ffe': 1.0, 'Kb': 0.0,
    'Ab': 1.0, 'Fb': 1.0, 'Xti': 3.0, 'Xtb': 0.0, 'Eg': 1.11, 'Tnom': 300.0, 'Area': 1.0,
}


class BJT(_Device):
    """yalrf/Devices/BJT.py:62-88"""
    kind = 'bjt'

    def __init__(self, name, n1, n2, n3, n4=0, ispnp=False):
        self.name = name
        self.n1 = n1  # base (B)
        self.n2 = n2  # collector (C)
        self.n3 = n3  # emitter (E)
        self.n4 = n4  # substrate (S)
        self.type = -1 if ispnp else 1
        self.options = bjt_options.copy()

    def is_nonlinear(self):
        return True

    def get_hb_params(self, Vb, Vc, Ve, Vs, s, Vbold, Vcold, Veold):
        """BJT.py:462-582 -> (Ib, Ic, Ie, Qbe, Qbc, Qsc, gmu, gpi, gmf, gmr, Cbc, Cbe, Cbebc, Csc) of one sample; Vs and the
        sample index s are accepted and ignored like the reference does (:504 forces Vs = 0)"""
        return tuple(float(v) for v in _model_sample('bjt', self, [Vb, Vc, Ve], [Vbold, Vcold, Veold]))

    def __str__(self):
        return 'BJT: {}\nNodes BCE nodes = {}, {}, {}\n'.format(self.name, self.n1, self.n2, self.n3)


class CubicNonLinearity(_Device):
    """yalrf/Devices/CubicNonLinearity.py:5-9"""
    kind = 'cubicnl'

    def __init__(self, name, n1, n2, alpha=1e-3):
        self.name = name
        self.n1 = n1
        self.n2 = n2
        self.alpha = alpha

    def is_nonlinear(self):
        return True

    def get_mthb_params(self, V, Vold):
        """CubicNonLinearity.py:36-41 -> (g, I) with g = 2 alpha V^2 as the reference stamps it"""
        g, i = _model_sample('cubicnl', self, [V, 0.], [Vold, 0.])
        return float(g), float(i)

    def __str__(self):
        return 'CubicNL: {}\nNodes = {} -> {}\n'.format(self.name, self.n1, self.n2)


class VoltageSourceEquivalent:
    """Handle of a voltage source entered as current source + unit gyrator (Netlist.vsource_extension): not a device of
    its own -- the netlist holds the two internal devices -- but the object scripts edit between runs."""

    def __init__(self, name, isrc, gyr):
        self.name, self._i, self._g = name, isrc, gyr
        self.n1, self.n2 = gyr.n2, gyr.n3

    dc = property(lambda self: self._i.dc, lambda self, v: setattr(self._i, 'dc', v))
    ac = property(lambda self: self._i.ac, lambda self, v: setattr(self._i, 'ac', v))
    phase = property(lambda self: self._i.phase, lambda self, v: setattr(self._i, 'phase', v))
    freq = property(lambda self: self._i.freq, lambda self, v: setattr(self._i, 'freq', v))
