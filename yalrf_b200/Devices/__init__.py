"""Device records with the reference's class names, constructor signatures and mutable attributes.

The reference's device classes (yalrf/Devices/*.py) carry the numerics of every analysis as methods
(add_dc_stamps, add_mthb_stamps, get_hb_params ...).  Here they are parameter holders only: the
harmonic-balance numerics of these devices run on the GPU (yalrf_b200/csrc), and the objects are
re-read at every ``run`` because users mutate them between runs (``i2.ac = vin / 2``,
``q1.options['Is'] = ...``, ``q2.options = q1.options``; tests/HB_DiffAmp.py:67-69, :116-117).
"""
from math import inf

import numpy as np


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


class IdealHarmonicFilter(_Device):
    """yalrf/Devices/IdealHarmonicFilter.py:5-12"""
    kind = 'ihf'

    def __init__(self, name, n1, n2, value):
        self.name = name
        self.n1 = n1
        self.n2 = n2
        self.freq = float(value)
        self.g = 1e9  # conductance of 'shorted' harmonics


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

    def __str__(self):
        return 'CubicNL: {}\nNodes = {} -> {}\n'.format(self.name, self.n1, self.n2)
