"""Netlist container with the reference's API (yalrf/Netlist.py:8-897), restricted to the devices the
harmonic-balance path can reach."""
from .Devices import (BJT, Capacitor, CubicNonLinearity, CurrentSource, Diode, Gyrator, IdealHarmonicFilter,
                      Inductor, Resistor)


class Netlist():
    """Device list + node name <-> index maps.  ``add_*`` return the device object, whose attributes and
    ``options`` dict the caller may edit until (and between) runs."""

    def __init__(self, name):
        self.name = name
        self.devices = []
        self.node_name_to_idx = {'gnd': 0}
        self.node_idx_to_name = ['gnd']

    def _add(self, dev):
        self.devices.append(dev)
        return dev

    def add_resistor(self, name, n1, n2, value):
        return self._add(Resistor(name, self.add_node(n1), self.add_node(n2), value))

    def add_capacitor(self, name, n1, n2, value):
        return self._add(Capacitor(name, self.add_node(n1), self.add_node(n2), value))

    def add_inductor(self, name, n1, n2, value):
        return self._add(Inductor(name, self.add_node(n1), self.add_node(n2), value))

    def add_idc(self, name, n1, n2, dc):
        return self._add(CurrentSource(name, self.add_node(n1), self.add_node(n2), itype='dc', dc=dc))

    def add_iac(self, name, n1, n2, ac, phase=0, freq=0):
        return self._add(CurrentSource(name, self.add_node(n1), self.add_node(n2), itype='ac', ac=ac,
                                       phase=phase, freq=freq))

    def add_isource(self, name, n1, n2, dc, ac, phase=0):
        # Netlist.py:265-295: itype 'both' -- stamped by DC, ignored by calc_Is (MultiToneHarmonicBalance.py:652,675)
        return self._add(CurrentSource(name, self.add_node(n1), self.add_node(n2), itype='both', dc=dc, ac=ac,
                                       phase=phase))

    def add_gyrator(self, name, n1, n2, n3, n4, G):
        return self._add(Gyrator(name, self.add_node(n1), self.add_node(n2), self.add_node(n3),
                                 self.add_node(n4), G))

    def add_idealharmonicfilter(self, name, n1, n2, freq):
        return self._add(IdealHarmonicFilter(name, self.add_node(n1), self.add_node(n2), freq))

    def add_diode(self, name, n1, n2):
        return self._add(Diode(name, self.add_node(n1), self.add_node(n2)))

    def add_bjt(self, name, n1, n2, n3, n4='gnd', ispnp=False):
        return self._add(BJT(name, self.add_node(n1), self.add_node(n2), self.add_node(n3), self.add_node(n4),
                             ispnp))

    def add_cubicnl(self, name, n1, n2, alpha):
        return self._add(CubicNonLinearity(name, self.add_node(n1), self.add_node(n2), alpha))

    def _unsupported(self, what):
        raise NotImplementedError(
            what + ' has no add_mthb_stamps in the reference (it is silently dropped by its harmonic balance, '
            'MultiToneHarmonicBalance.py:692) and is outside this engine; model sources as current source + '
            'gyrator like the reference HB scripts do')

    # Extension (off by default = reference behaviour): with ``netlist.vsource_extension = True`` an ideal
    # voltage source is entered the way the reference's HB scripts write it by hand (tests/HB_Diode.py:12-13) -- a current
    # source of the same value into an internal node and a unit gyrator from that node to (n1, n2), which forces
    # V(n1) - V(n2) = value -- instead of being refused.  The returned object forwards dc / ac / phase / freq to the
    # internal current source, so scripts can keep editing the source between runs.
    vsource_extension = False

    def _vsource(self, name, n1, n2, itype, **kw):
        if not self.vsource_extension:
            self._unsupported('VoltageSource')
        from .Devices import VoltageSourceEquivalent
        x = name + '.x'
        isrc = self._add(CurrentSource(name + '.I', self.add_node(x), 0, itype=itype, **kw))
        gyr = self.add_gyrator(name + '.G', x, n1, n2, 'gnd', 1)
        return VoltageSourceEquivalent(name, isrc, gyr)

    def add_vdc(self, name, n1, n2, dc):
        return self._vsource(name, n1, n2, 'dc', dc=dc)

    def add_vac(self, name, n1, n2, ac, phase=0, freq=0):
        """freq: the tone this source belongs to (the reference's VoltageSource has no frequency of its own)"""
        return self._vsource(name, n1, n2, 'ac', ac=ac, phase=phase, freq=freq)

    def add_vsource(self, *a, **k):
        self._unsupported('VoltageSource with DC and AC values (the HB source vector ignores such sources, :652)')

    add_vpulse = add_vsine = add_vsource

    def add_vcvs(self, *a, **k):
        self._unsupported('controlled source')

    add_vccs = add_ccvs = add_cccs = add_vcvs

    def add_opamp(self, *a, **k):
        self._unsupported('Opamp')

    # Netlist.py:703-725
    def add_node(self, n):
        if n not in self.node_name_to_idx:
            self.node_name_to_idx[n] = len(self.node_name_to_idx)
            self.node_idx_to_name.append(n)
        return self.node_name_to_idx[n]

    def is_nonlinear(self):
        return True  # Netlist.py:727-732 returns True on every path

    def get_num_nodes(self):
        return len(self.node_idx_to_name)

    def get_num_vsources(self):
        return sum(dev.get_num_vsources() for dev in self.devices)

    def get_device(self, name):
        for dev in self.devices:
            if dev.name == name:
                return dev
        return None

    def get_devices(self):
        return self.devices

    def get_linear_devices(self):
        return [dev for dev in self.devices if not dev.is_nonlinear()]

    def get_nonlinear_devices(self):
        return [dev for dev in self.devices if dev.is_nonlinear()]

    def get_node_idx(self, name):
        return self.node_name_to_idx.get(name)

    def get_node_name(self, idx):
        return self.node_idx_to_name[idx] if idx < len(self.node_idx_to_name) else None

    def get_voltage_idx(self, name):
        return self.node_name_to_idx[name] - 1 if name in self.node_name_to_idx else None

    # Netlist.py:889-897: shallow -- the copy shares the device objects
    def copy(self):
        netlist = Netlist(self.name)
        netlist.devices = self.devices.copy()
        netlist.node_name_to_idx = self.node_name_to_idx.copy()
        netlist.node_idx_to_name = self.node_idx_to_name.copy()
        return netlist
