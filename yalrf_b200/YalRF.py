"""``YalRF(Netlist)`` facade (yalrf/YalRF.py:9-324): only the DC analysis registry survives here, because
the DC operating point is the cold-start guess of harmonic balance; AC and transient are other analyses."""
from .Netlist import Netlist


class YalRF(Netlist):

    def __init__(self, name):
        Netlist.__init__(self, name)
        self.analyses = {}

    # YalRF.py:106-128
    def add_dc_analysis(self, name):
        from .Analyses.DC import DC
        dc = DC(name)
        self.analyses[name] = dc
        return dc

    # YalRF.py:130-156
    def run(self, name, x0=None, nodeset=None):
        return self.analyses[name].run(self)

    def get_analysis(self, name):
        return self.analyses.get(name)

    # YalRF.py:180-209 (DC only)
    def get_voltage(self, analysis, node):
        x = self.analyses[analysis].get_dc_solution()
        idx = self.get_voltage_idx(node)
        return x[idx, 0] if idx is not None and idx >= 0 else 0.0

    # YalRF.py:281-298
    def print_dc_voltages(self, name):
        x = self.analyses[name].get_dc_solution()
        print('DC Voltages:')
        for i in range(1, len(self.node_idx_to_name)):
            print('{}:\t{:0.4f} V'.format(self.node_idx_to_name[i], x[i - 1, 0]))
        print()
