"""yalrf_b200 -- B200-native harmonic-balance engine behind YalRF's Netlist / MultiToneHarmonicBalance API."""
from .Netlist import Netlist
from .YalRF import YalRF

__all__ = ['Netlist', 'YalRF']
