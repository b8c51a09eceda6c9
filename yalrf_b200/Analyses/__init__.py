from .MultiToneHarmonicBalance import MultiToneHarmonicBalance
from .DC import DC
