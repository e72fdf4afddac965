"""Physical constants and record sizes used by the line-by-line path.

Two sets coexist in the reference and both are kept:
  * the CODATA values of pyratbay/constants/astrophysical_constants.py:66-113 (taken from
    scipy.constants like the reference does) -- used for number densities and for the
    Voigt-grid width bounds;
  * the older CGS values hard-wired in src_c/include/constants.h:11-21 -- used inside the
    extinction kernel (they live in csrc/common.cuh, not here).
"""
import scipy.constants as sc

# Universal constants in CGS units (astrophysical_constants.py:66-70)
h = sc.h * 1e7
k = sc.k * 1e7
c = sc.c * 1e2

# Conversion factors (astrophysical_constants.py:81-113)
A = 1e-8
nm = 1e-7
um = 1e-4
cm = 1.0
barye = 1.0
mbar = 1e3
pascal = 1e1
bar = 1e6
atm = 1.01e6
amu = sc.physical_constants['unified atomic mass unit'][0] * 1e3

# TLI record lengths (constants/code_constants.py:37-41)
tlireclen = 26
dreclen = 8
ireclen = 4
sreclen = 2

_UNITS = {
    'A': A, 'nm': nm, 'um': um, 'cm': cm, 'mm': 0.1, 'm': 100.0,
    'barye': barye, 'mbar': mbar, 'pascal': pascal, 'bar': bar, 'atm': atm,
    'kelvin': 1.0, 'K': 1.0,
}


def u(units):
    """Conversion factor to CGS for a unit name (tools.u, tools/tools.py:380-400)."""
    if units not in _UNITS:
        raise ValueError(f"Units '{units}' does not exist in pyratbay_b200.constants")
    return _UNITS[units]
