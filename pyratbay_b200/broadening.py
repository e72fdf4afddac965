"""Width bounds for the Voigt grid (pyratbay/opacity/broadening/broadening.py:367-498)."""
import numpy as np

from . import constants as pc

_H2_RADIUS = 1.445e-8  # cm
_H2_MASS = 2.01588     # amu


def min_widths(min_temp, max_temp, min_wn, max_mass, min_rad, min_press):
    """Minimum Doppler and Lorentz HWHM (cm-1) for an H2-dominated atmosphere
    (broadening.py:367-428).  Pressure in bar, radius in cm, mass in amu."""
    dmin = (np.sqrt(2.0 * np.log(2.0) * pc.k * min_temp / (max_mass * pc.amu))
            * min_wn / pc.c)
    min_diam = _H2_RADIUS + min_rad
    lmin = (np.sqrt(2.0 / (np.pi * pc.k * max_temp * pc.amu))
            * min_press * pc.bar * min_diam**2.0 / pc.c
            * np.sqrt(1.0 / max_mass + 1.0 / _H2_MASS))
    return dmin, lmin


def max_widths(min_temp, max_temp, max_wn, min_mass, max_rad, max_press):
    """Maximum Doppler and Lorentz HWHM (cm-1) (broadening.py:431-498)."""
    dmax = (np.sqrt(2.0 * np.log(2.0) * pc.k * max_temp / (min_mass * pc.amu))
            * max_wn / pc.c)
    max_diam = _H2_RADIUS + max_rad
    lmax = (np.sqrt(2.0 / (np.pi * pc.k * min_temp * pc.amu))
            * max_press * pc.bar * max_diam**2.0 / pc.c
            * np.sqrt(1.0 / min_mass + 1.0 / _H2_MASS))
    return dmax, lmax
