"""Spectral grids feeding the line-by-line engine.

Restates the grid part of the reference's Spectrum.__init__
(pyratbay/pyrat/spectrum.py:181-228): output grid `wn`, fine grid `own` with step
wnstep/wnosamp, and the divisors of wnosamp used for dynamic sampling.
"""
import numpy as np

from . import constants as pc

# highly composite numbers (spectrum.py:181-185)
_HCN = np.array([
    1, 2, 4, 6, 12, 24, 36, 48, 60, 120, 180, 240, 360, 720, 840,
    1260, 1680, 2160, 2520, 5040, 7560, 10080, 15120, 20160, 25200,
    27720, 45360, 50400, 55440, 83160, 110880, 221760, 277200,
])


def divisors(number):
    """All integer divisors of `number` (tools.divisors, tools/tools.py:314-323)."""
    number = int(number)
    # candidates 1 .. number/2 (for number == 1 this yields [1, 1], as in the reference)
    upper = int(np.ceil(number / 2 + 1))
    divs = [i for i in range(1, upper) if number % i == 0]
    divs.append(number)
    return np.asarray(divs, int)


def constant_resolution_spectrum(wave_min, wave_max, resolution):
    """Constant resolving-power sampling (spectrum/spec_tools.py:461-504)."""
    f = 0.5 / resolution
    g = (1.0 + f) / (1.0 - f)
    nwave = int(np.ceil(-np.log(wave_min / wave_max) / np.log(g)))
    return wave_min * g**np.arange(nwave)


class Spectrum:
    """Wavenumber sampling (the attributes of pyrat.spec that the hot path reads)."""

    def __init__(self, wnlow=None, wnhigh=None, wl_low=None, wl_high=None, wnstep=None,
                 wnosamp=None, resolution=None, wlstep=None, log=None):
        # Boundaries (spectrum.py:77-118); wavelengths in cm.
        if wnlow is None and wl_high is None:
            raise ValueError('Undefined low wavenumber boundary.  Either set wnlow or wl_high')
        if wnhigh is None and wl_low is None:
            raise ValueError('Undefined high wavenumber boundary. Either set wnhigh or wl_low')
        if wnlow is not None:
            self.wnlow = float(wnlow)
            self.wl_high = 1.0 / self.wnlow
        else:
            self.wl_high = float(wl_high)
            self.wnlow = 1.0 / self.wl_high
        if wnhigh is not None:
            self.wnhigh = float(wnhigh)
            self.wl_low = 1.0 / self.wnhigh
        else:
            self.wl_low = float(wl_low)
            self.wnhigh = 1.0 / self.wl_low
        if self.wnlow > self.wnhigh:
            raise ValueError(
                f'Wavenumber low boundary ({self.wnlow:.1f} cm-1) must be '
                f'larger than the high boundary ({self.wnhigh:.1f} cm-1)')
        if wnstep is None and wlstep is None and resolution is None:
            raise ValueError(
                'Undefined spectral sampling rate, either set resolution, wnstep, or wlstep')

        # Sampling (spectrum.py:187-217)
        if wnstep is not None:
            self.wnstep = float(wnstep)
        if wnosamp is None:
            if wnstep is None:
                self.wnstep = 1.0
            self.wnosamp = int(_HCN[self.wnstep / _HCN <= 0.0004][0])
        else:
            self.wnosamp = int(wnosamp)
        self.resolution = resolution
        self.wlstep = wlstep
        if resolution is not None:
            self.wn = constant_resolution_spectrum(self.wnlow, self.wnhigh, resolution)
            self.wlstep = None
        elif wlstep is not None:
            wl = np.arange(self.wl_low, self.wl_high, wlstep)
            self.wn = 1.0 / np.flip(wl)
            self.wnlow = self.wn[0]
            self.resolution = None
        else:
            nwave = int((self.wnhigh - self.wnlow) / self.wnstep) + 1
            self.wn = self.wnlow + np.arange(nwave) * self.wnstep
        self.nwave = len(self.wn)

        # Fine-sampled grid (spectrum.py:219-228)
        self.ownstep = self.wnstep / self.wnosamp
        self.onwave = int(np.ceil((self.wn[-1] - self.wnlow) / self.ownstep)) + 1
        self.own = self.wnlow + np.arange(self.onwave) * self.ownstep
        self.odivisors = divisors(self.wnosamp)
        if log is not None and self.wn[-1] != self.wnhigh:
            log.warning(
                f'Final wavenumber modified from {self.wnhigh:.4f} cm-1 (input)'
                f'\n                            to {self.wn[-1]:.4f} cm-1')

    @property
    def interpolate(self):
        """True when the output grid needs 2-point interpolation (extinction.py:163)."""
        return self.resolution is not None or self.wlstep is not None

    @property
    def wl(self):
        return 1.0 / (self.wn * pc.um)
