"""Grid of Voigt profiles, mirroring pyratbay/pyrat/voigt.py:11-194.

Sizing (Doppler/Lorentz grids, half-sizes, dlratio skipping) is host NumPy exactly like
the reference; the profile table itself is computed on the GPU (csrc/voigt.cu) and stays
there.  `Voigt.profile` is fetched lazily for callers that want the host array.
"""
import numpy as np

from . import broadening
from .engine import Engine


def voigt_sizes(lorentz, doppler, extent, cutoff, ownstep, onwave, dlratio):
    """Half-sizes [nlor, ndop] with 0 marking profiles not to be computed
    (pyrat/voigt.py:105-131)."""
    lorentz = np.asarray(lorentz, np.double)
    doppler = np.asarray(doppler, np.double)
    size = np.zeros((len(lorentz), len(doppler)), int)
    for i in range(len(lorentz)):
        pwidth = extent * (
            0.5346 * lorentz[i] + np.sqrt(0.2166 * lorentz[i]**2 + doppler**2))
        if cutoff > 0:
            pwidth = np.minimum(pwidth, cutoff)
        psize = 1 + 2 * np.asarray(pwidth / ownstep + 0.5, int)
        psize = np.clip(psize, 3, 1 + 2 * onwave)
        skip = doppler / lorentz[i] < dlratio
        skip[0] = False
        psize[skip] = 0
        size[i] = psize // 2
    return size


class Voigt:
    """Voigt-profile grid with the reference's attribute names: profile, size, index,
    lorentz, doppler, extent, cutoff, dlratio, dmin/dmax/ndop, lmin/lmax/nlor."""

    def __init__(self, spec, atm, iso_atm_index, engine, extent=300.0, cutoff=25.0,
                 dlratio=0.1, ndop=50, nlor=100, dmin=None, dmax=None, lmin=None,
                 lmax=None, tmin=None, tmax=None, log=None):
        self.extent = extent
        self.cutoff = cutoff
        self.dlratio = dlratio
        self._engine = engine
        self._profile = None

        # Width boundaries from the atmosphere (pyrat/voigt.py:29-54)
        min_wn, max_wn = np.amin(spec.wn), np.amax(spec.wn)
        min_pressure, max_pressure = np.amin(atm.press), np.amax(atm.press)
        min_temp = 100.0 if tmin is None else tmin
        max_temp = 3000.0 if tmax is None else tmax
        mol_indices = np.unique(iso_atm_index)
        min_mass = np.amin(atm.mol_mass[mol_indices])
        max_mass = np.amax(atm.mol_mass[mol_indices])
        min_rad = np.amin(atm.mol_radius[mol_indices])
        max_rad = np.amax(atm.mol_radius[mol_indices])
        est_dmin, est_lmin = broadening.min_widths(
            min_temp, max_temp, min_wn, max_mass, min_rad, min_pressure)
        est_dmax, est_lmax = broadening.max_widths(
            min_temp, max_temp, max_wn, min_mass, max_rad, max_pressure)

        self.dmin = est_dmin if dmin is None else dmin
        self.dmax = est_dmax if dmax is None else dmax
        if self.dmax <= self.dmin:
            raise ValueError(
                f'Voigt dmax ({self.dmax:.4e} cm-1) must be > dmin ({self.dmin:.4e} cm-1)')
        self.ndop = ndop
        self.doppler = np.logspace(np.log10(self.dmin), np.log10(self.dmax), self.ndop)

        self.lmin = est_lmin if lmin is None else lmin
        self.lmax = est_lmax if lmax is None else lmax
        if self.lmax <= self.lmin:
            raise ValueError(
                f'Voigt lmax ({self.lmax:.4e} cm-1) must be > lmin ({self.lmin:.4e} cm-1)')
        self.nlor = nlor
        self.lorentz = np.logspace(np.log10(self.lmin), np.log10(self.lmax), self.nlor)

        self.size = voigt_sizes(self.lorentz, self.doppler, self.extent, self.cutoff,
                                spec.ownstep, spec.onwave, self.dlratio)
        self.index = np.zeros((self.nlor, self.ndop), int)
        if log is not None:
            log.msg('Calculating Voigt profiles with max extent: '
                    f'{self.extent:.1f} HWHM.', indent=2)
        # Table on the device (replaces vp.grid, pyrat/voigt.py:145-149).  With engine=None
        # only the host-side sizing is done (size still holds 0 for skipped profiles).
        self.size = np.ascontiguousarray(self.size, np.int64)
        self.index = np.ascontiguousarray(self.index, np.int64)
        self.profile_len = int(np.sum(2 * self.size + 1))  # pyrat/voigt.py:142
        if engine is not None:
            engine.build_voigt(self.lorentz, self.doppler, spec.ownstep, self.size,
                               self.index, self.cutoff)

    @property
    def profile(self):
        """Host copy of the concatenated profiles [sum(2*size+1)] (fetched on first use)."""
        if self._profile is None:
            if self._engine is None:
                from ._lib import PB200Error
                raise PB200Error("Voigt built without an engine: no profile table "
                                 "(there is no CPU fallback)")
            self._profile = self._engine.get_profile()
        return self._profile

    def __str__(self):
        """Same text as the reference's Voigt.__str__ (pyrat/voigt.py:154-194; pinned by the
        reference's tests/test_str.py:338-366)."""
        from .tools import Formatted_Write
        fw = Formatted_Write(fmt={'float': '{:.3e}'.format}, edge=3)
        fw.write('Voigt-profile information:')
        fw.write('\nNumber of Doppler-width samples (ndop): {:d}', self.ndop)
        fw.write('Number of Lorentz-width samples (nlor): {:d}', self.nlor)
        fw.write('Doppler HWHM (doppler, cm-1):\n    {}', self.doppler)
        fw.write('Lorentz HWMH (lorentz, cm-1):\n    {}', self.lorentz)
        fw.write(f'Doppler--Lorentz ratio threshold (dlratio): {self.dlratio:.3e}')
        fw.write(f"\nVoigt-profiles extent (extent, in HWHMs): {self.extent:.1f}")
        fw.write(f"Voigt-profiles cutoff extent (cutoff in cm-1): {self.cutoff:.1f}")
        fw.write('Voigt-profile half-sizes (size) of shape [nlor, ndop]:\n{}', self.size,
                 edge=2)
        fw.write('Voigt-profile indices (index) of shape [nlor, ndop]:\n{}', self.index,
                 edge=2)
        index, size = self.index[0, 0], 2 * self.size[0, 0] + 1
        fw.write('\nVoigt profiles:\n  profile[ 0, 0]: {}', self.profile[index:index+size],
                 fmt={'float': '{:.5e}'.format}, edge=2)
        index = self.index[self.nlor-1, self.ndop-1]
        size = 2 * self.size[self.nlor-1, self.ndop-1] + 1
        fw.write('  ...\n  profile[{:2d},{:2d}]: {}', self.nlor-1, self.ndop-1,
                 self.profile[index:index+size], fmt={'float': '{:.5e}'.format}, edge=2)
        return fw.text
