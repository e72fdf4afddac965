"""Line-sampled cross-section tables (consumer of the .npz tables), mirroring the core of
pyratbay/opacity/line_sampling.py:19-463 with the temperature interpolation on the GPU
(csrc/lbl_kernels.cu: interp_ec_kernel, replacing _extcoeff.c:367-472).

The table is uploaded once and stays resident in HBM; each call moves only the layer
temperatures/densities in and the [nlayers, nwave] (or per-species) spectrum out.
"""
import os

import numpy as np
import scipy.interpolate as sip

from . import constants as pc
from . import io
from . import engine as eng


def wn_mask(wn, wn_min, wn_max, tol=1.0e-8):
    """Mask of wn within [wn_min, wn_max] with an edge tolerance
    (spectrum/spec_tools.py:778-815)."""
    mask = (wn >= wn_min) & (wn <= wn_max)
    if np.sum(mask) < 2:
        min_dwn = max_dwn = 0
    else:
        min_dwn = np.abs(np.ediff1d(wn[mask][0:2]))
        max_dwn = np.abs(np.ediff1d(wn[mask][-2:]))
    return (wn >= wn_min - min_dwn * tol) & (wn <= wn_max + max_dwn * tol)


def check_pressure_boundaries(press, tabulated_press):
    """line_sampling.py:501-511."""
    if np.amax(press) / np.amax(tabulated_press) - 1 > 1e-3:
        raise ValueError('Pressure profile extends beyond the maximum tabulated pressure')


def interpolate_opacity(cs_file, temperature=None, pressure=None, mask=None, wl_thinning=1):
    """Re-grid a table in log-opacity over pressure and temperature
    (tools/tools.py:1026-1107; slinear in log p and T, floor -230)."""
    _, temp, press, wn = io.read_opacity(cs_file, extract='arrays')
    logp_table = np.log(press)
    if mask is None:
        mask = np.ones(len(wn), bool)
    resample_pressure = (
        pressure is not None and
        (len(press) != len(pressure) or np.any(np.abs(1.0 - press / pressure) > 0.01)))
    resample_temperature = (
        temperature is not None and
        (len(temp) != len(temperature) or np.any(np.abs(1.0 - temp / temperature) > 0.01)))
    cross_section = io.read_opacity(cs_file, extract='opacity')[:, :, mask]
    cross_section = cross_section[:, :, ::wl_thinning]
    if not resample_pressure and not resample_temperature:
        return cross_section
    with np.errstate(divide='ignore'):
        log_cs = np.log(cross_section)
    log_cs[~np.isfinite(log_cs)] = -230.0
    if resample_pressure:
        interp = sip.interp1d(logp_table, log_cs, axis=1, kind='slinear', bounds_error=False,
                              fill_value=(log_cs[:, 0], log_cs[:, -1]))
        log_cs = interp(np.log(pressure))
    if resample_temperature:
        interp = sip.interp1d(temp, log_cs, axis=0, kind='slinear', bounds_error=False,
                              fill_value=(log_cs[0], log_cs[-1]))
        log_cs = interp(temperature)
    return np.exp(log_cs)


class Line_Sample:
    """Line-by-line sampled opacities: cs_table [nspec, ntemp, nlayers, nwave]."""

    def __init__(self, cs_files, *, pressure=None, temperature=None, min_wl=None,
                 max_wl=None, min_wn=None, max_wn=None, isotope_ratios=None, wl_thinning=1,
                 device=0, log=None):
        self.name = 'line sampling'
        if isinstance(cs_files, str):
            cs_files = [cs_files]
        self.cs_files = list(cs_files)
        missing = [f for f in self.cs_files if not os.path.isfile(f)]
        if missing:
            raise ValueError(f'Missing opacity files: {missing}')
        self.device = device

        _, temp, press, wn = io.read_opacity(self.cs_files[0], extract='arrays')
        self.temp = temp if temperature is None else np.asarray(temperature, np.double)
        self.ntemp = len(self.temp)
        self.press = press if pressure is None else np.asarray(pressure, np.double)
        self.nlayers = len(self.press)
        if min_wn is not None and max_wl is not None:
            raise ValueError('Either define min_wn or max_wl, not both')
        if max_wn is not None and min_wl is not None:
            raise ValueError('Either define min_wl or max_wn, not both')
        if min_wn is None:
            min_wn = 0.0 if max_wl is None else 1.0 / (max_wl * pc.um)
        if max_wn is None:
            max_wn = np.inf if min_wl is None else 1.0 / (min_wl * pc.um)
        mask = wn_mask(wn, min_wn, max_wn)
        self.wn = wn[mask][::wl_thinning]
        self.nwave = len(self.wn)

        # Isotopic parameters: lines "<file key> <label> <log10 ratio | fill_a_b>"
        # (line_sampling.py:142-156)
        iso_keys, iso_labels, iso_ratios = [], [], []
        if isotope_ratios is not None:
            for iso_data in isotope_ratios.strip().split('\n'):
                ext_label, label, ratio = iso_data.split()
                iso_keys.append(ext_label)
                iso_labels.append('iso_' + label)
                iso_ratios.append(ratio)

        self.species, self.isotopes = [], []
        iso_species, species_per_file = [], []
        masks = []
        for cs_file in self.cs_files:
            species, _t, p, w = io.read_opacity(cs_file, extract='arrays')
            m = wn_mask(w, min_wn, max_wn)
            w = w[m][::wl_thinning]
            masks.append(m)
            if len(w) != self.nwave or np.any(np.abs(1.0 - w / self.wn) > 0.01):
                raise ValueError(
                    f"Wavenumber array of cross-section file '{cs_file}' "
                    "does not match with previous arrays")
            check_pressure_boundaries(self.press, p)
            iso = ''
            for i, key in enumerate(iso_keys):
                if key in cs_file and iso != '':
                    raise ValueError(f'Multiple isotope labels match {repr(cs_file)}')
                elif key in cs_file:
                    iso = iso_labels[i]
            species_per_file.append(species + iso)
            if species + iso not in iso_species:
                iso_species.append(species + iso)
                self.species.append(species)
                self.isotopes.append(iso)
        spec_indices = [iso_species.index(sp) for sp in species_per_file]
        self.species = np.array(self.species)
        self.nspec = len(self.species)

        # Isotopic ratios: free parameters and fillers (line_sampling.py:198-229)
        self.iso_ratios = np.ones(self.nspec, float)
        self.iso_fill = [None] * self.nspec
        self._iso_free = []
        self.pnames = []
        pars = []
        for i, iso in enumerate(self.isotopes):
            if iso == '':
                continue
            ratio = iso_ratios[iso_labels.index(iso)]
            if not ratio.startswith('fill_'):
                self.iso_ratios[i] = 10.0**float(ratio)
                self.pnames.append(iso)
                self._iso_free.append(i)
                pars.append(ratio)
                continue
            fillers = [f'iso_{filler}' for filler in ratio[5:].split('_')]
            for filler in fillers:
                if filler not in self.isotopes:
                    raise ValueError('Invalid filler')
            self.iso_fill[i] = [self.isotopes.index(filler) for filler in fillers]
        self._update_iso_ratios()
        self.pars = np.array(pars, float)
        self.npars = len(self.pars)
        self.texnames = list(self.pnames)

        self.cs_table = np.zeros((self.nspec, self.ntemp, self.nlayers, self.nwave))
        for i, cs_file in enumerate(self.cs_files):
            self.cs_table[spec_indices[i]] += interpolate_opacity(
                cs_file, self.temp, self.press, masks[i], wl_thinning)
        self.tmin = np.amin(self.temp)
        self.tmax = np.amax(self.temp)
        self._dev_table = None

    def _update_iso_ratios(self, pars=None):
        """Update the isotopic ratios, keeping the fillers complementary
        (line_sampling.py:278-293)."""
        if pars is not None:
            self.iso_ratios[self._iso_free] = 10.0**np.array(pars)
        for i, fillers in enumerate(self.iso_fill):
            if fillers is not None:
                self.iso_ratios[i] = 1.0 - np.sum(self.iso_ratios[fillers])

    def get_wl(self, units='um'):
        return 1.0 / (self.wn * pc.u(units))

    def _device_table(self):
        """Upload the table once; it stays in HBM for all later calls."""
        if self._dev_table is None:
            import torch
            self._dev_table = torch.from_numpy(self.cs_table).to(f'cuda:{self.device}')
        return self._dev_table

    def _layers(self, layer):
        if layer is None:
            return 0, self.nlayers
        if np.isscalar(layer):
            return layer, layer + 1
        if len(layer) == 2:
            return layer[0], layer[1]
        raise ValueError('Invalid layer input')

    def _interp(self, temperature, density, layer, per_mol):
        import torch
        if np.amax(temperature) > self.tmax or np.amin(temperature) < self.tmin:
            raise ValueError('Temperatures are out of line-sample bounds')
        lay1, lay2 = self._layers(layer)
        table = self._device_table()
        shape = (self.nspec, self.nlayers, self.nwave) if per_mol else (self.nlayers, self.nwave)
        out = torch.zeros(shape, dtype=torch.float64, device=table.device)
        eng.interp_ec_device(out.data_ptr(), table.data_ptr(), self.temp, temperature, density,
                             self.cs_table.shape, lay1, lay2, per_mol=per_mol,
                             device=self.device)
        out = out.cpu().numpy()
        if np.isscalar(layer):
            out = out[:, layer] if per_mol else out[layer]
        return out

    def calc_cross_section(self, temperature, layer=None, per_mol=False, pars=None):
        """Cross sections (cm2 molec-1) at the given layer temperatures
        (line_sampling.py:317-391)."""
        if pars is not None:
            self._update_iso_ratios(pars)
        density = np.ones((self.nlayers, self.nspec)) * self.iso_ratios
        return self._interp(np.asarray(temperature, np.double), density, layer, per_mol)

    def calc_extinction_coefficient(self, temperature, density, layer=None, per_mol=False,
                                    pars=None):
        """Extinction coefficient (cm-1) (line_sampling.py:394-463); density [nlayers, nspec]."""
        if pars is not None:
            self._update_iso_ratios(pars)
        density = np.asarray(density, np.double) * self.iso_ratios
        return self._interp(np.asarray(temperature, np.double), density, layer, per_mol)
