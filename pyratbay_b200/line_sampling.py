"""Line-sampled cross-section tables (the consumer of the .npz tables), with the surface of
pyratbay/opacity/line_sampling.py:19-498 and everything numerical on the GPU:

  * the table lives in HBM; it comes from .npz files (uploaded once) or straight from device
    memory (`tables=`: e.g. the `ex.etable_dev` that extinction.compute_opacity leaves there,
    no .npz round trip);
  * the p/T re-gridding of tools/tools.py:1026-1107 is a kernel (pb200_regrid_table_dev);
  * the temperature interpolation of _extcoeff.c:367-472 goes through a table handle
    (pb200_table_interp): no per-call allocation, lock or synchronisation; with
    `device_out=True` the result stays in a persistent device buffer and the call returns once
    the kernel is queued.
"""
import os

import numpy as np

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


def _needs_resampling(tabulated, requested):
    """tools.py:1063-1076: a different length or any sample more than 1 % off."""
    if requested is None:
        return False
    requested = np.asarray(requested, np.double)
    return len(tabulated) != len(requested) or bool(np.any(np.abs(1.0 - tabulated / requested) > 0.01))


def _brackets(nodes, x):
    """Piecewise-linear brackets of `x` on the increasing `nodes` with edge values outside
    (interp1d kind='slinear', fill_value=(first, last)): (lo, hi, weight of hi)."""
    nodes = np.asarray(nodes, np.double)
    x = np.atleast_1d(np.asarray(x, np.double))
    n = len(nodes)
    lo = np.clip(np.searchsorted(nodes, x, side='right') - 1, 0, max(n - 2, 0))
    hi = np.minimum(lo + 1, n - 1)
    with np.errstate(divide='ignore', invalid='ignore'):
        f = np.where(hi > lo, (x - nodes[lo]) / (nodes[hi] - nodes[lo]), 0.0)
    below, above = x <= nodes[0], x >= nodes[-1]
    lo = np.where(below, 0, np.where(above, n - 1, lo))
    hi = np.where(below, 0, np.where(above, n - 1, hi))
    f = np.where(below | above, 0.0, f)
    exact = np.isin(x, nodes) & ~below & ~above          # a node: that row itself
    node_at = np.searchsorted(nodes, x)
    lo = np.where(exact, node_at, lo)
    hi = np.where(exact, node_at, hi)
    f = np.where(exact, 0.0, f)
    return lo.astype(np.int32), hi.astype(np.int32), f


def _identity(n):
    idx = np.arange(n, dtype=np.int32)
    return idx, idx, np.zeros(n)


class _Source:
    """One tabulated cross-section set: grids on the host, opacity on the device."""

    def __init__(self, name, species, temp, press, wn, opacity, device):
        import torch
        self.name = name
        self.species = species
        self.temp = np.asarray(temp, np.double)
        self.press = np.asarray(press, np.double)
        self.wn = np.asarray(wn, np.double)
        dev = torch.device('cuda', device)
        if isinstance(opacity, torch.Tensor):
            opacity = opacity.to(device=dev, dtype=torch.float64)
        else:
            opacity = torch.from_numpy(np.ascontiguousarray(opacity, np.float64)).to(dev)
        if opacity.dim() == 4:                      # pyratbay 2.0beta layout
            opacity = opacity[0]
        self.opacity = opacity.contiguous()
        if tuple(self.opacity.shape) != (len(self.temp), len(self.press), len(self.wn)):
            raise ValueError(f"Table '{name}' has shape {tuple(self.opacity.shape)}, expected "
                             f"[{len(self.temp)}, {len(self.press)}, {len(self.wn)}]")

    @classmethod
    def from_file(cls, cs_file, device):
        _units, species, temp, press, wn, opacity = io.read_opacity(cs_file, extract='all')
        return cls(cs_file, species, temp, press, wn, opacity, device)

    @classmethod
    def from_mapping(cls, table, device):
        get = table.get if isinstance(table, dict) else lambda k, d=None: getattr(table, k, d)
        species = get('species')
        if not isinstance(species, str):
            species = str(np.atleast_1d(species)[0])
        opacity = get('opacity')
        if opacity is None:
            opacity = get('etable_dev') if get('etable_dev') is not None else get('etable')
        temp = get('temperature') if get('temperature') is not None else get('temp')
        press = get('pressure') if get('pressure') is not None else get('press')
        wn = get('wavenumber') if get('wavenumber') is not None else get('wn')
        return cls(get('name') or f'<device table {species}>', species, temp, press, wn, opacity,
                   device)


def regrid_on_device(src, temperature, pressure, mask=None, wl_thinning=1, out=None,
                     accumulate=False, device=0):
    """tools/tools.py:1026-1107 on the device: `src` (a _Source) re-gridded in log-opacity over
    log p, then T; the table itself where nothing needs resampling.  Returns (or adds into)
    a device tensor [ntemp_out, nlayers_out, nwave_out]."""
    import torch
    if mask is None:
        mask = np.ones(len(src.wn), bool)
    wave_idx = np.where(mask)[0][::wl_thinning].astype(np.int32)
    resample_p = _needs_resampling(src.press, pressure)
    resample_t = _needs_resampling(src.temp, temperature)
    t_br = _brackets(src.temp, temperature) if resample_t else _identity(len(src.temp))
    p_br = (_brackets(np.log(src.press), np.log(np.asarray(pressure, np.double)))
            if resample_p else _identity(len(src.press)))
    shape = (len(t_br[0]), len(p_br[0]), len(wave_idx))
    if out is None:
        out = torch.empty(shape, dtype=torch.float64, device=src.opacity.device)
        accumulate = False
    elif tuple(out.shape) != shape:
        raise ValueError(f"regrid_on_device: destination has shape {tuple(out.shape)}, need {shape}")
    stream = torch.cuda.current_stream(src.opacity.device).cuda_stream
    eng.regrid_table_device(src.opacity.data_ptr(), src.opacity.shape, t_br, p_br, wave_idx,
                            out.data_ptr(), take_log=resample_p or resample_t,
                            accumulate=accumulate, device=device, stream=stream)
    return out


def interpolate_opacity(cs_file, temperature=None, pressure=None, mask=None, wl_thinning=1,
                        device=0):
    """Same contract as pyratbay.tools.interpolate_opacity (tools/tools.py:1026-1107): the
    cross sections of `cs_file` over the requested temperature (K) and pressure (bar) arrays,
    as a host array; computed on the GPU."""
    src = _Source.from_file(cs_file, device)
    return regrid_on_device(src, temperature, pressure, mask, wl_thinning, device=device).cpu().numpy()


def _parse_isotope_ratios(isotope_ratios):
    """Lines '<file key> <label> <log10 ratio | fill_a_b>' (line_sampling.py:142-156) ->
    {file key: (label, value string)}, insertion ordered."""
    entries = {}
    if isotope_ratios is None:
        return entries
    for line in isotope_ratios.strip().split('\n'):
        key, label, value = line.split()
        entries[key] = ('iso_' + label, value)
    return entries


class Line_Sample:
    """Line-by-line sampled opacities: cs_table [nspec, ntemp, nlayers, nwave] in HBM.

    cs_files : .npz cross-section file(s), as in the reference;
    tables   : instead of / in addition to files, tables already in memory: dicts or objects
               with species, temperature (or temp), pressure (press), wavenumber (wn) and
               opacity (or etable_dev / etable) [ntemp, nlayers, nwave], on the device or host,
               e.g. `Line_Sample(tables=[pyrat.ex])` right after compute_opacity.
    """

    def __init__(self, cs_files=None, *, tables=None, pressure=None, temperature=None,
                 min_wl=None, max_wl=None, min_wn=None, max_wn=None, isotope_ratios=None,
                 wl_thinning=1, device=0, log=None):
        import torch
        self.name = 'line sampling'
        if cs_files is None:
            cs_files = []
        elif isinstance(cs_files, str):
            cs_files = [cs_files]
        self.cs_files = list(cs_files)
        missing = [f for f in self.cs_files if not os.path.isfile(f)]
        if missing:
            raise ValueError(f'Missing opacity files: {missing}')
        if not self.cs_files and not tables:
            raise ValueError('Line_Sample needs cross-section files or tables')
        self.device = device
        sources = [_Source.from_file(f, device) for f in self.cs_files]
        sources += [_Source.from_mapping(t, device) for t in (tables or [])]
        self.cs_files += [s.name for s in sources[len(self.cs_files):]]

        first = sources[0]
        self.temp = first.temp if temperature is None else np.asarray(temperature, np.double)
        self.ntemp = len(self.temp)
        self.press = first.press if pressure is None else np.asarray(pressure, np.double)
        self.nlayers = len(self.press)
        if min_wn is not None and max_wl is not None:
            raise ValueError('Either define min_wn or max_wl, not both')
        if max_wn is not None and min_wl is not None:
            raise ValueError('Either define min_wl or max_wn, not both')
        if min_wn is None:
            min_wn = 0.0 if max_wl is None else 1.0 / (max_wl * pc.um)
        if max_wn is None:
            max_wn = np.inf if min_wl is None else 1.0 / (min_wl * pc.um)
        self.wn = first.wn[wn_mask(first.wn, min_wn, max_wn)][::wl_thinning]
        self.nwave = len(self.wn)

        # Which table feeds which row: a species, optionally split into isotopologues whose
        # label comes from the first isotope key found in the table's name.
        iso_entries = _parse_isotope_ratios(isotope_ratios)
        rows = {}                                  # (species, isotope label) -> row
        row_of_source, masks = [], []
        for src in sources:
            mask = wn_mask(src.wn, min_wn, max_wn)
            wn_src = src.wn[mask][::wl_thinning]
            if len(wn_src) != self.nwave or np.any(np.abs(1.0 - wn_src / self.wn) > 0.01):
                raise ValueError(
                    f"Wavenumber array of cross-section file '{src.name}' "
                    "does not match with previous arrays")
            check_pressure_boundaries(self.press, src.press)
            matches = [label for key, (label, _v) in iso_entries.items() if key in src.name]
            if len(matches) > 1:
                raise ValueError(f'Multiple isotope labels match {repr(src.name)}')
            tag = (src.species, matches[0] if matches else '')
            row_of_source.append(rows.setdefault(tag, len(rows)))
            masks.append(mask)
        self.species = np.array([species for species, _iso in rows])
        self.isotopes = [iso for _species, iso in rows]
        self.nspec = len(rows)

        # Isotopic ratios (line_sampling.py:198-229): a numeric value is a free parameter
        # (log10 of the ratio), 'fill_a_b' makes the ratio the complement of isotopes a and b.
        value_of = {label: value for label, value in iso_entries.values()}
        self.iso_ratios = np.ones(self.nspec, float)
        self.iso_fill = [None] * self.nspec
        self._iso_free, self.pnames, pars = [], [], []
        for row, label in enumerate(self.isotopes):
            value = value_of.get(label)
            if value is None:
                continue
            if value.startswith('fill_'):
                fillers = ['iso_' + name for name in value[5:].split('_')]
                if any(name not in self.isotopes for name in fillers):
                    raise ValueError('Invalid filler')
                self.iso_fill[row] = [self.isotopes.index(name) for name in fillers]
            else:
                self.iso_ratios[row] = 10.0 ** float(value)
                self._iso_free.append(row)
                self.pnames.append(label)
                pars.append(value)
        self._update_iso_ratios()
        self.pars = np.array(pars, float)
        self.npars = len(self.pars)
        self.texnames = list(self.pnames)

        # The table, re-gridded on the device and summed per row (line_sampling.py:243-250)
        dev = torch.device('cuda', device)
        self._dev_table = torch.zeros((self.nspec, self.ntemp, self.nlayers, self.nwave),
                                      dtype=torch.float64, device=dev)
        for src, row, mask in zip(sources, row_of_source, masks):
            regrid_on_device(src, self.temp, self.press, mask, wl_thinning,
                             out=self._dev_table[row], accumulate=True, device=device)
        self.tmin = np.amin(self.temp)
        self.tmax = np.amax(self.temp)
        self._cs_table = None
        self._handle = None
        self._out = {}

    @property
    def cs_table(self):
        """Host copy of the table (fetched on first use; the computations read the device one)."""
        if self._cs_table is None:
            self._cs_table = self._dev_table.cpu().numpy()
        return self._cs_table

    @property
    def cs_table_device(self):
        return self._dev_table

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

    def _layers(self, layer):
        if layer is None:
            return 0, self.nlayers
        if np.isscalar(layer):
            return layer, layer + 1
        if len(layer) == 2:
            return layer[0], layer[1]
        raise ValueError('Invalid layer input')

    def _interp(self, temperature, density, layer, per_mol, device_out):
        import torch
        if np.amax(temperature) > self.tmax or np.amin(temperature) < self.tmin:
            raise ValueError('Temperatures are out of line-sample bounds')
        lay1, lay2 = self._layers(layer)
        if self._handle is None:
            self._handle = eng.DeviceTable(self._dev_table, self.temp, self.device)
        out = self._out.get(per_mol)
        if out is None:
            shape = (self.nspec, self.nlayers, self.nwave) if per_mol else (self.nlayers, self.nwave)
            out = self._out[per_mol] = torch.empty(shape, dtype=torch.float64,
                                                   device=self._dev_table.device)
        stream = torch.cuda.current_stream(self._dev_table.device).cuda_stream
        self._handle.interp(temperature, density, lay1, lay2, per_mol, out.data_ptr(),
                            overwrite=True, stream=stream, sync=False)
        if not device_out:
            out = out.cpu().numpy()
        if np.isscalar(layer):
            out = out[:, layer] if per_mol else out[layer]
        return out

    def calc_cross_section(self, temperature, layer=None, per_mol=False, pars=None,
                           device_out=False):
        """Cross sections (cm2 molec-1) at the given layer temperatures
        (line_sampling.py:317-391).  device_out=True: a view of the persistent device buffer
        (valid until the next call), queued on torch's current stream, no synchronisation."""
        if pars is not None:
            self._update_iso_ratios(pars)
        density = np.ones((self.nlayers, self.nspec)) * self.iso_ratios
        return self._interp(np.asarray(temperature, np.double), density, layer, per_mol,
                            device_out)

    def calc_extinction_coefficient(self, temperature, density, layer=None, per_mol=False,
                                    pars=None, device_out=False):
        """Extinction coefficient (cm-1) (line_sampling.py:394-463); density [nlayers, nspec]."""
        if pars is not None:
            self._update_iso_ratios(pars)
        density = np.asarray(density, np.double) * self.iso_ratios
        return self._interp(np.asarray(temperature, np.double), density, layer, per_mol,
                            device_out)

    def __str__(self):
        """Same text as the reference's Line_Sample.__str__ (line_sampling.py:466-498)."""
        from .tools import Formatted_Write
        fw = Formatted_Write()
        fw.write(f"Line-sampling cross-section files (cs_files):\n{self.cs_files}")
        fw.write(f'Number of species (nspec): {self.nspec}')
        fw.write(f'Number of temperature samples (ntemp): {self.ntemp}')
        fw.write(f'Number of pressure layers (nlayers): {self.nlayers}')
        fw.write(f'Number of wavenumber samples (nwave): {self.nwave}')
        fw.write('\nMinimum and maximum temperatures (tmin, tmax) in K: '
                 f'[{self.tmin:.1f}, {self.tmax:.1f}]')
        fw.write('Minimum and maximum pressures in bar: '
                 f'[{np.amin(self.press):.3e}, {np.amax(self.press):.3e}]')
        fw.write('Minimum and maximum wavelengths in um: '
                 f'[{np.amin(self.get_wl()):.3f}, {np.amax(self.get_wl()):.3f}]')
        fw.write(f'\nLine-sample species (species): {self.species}')
        fw.write(f'Temperature array (temps, K):\n{self.temp}')
        with np.printoptions(precision=3):
            fw.write(f'Pressure layers (pressure, bar):\n{self.press}')
            fw.write(f'Wavenumber array (wn, cm-1):\n  {self.wn}')
        fw.write('The tabulated cross sections (cs_table, cm2 molecule-1) are an array\nof '
                 f'dimensions [nspec,ntemp,nlayers,nwave] and shape {self.cs_table.shape}')
        return fw.text
