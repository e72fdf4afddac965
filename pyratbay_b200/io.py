"""Cross-section table files (.npz), the layout of pyratbay/io/io.py:570-694."""
import numpy as np

from . import constants as pc

_UNITS = {
    'temperature': 'K',
    'pressure': 'bar',
    'wavenumber': 'cm-1',
    'cross section': 'cm2 molecule-1',
}


def write_opacity(ofile, species, temp, press, wn, opacity):
    """Write an opacity table; keys species, temperature, pressure, wavenumber,
    opacity[ntemp, nlayers, nwave], units (io.py:570-606)."""
    if not isinstance(species, str):
        raise ValueError("'species' input must be a string")
    np.savez(
        ofile,
        species=[species],
        temperature=temp,
        pressure=press,
        wavenumber=wn,
        opacity=opacity,
        units=dict(_UNITS),
    )


def read_opacity(ofile, extract='all'):
    """Read an opacity table (io.py:609-694).  extract in {'arrays','opacity','all'}."""
    if ofile.endswith('petitRADTRANS.h5'):
        raise NotImplementedError("petitRADTRANS tables need h5py (out of scope here)")
    with np.load(ofile, allow_pickle=True) as f:
        if len(f['species']) > 1:
            raise ValueError('Opacity files must contain a single species')
        species = str(f['species'][0])
        temp = f['temperature']
        press = f['pressure']
        wn = f['wavenumber']
        if extract in ['opacity', 'all']:
            opacity = f['opacity']
            if np.ndim(opacity) == 4:  # pyratbay 2.0beta layout
                opacity = opacity[0]
        units = np.ndarray.item(f['units']) if 'units' in f else None
    if units is None:  # pyratbay < 2.0 stored barye
        press = press / pc.bar
        units = dict(_UNITS)
    if extract == 'opacity':
        return opacity
    if extract == 'arrays':
        return species, temp, press, wn
    if extract == 'all':
        return units, species, temp, press, wn, opacity
    raise ValueError(f"Invalid extract mode '{extract}'")
