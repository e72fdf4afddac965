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


class OpacityWriter:
    """Write an opacity table file incrementally: same members and layout as write_opacity
    (np.savez, io.py:570-606), but the `opacity` array [ntemp, nlayers, nwave] is streamed row
    block by row block, so the file can be written while later rows are still being computed.

        with OpacityWriter(ofile, species, temp, press, wn) as w:
            w.write(rows)            # any number of consecutive [*, nwave] row blocks
    """

    def __init__(self, ofile, species, temp, press, wn):
        import zipfile
        if not isinstance(species, str):
            raise ValueError("'species' input must be a string")
        if not ofile.endswith('.npz'):
            ofile += '.npz'                      # as np.savez does
        self.shape = (len(temp), len(press), len(wn))
        self._left = int(np.prod(self.shape)) * 8
        self._zf = zipfile.ZipFile(ofile, 'w', zipfile.ZIP_STORED, allowZip64=True)
        for name, arr in (('species', [species]), ('temperature', temp), ('pressure', press),
                          ('wavenumber', wn)):
            self._member(name, np.asanyarray(arr))
        self._f = self._zf.open('opacity.npy', 'w', force_zip64=True)
        np.lib.format.write_array_header_1_0(
            self._f, {'descr': '<f8', 'fortran_order': False, 'shape': self.shape})

    def _member(self, name, arr):
        with self._zf.open(name + '.npy', 'w', force_zip64=True) as f:
            np.lib.format.write_array(f, arr, allow_pickle=True)

    def write(self, rows):
        rows = np.ascontiguousarray(rows, '<f8')
        self._left -= rows.nbytes
        if self._left < 0:
            raise ValueError('OpacityWriter: more rows than the table holds')
        self._f.write(memoryview(rows).cast('B'))

    def close(self):
        if self._zf is None:
            return
        self._f.close()
        self._member('units', np.asanyarray(dict(_UNITS)))
        self._zf.close()
        self._zf = None
        if self._left != 0:
            raise ValueError('OpacityWriter: the table was not written completely')

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        if exc_type is None:
            self.close()
        elif self._zf is not None:
            self._f.close()
            self._zf.close()
            self._zf = None
        return False


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
