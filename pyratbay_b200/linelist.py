"""Line-list database readers: HITRAN/HITEMP `.par`, ExoMol
`.trans`/`.states` and repack binary files -> (wn, gf, elow, iso_id) arrays.

Same classes, methods and results as pyratbay/opacity/linelist/{driver,hitran,exomol,repack}.py,
but the records are parsed column-wise with NumPy from one read of the file instead of one
seek + read + float() per record (hitran.py:173-181, exomol.py:166-170, repack.py:122-127), so
ExoMol-scale inputs (1e8-1e9 records) are ingested at memory speed.  Window selection keeps
the reference's binary-then-linear record search (driver.py:80-137) on the in-memory
wavenumber column, including its behaviour on repeated values.
"""
import bz2
import itertools
import json
import os
import re

import numpy as np

from . import constants as pc
from . import tli as ptli

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')

# gf = g_up * A21 * C1 / (8 pi c) / wn^2, equation (36) of Simeckova et al. (2006)
# (hitran.py:198, exomol.py:198); C1 = m_e c^2 / (e^2 pi) (astrophysical_constants.py:129).
import scipy.constants as _sc
_ME = _sc.m_e * 1e3
_E = 4.803205e-10
C1 = _ME * pc.c**2 / (_E**2 * _sc.pi)


def _isotope_table():
    with open(os.path.join(_DATA, 'isotopes_subset.json')) as f:
        return json.load(f)


def read_pf(pffile):
    """Partition-function file reader (io/io.py:740-810): returns (pf[niso, ntemp],
    isotopes, temp)."""
    if not os.path.isfile(pffile):
        raise ValueError(f"Partition-function file '{pffile}' does not exist.")
    with open(pffile, 'r') as f:
        lines = f.readlines()
    isotopes = None
    start = None
    it = iter(enumerate(lines))
    for i, line in it:
        line = line.strip()
        if line == '@DATA':
            start = i + 1
            break
        if line == '@ISOTOPES':
            isotopes = np.asarray(next(it)[1].split())
    if isotopes is None or start is None:
        raise ValueError(f"Partition-function file '{pffile}' has no @ISOTOPES/@DATA section.")
    data = np.array([row.split() for row in lines[start:] if row.strip()], np.double)
    return data[:, 1:].T.copy(), isotopes, data[:, 0].copy()


def get_exomol_mol(dbfile):
    """Molecule and isotope name from an ExoMol file name (tools/tools.py:846-901)."""
    atoms = os.path.split(dbfile)[1].split('_')[0].split('-')
    elements, isotope = [], ''
    for atom in atoms:
        match = re.match(r"([0-9]+)([a-z]+)([0-9]*)", atom, re.I)
        n = 1 if match.group(3) == '' else int(match.group(3))
        elements += n * [match.group(2)]
        isotope += match.group(1)[-1:] * n
    composition = [list(g[1]) for g in itertools.groupby(elements)]
    molecule = ''.join([c[0] + str(len(c)) * (len(c) > 1) for c in composition])
    if molecule == 'OCO':
        molecule = 'CO2'
    return molecule, isotope


class _NullLog:
    verb = 0

    def msg(self, *a, **k): pass
    head = debug = warning = msg

    def error(self, message, *a, **k):
        raise ValueError(message)


class Linelist:
    """Base class (linelist/driver.py:10-163)."""

    def __init__(self, dbfile, pffile, log=None):
        self.dbfile = dbfile
        self.pffile = pffile
        self.log = log if log is not None else _NullLog()

    def getpf(self, verbose=0):
        """(temp, pf[niso, ntemp], isotope names) (driver.py:17-62).  'tips' is served from
        the bundled TIPS-2021 tables (data/tips_subset.npz: H2O, CO2, CO, CH4, NH3, HCN); other
        molecules need a tabulated file."""
        if self.pffile == 'tips':
            with np.load(os.path.join(_DATA, 'tips_subset.npz')) as tips:
                if f'{self.molecule}_z' not in tips:
                    self.log.error(
                        f"pflist = tips: no bundled TIPS table for {self.molecule}; give a "
                        "partition-function file instead")
                return (tips[f'{self.molecule}_temp'].copy(), tips[f'{self.molecule}_z'].copy(),
                        [str(i) for i in tips[f'{self.molecule}_iso']])
        pf, iso, temp = read_pf(self.pffile)
        return temp, pf, iso.tolist()

    def get_iso(self, molname):
        """Isotope names (ExoMol convention), masses and ratios (driver.py:139-163)."""
        table = _isotope_table()['molecules']
        if molname not in table:
            self.log.error(f"No isotope data bundled for molecule '{molname}'")
        m = table[molname]
        return list(m['exomol_iso']), list(m['mass']), list(m['ratio'])

    @staticmethod
    def binsearch_array(wave_of, target, ilo, ihi, searchup=True):
        """Record index for `target` with the semantics of the reference's file search
        (driver.py:80-137), `wave_of(irec)` replacing its seek + read: a bisection that keeps
        wave(lo) <= target < wave(hi), then a walk over neighbouring records (up from `lo`
        while the next record is still below the target, or down from `hi` while the previous
        one is still above it) that stops at either end of the search range."""
        first, last = ilo, ihi
        lo, hi = ilo, ihi
        while hi - lo > 1:
            mid = (hi + lo) // 2
            if wave_of(mid) > target:
                hi = mid
            else:
                lo = mid
        step = 1 if searchup else -1
        pos = lo if searchup else hi
        while pos != first and pos != last:
            neighbour = wave_of(pos + step)
            if not (neighbour < target if searchup else neighbour > target):
                break
            pos += step
        return pos

    def _window(self, wn_all, iwn, fwn):
        """Record range [istart, istop] of the reference's dbread, or None (no overlap)."""
        nlines = len(wn_all)
        db_iwn, db_fwn = wn_all[0], wn_all[nlines - 1]
        if iwn > db_fwn or fwn < db_iwn:
            self.log.warning(
                f"Database ('{os.path.basename(self.dbfile)}') wavenumber "
                f"range ({db_iwn:.2f}--{db_fwn:.2f} cm-1) does not overlap with "
                f"the requested wavenumber range ({iwn:.2f}--{fwn:.2f} cm-1).")
            return None
        wave_of = wn_all.__getitem__
        istart = self.binsearch_array(wave_of, iwn, 0, nlines - 1, False)
        istop = self.binsearch_array(wave_of, fwn, istart, nlines - 1, True)
        self.log.msg(f'Process {self.name} database between records '
                     f'{istart:,d} and {istop:,d}.', indent=2)
        return istart, istop


def _fixed_width_records(path):
    """The file as a [nrecords, recsize] uint8 array (records = lines of equal length)."""
    raw = np.fromfile(path, dtype=np.uint8)
    first = np.flatnonzero(raw[:4096] == 10)
    if len(first) == 0:
        raise ValueError(f"'{path}': no end of line found in the first 4096 bytes")
    recsize = int(first[0]) + 1
    nrec = len(raw) // recsize
    return raw[:nrec * recsize].reshape(nrec, recsize), recsize


def _column_float(rec, lo, hi):
    """float() of the text in columns [lo, hi) of every record."""
    return np.ascontiguousarray(rec[:, lo:hi]).view(f'S{hi - lo}').ravel().astype(np.double)


def _first_three_fields(rec):
    """(int, int, float) columns of fixed-format text records (ExoMol .trans: '%12d %12d
    %10.4e ...').  The column boundaries are taken from the first record; when every record
    has blanks exactly there the fields are converted column-wise, otherwise (ragged input)
    record by record like the reference does."""
    first = rec[0].tobytes()
    spans, pos = [], 0
    for _ in range(3):
        while pos < len(first) and first[pos:pos + 1].isspace():
            pos += 1
        start = pos
        while pos < len(first) and not first[pos:pos + 1].isspace():
            pos += 1
        spans.append((start, pos))
    # field k occupies [previous field's end, its own end): right-aligned numbers
    edges = [0] + [e for _, e in spans]
    seps = [e for e in edges[1:] if e < rec.shape[1]]
    blank = np.isin(rec[:, seps], (32, 9, 10, 13)).all() if seps else True
    if blank and spans[2][1] > spans[2][0]:
        cols = [np.ascontiguousarray(rec[:, edges[k]:edges[k + 1]])
                .view(f'S{edges[k + 1] - edges[k]}').ravel() for k in range(3)]
        try:
            return cols[0].astype(np.int64), cols[1].astype(np.int64), cols[2].astype(np.double)
        except ValueError:
            pass  # a record with shifted columns: fall through to the generic split
    fields = np.array([r.split(None, 3)[:3]
                       for r in rec.view(f'S{rec.shape[1]}').ravel().tolist()])
    return (fields[:, 0].astype(np.int64), fields[:, 1].astype(np.int64),
            fields[:, 2].astype(np.double))


class Hitran(Linelist):
    """HITRAN/HITEMP 160-character `.par` reader (linelist/hitran.py)."""
    rec_iso, rec_wn, rec_strength, rec_A21, rec_air = 2, 3, 15, 25, 35
    rec_elow, rec_elow_end, rec_g2, rec_g2_end = 45, 55, 146, 153
    _iso_map = {c: i for i, c in enumerate('1234567890AB')}

    def __init__(self, dbfile, pffile, log=None):
        super().__init__(dbfile, pffile, log)
        if not os.path.isfile(self.dbfile):
            self.log.error(f"Input database file '{self.dbfile}' does not exist.")
        with open(self.dbfile, 'r') as data:
            mol_id = int(data.read(2))
        table = _isotope_table()
        if str(mol_id) not in table['hitran_mol_id']:
            self.log.error(f'No isotope data bundled for HITRAN molecule ID: {mol_id}')
        self.molecule = table['hitran_mol_id'][str(mol_id)]
        self.name = 'HITRAN ' + self.molecule
        iso_names, mass, ratio = self.get_iso(self.molecule)
        # HITRAN isotope order (hitran.py:55-62)
        isotopes = table['molecules'][self.molecule]['tips_order']
        isort = [iso_names.index(iso) for iso in isotopes]
        self.isotopes = list(isotopes)
        self.mass = np.array(mass)[isort]
        self.isoratio = np.array(ratio)[isort]

    def dbread(self, iwn, fwn, verb=0):
        rec, self.recsize = _fixed_width_records(self.dbfile)
        wn_all = _column_float(rec, self.rec_wn, self.rec_strength)
        window = self._window(wn_all, iwn, fwn)
        if window is None:
            return None
        istart, istop = window
        rec = rec[istart:istop + 1]
        wnumber = wn_all[istart:istop + 1]
        elow = _column_float(rec, self.rec_elow, self.rec_elow_end)
        a21 = _column_float(rec, self.rec_A21, self.rec_air)
        g2 = _column_float(rec, self.rec_g2, self.rec_g2_end)
        lut = np.full(256, -1, np.int64)
        for ch, idx in self._iso_map.items():
            lut[ord(ch)] = idx
        iso_id = lut[rec[:, self.rec_iso]]
        if np.any(iso_id < 0):
            bad = chr(int(rec[np.flatnonzero(iso_id < 0)[0], self.rec_iso]))
            raise KeyError(bad)
        gf = g2 * a21 * C1 / (8.0 * np.pi * pc.c) / wnumber**2.0
        good = np.where(elow > 0)   # unknown Elow, Rothman et al. (1996)  (hitran.py:201-202)
        return wnumber[good], gf[good], elow[good], iso_id[good]


class Exomol(Linelist):
    """ExoMol `.trans` + `.states` reader (linelist/exomol.py)."""

    def __init__(self, dbfile, pffile, log=None):
        super().__init__(dbfile, pffile, log)
        if not os.path.isfile(self.dbfile):
            self.log.error(f"Exomol file '{self.dbfile}' does not exist")
        sfile = self.dbfile.replace('trans', 'states')
        if sfile.count('__') == 2:
            suffix = sfile[sfile.rindex('__'):sfile.index('.')]
            sfile = sfile.replace(suffix, '')
        if sfile.count('__') == 2:
            sfile = sfile.replace(sfile[sfile.rindex('__'):sfile.index('.')], '')
        if os.path.isfile(sfile):
            with open(sfile, 'rb') as f:
                text = f.read()
        elif os.path.isfile(sfile + '.bz2'):
            with bz2.open(sfile + '.bz2', 'rb') as f:
                text = f.read()
        else:
            self.log.error(f"Exomol file '{sfile}' does not exist")
        # first three whitespace-separated fields of every line: id, energy, degeneracy
        rows = [line.split(None, 3)[:3] for line in text.splitlines() if line.strip()]
        cols = np.array(rows)
        state_id = cols[:, 0].astype(np.int64)
        state_e = cols[:, 1].astype(np.double)
        state_g = cols[:, 2].astype(np.int64)
        nstates = int(np.amax(state_id)) + 1      # in case of gaps (exomol.py:64-69)
        self.E = np.zeros(nstates, np.double)
        self.g = np.zeros(nstates, np.int64)
        self.E[state_id] = state_e
        self.g[state_id] = state_g
        self.molecule, self.iso = get_exomol_mol(dbfile)
        self.name = 'Exomol ' + self.molecule
        self.isotopes, self.mass, self.isoratio = self.get_iso(self.molecule)

    def dbread(self, iwn, fwn, verb=0):
        with open(self.dbfile, 'rb') as f:
            text = f.read()
        self.recsize = text.index(b'\n') + 1
        nlines = len(text) // self.recsize
        # upper id, lower id, A21: the first three fields of every (equal-length) record
        rec = np.frombuffer(text, np.uint8, nlines * self.recsize).reshape(nlines, self.recsize)
        up, lo, a21 = _first_three_fields(rec)
        wn_all = self.E[up] - self.E[lo]
        window = self._window(wn_all, iwn, fwn)
        if window is None:
            return None
        istart, istop = window
        sl = slice(istart, istop + 1)
        wnumber = wn_all[sl]
        gf = self.g[up[sl]] * a21[sl] * C1 / (8.0 * np.pi * pc.c) / wnumber**2.0
        elow = self.E[lo[sl]]
        iso_id = np.full(istop - istart + 1, self.isotopes.index(self.iso), np.int64)
        return wnumber, gf, elow, iso_id


class Repack(Linelist):
    """repack binary line lists, records 'dddi' = wn, elow, gf, isotope (linelist/repack.py)."""
    _record = np.dtype([('wn', '<f8'), ('elow', '<f8'), ('gf', '<f8'), ('iso', '<i4')])

    def __init__(self, dbfile, pffile, log=None):
        super().__init__(dbfile, pffile, log)
        import struct
        self.fmt = 'dddi'
        self.recsize = struct.calcsize(self.fmt)   # 32 with native alignment padding
        self.molecule, self.dbtype = os.path.split(dbfile)[1].split('_')[0:2]
        self.name = f'repack {self.dbtype} {self.molecule}'
        self.isotopes, self.mass, self.isoratio = self.get_iso(self.molecule)

    def dbread(self, iwn, fwn, verb=0):
        raw = np.fromfile(self.dbfile, dtype=np.uint8)
        nlines = len(raw) // self.recsize
        rec = raw[:nlines * self.recsize].reshape(nlines, self.recsize)
        body = np.ascontiguousarray(rec[:, :self._record.itemsize]).view(self._record).ravel()
        wn_all = body['wn'].astype(np.double)
        window = self._window(wn_all, iwn, fwn)
        if window is None:
            return None
        istart, istop = window
        sl = slice(istart, istop + 1)
        wnumber, elow, gf = wn_all[sl], body['elow'][sl].astype(np.double), \
            body['gf'][sl].astype(np.double)
        unique_iso, inverse = np.unique(body['iso'][sl], return_inverse=True)
        iso_len = len(self.isotopes[0])
        names = [str(i).zfill(iso_len) for i in unique_iso]
        missing = [n for n in names if n not in self.isotopes]
        if missing:
            raise ValueError(
                f'Unrecognized isotope names for {self.molecule} line-list: {missing}\n'
                'See list of known isotopes at pyratbay_b200/data/isotopes_subset.json')
        idx = np.array([self.isotopes.index(n) for n in names], np.int64)
        return wnumber, gf, elow, idx[inverse]


DB_READERS = {'hitran': Hitran, 'exomol': Exomol, 'repack': Repack}
