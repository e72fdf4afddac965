"""TLI (transition line information) binary files: reader, writer, synthetic generator.

Layout (writer pyratbay/opacity/lread.py:276-314, reader
pyratbay/pyrat/line_by_line.py:298-482).  The reader here uses np.memmap / np.fromfile per
column segment instead of struct.unpack tuples, so it scales to 1e8-1e9 lines.
"""
import os
import struct
import sys

import numpy as np

from . import constants as pc

TLI_VERSION = (6, 5, 0)   # pyratbay/version.py:10-12


class Database:
    """Header block of one line-list database (line_by_line.py:16-68)."""

    def __init__(self, name, molname, temp, iso_name, iso_mass, iso_ratio, iso_pf):
        self.name = name
        self.molname = molname
        self.temp = np.asarray(temp, np.double)
        self.ntemp = len(self.temp)
        self.iso_name = np.asarray(iso_name, 'U20')
        self.niso = len(self.iso_name)
        self.iso_mass = np.asarray(iso_mass, np.double)
        self.iso_ratio = np.asarray(iso_ratio, np.double)
        self.iso_pf = np.asarray(iso_pf, np.double).reshape(self.niso, self.ntemp)

    def __str__(self):
        """Same text as the reference's Database.__str__ (line_by_line.py:58-68)."""
        from .tools import Formatted_Write
        fw = Formatted_Write()
        fw.write('Database name (name): {:s}', self.name)
        fw.write('Species name (molname):  {:s}', self.molname)
        fw.write('Number of isotopes (niso): {:d}', self.niso)
        fw.write('Number of temperature samples (ntemp): {:d}', self.ntemp)
        fw.write('Temperature (temp, K):\n    {}', self.temp, prec=3, edge=3)
        fw.write('Partition function for each isotope (z):')
        for z in self.iso_pf:
            fw.write('    {}', z, fmt={'float': '{: .3e}'.format}, edge=3)
        return fw.text


def _read(fmt, f):
    size = struct.calcsize(fmt)
    data = f.read(size)
    if len(data) != size:
        raise ValueError("TLI file truncated while reading the header")
    out = struct.unpack(fmt, data)
    return out[0] if len(out) == 1 else out


def _read_str(f):
    n = _read('h', f)
    return f.read(n).decode('utf-8')


def read_tli_header(f):
    """Parse the header; returns (databases, n_transitions, niso_tran, data_offset, limits)."""
    f.seek(0)
    endian = f.read(1).decode()
    if sys.byteorder[0:1] != endian:
        raise ValueError(
            f"Incompatible endianness between TLI file ({endian}) and "
            f"Pyrat ({sys.byteorder[0:1]})")
    ver, vmin, rev = _read('3h', f)
    if ver != 6 or vmin not in [1, 2, 3, 4, 5]:
        raise ValueError(
            "Incompatible TLI version.  The TLI file must be created "
            "with Lineread version 6.1-6.5.")
    wn_lo, wn_hi = _read('2d', f)
    n_db = _read('h', f)
    databases = []
    for _ in range(n_db):
        name = _read_str(f)
        molname = _read_str(f)
        ntemp, niso = _read('2h', f)
        temp = np.frombuffer(f.read(8 * ntemp), np.double)
        iso_name, iso_mass, iso_ratio, iso_pf = [], [], [], []
        for _j in range(niso):
            iso_name.append(_read_str(f))
            mass, ratio = _read('2d', f)
            iso_mass.append(mass)
            iso_ratio.append(ratio)
            iso_pf.append(np.frombuffer(f.read(8 * ntemp), np.double))
        databases.append(Database(name, molname, temp, iso_name, iso_mass, iso_ratio, iso_pf))
    n_transitions = _read('i', f)
    n_iso = _read('i', f)
    niso_tran = np.atleast_1d(np.frombuffer(f.read(4 * n_iso), np.int32)).astype(np.int64)
    return databases, int(n_transitions), niso_tran, f.tell(), (wn_lo, wn_hi, (ver, vmin, rev))


def read_tli_file(tli_file, wn_low, wn_high, log=None):
    """Extract the transitions of `tli_file` inside [wn_low, wn_high] (cm-1).

    Same return value as the reference's read_tli_file (line_by_line.py:298-482):
    (databases, wn, gf, elow, iso_id), per-isotope blocks in file order.
    """
    with open(tli_file, "rb") as f:
        databases, n_transitions, niso_tran, init_wl, (lo, hi, _v) = read_tli_header(f)
        f.seek(0, 2)
        endrec = f.tell()
    if log is not None:
        if lo > wn_high or hi < wn_low:
            log.warning(
                f"TLI wavenumber range ({lo:.1f}--{hi:.1f} cm-1) does not overlap with "
                f"Pyrat wavenumber range ({wn_low:.1f}--{wn_high:.1f} cm-1).")
        elif lo > wn_low or hi < wn_high:
            log.warning(
                f"TLI wavenumber range ({lo:.1f}--{hi:.2f} cm-1) does not cover the full "
                f"Pyrat wavenumber range ({wn_low:.1f}--{wn_high:.1f} cm-1).")
    nrec = (endrec - init_wl) / pc.tlireclen
    if nrec != n_transitions:
        raise ValueError(
            f'The remaining data file size ({nrec:.1f}) does not '
            f'correspond to the number of transitions ({n_transitions})')
    init_iso = init_wl + n_transitions * pc.dreclen
    init_el = init_iso + n_transitions * pc.sreclen
    init_gf = init_el + n_transitions * pc.dreclen

    if n_transitions == 0:
        z = np.zeros(0)
        return databases, z, z.copy(), z.copy(), np.zeros(0, np.short)

    wn_all = np.memmap(tli_file, np.double, 'r', offset=init_wl, shape=(n_transitions,))
    segments = []
    offset = 0
    for n in niso_tran:
        n = int(n)
        block = wn_all[offset:offset + n]
        if n > 0:
            # tools.binsearch (tools/tools.py:219-311): index of the first record >= wn_low
            # (-1 if the whole block lies below) and of the last record <= wn_high (-1 if
            # the whole block lies above).
            ifirst = -1 if block[-1] < wn_low else int(np.searchsorted(block, wn_low, 'left'))
            ilast = -1 if wn_high < block[0] else int(np.searchsorted(block, wn_high, 'right')) - 1
            # line_by_line.py:423-432 adds the block offset BEFORE testing the -1 sentinels,
            # so for every block but the first a block lying entirely below the window is
            # not skipped: the reader returns the previous block's last record again plus
            # the whole block.  Kept bug-for-bug: the reference's tables contain that
            # duplicated line, and parity is defined on its output.  (The out-of-window
            # lines are dropped later by the window test of _extcoeff.c:215.)
            ifirst += offset
            ilast += offset
            if ifirst >= 0 and ilast >= 0:
                nread = ilast - ifirst + 1
                if nread > 0:
                    segments.append((ifirst, nread))
        offset += n
    nlt = sum(n for _, n in segments)
    wn = np.empty(nlt, np.double)
    elow = np.empty(nlt, np.double)
    gf = np.empty(nlt, np.double)
    isoid = np.empty(nlt, np.short)
    pos = 0
    del wn_all

    def read_into(offset, dest):
        # straight into the destination slice: no temporary array, no second pass; every
        # task has its own descriptor so that the column segments are read concurrently
        # (readinto releases the GIL; at 1e8 lines the file is 2.6 GB)
        with open(tli_file, "rb", buffering=0) as f:
            f.seek(offset)
            view = memoryview(dest).cast('B')
            done = 0
            while done < dest.nbytes:
                got = f.readinto(view[done:])
                if not got:
                    raise ValueError("TLI file truncated while reading the line records")
                done += got

    tasks = []
    for first, n in segments:
        # large segments are split so that the pool stays busy with few isotope blocks
        for lo in range(0, n, 1 << 24):
            m = min(1 << 24, n - lo)
            a, b = pos + lo, pos + lo + m
            tasks += [(init_wl + (first + lo) * pc.dreclen, wn[a:b]),
                      (init_iso + (first + lo) * pc.sreclen, isoid[a:b]),
                      (init_el + (first + lo) * pc.dreclen, elow[a:b]),
                      (init_gf + (first + lo) * pc.dreclen, gf[a:b])]
        pos += n
    if nlt < (1 << 20):
        for offset, dest in tasks:
            read_into(offset, dest)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
            for result in [pool.submit(read_into, *t) for t in tasks]:
                result.result()
    if log is not None:
        log.msg(f'There are {n_transitions:,d} line transitions in TLI file.', indent=2)
    return databases, wn, gf, elow, isoid


def write_tli(tli_file, databases, lines, wn_min, wn_max):
    """Write a TLI file (layout of lread.py:276-314).

    databases: list of Database.
    lines: list (one entry per database) of dicts with 'wn', 'elow', 'gf', 'iso_id'
        (0-based within the database) already sorted by (iso_id, wn), and 'n_lines_iso'.
    """
    def pack_str(f, s):
        b = s.encode('utf-8')
        f.write(struct.pack('h', len(b)))
        f.write(b)

    with open(tli_file, 'wb') as f:
        f.write(sys.byteorder[0].encode('utf-8'))
        f.write(struct.pack('3h', *TLI_VERSION))
        f.write(struct.pack('2d', wn_min, wn_max))
        f.write(struct.pack('h', len(databases)))
        for db in databases:
            pack_str(f, db.name)
            pack_str(f, db.molname)
            f.write(struct.pack('hh', db.ntemp, db.niso))
            f.write(np.asarray(db.temp, np.double).tobytes())
            for j in range(db.niso):
                pack_str(f, str(db.iso_name[j]))
                f.write(struct.pack('d', db.iso_mass[j]))
                f.write(struct.pack('d', db.iso_ratio[j]))
                f.write(np.asarray(db.iso_pf[j], np.double).tobytes())
        n_lines = int(sum(len(d['wn']) for d in lines))
        f.write(struct.pack('i', n_lines))
        n_lines_iso = np.concatenate([np.asarray(d['n_lines_iso'], np.int32) for d in lines])
        f.write(struct.pack('i', len(n_lines_iso)))
        f.write(n_lines_iso.astype(np.int32).tobytes())
        for d in lines:
            f.write(np.asarray(d['wn'], np.double).tobytes())
        for d in lines:
            f.write(np.asarray(d['iso_id'], np.int16).tobytes())
        for d in lines:
            f.write(np.asarray(d['elow'], np.double).tobytes())
        for d in lines:
            f.write(np.asarray(d['gf'], np.double).tobytes())


# H2O isotopologues of the benchmark line lists (values as in the reference's
# pyratbay/data/isotopes.dat:146-149 and tests/test_tli.py:32-35).
H2O_ISOTOPES = {
    'names': ['116', '118', '117', '126'],
    'mass': [18.010560, 20.014810, 19.014780, 19.016740],
    'ratio': [0.997317300, 0.001999827, 0.000371884, 0.000310693],
}


def h2o_partition_table():
    """(temp[ntemp], Z[4, ntemp]) TIPS-2021 H2O partition functions, extracted from a TLI
    written by the reference (pyratbay_b200/data/h2o_partition.npz; tests/golden/make_golden.py)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data',
                        'h2o_partition.npz')
    with np.load(path) as d:
        return d['temp'].copy(), d['z'].copy()


def synthetic_lines(nlines, wn_low, wn_high, fractions=(0.75, 0.15, 0.07, 0.03), seed=0):
    """Seeded synthetic line list with the statistics of SURVEY.md section 8d:
    wn ~ U(wn_low, wn_high) sorted per isotope, elow ~ U(10, 8000) cm-1,
    log10 gf ~ U(-12, -4).  Returns (wn, elow, gf, iso_id, n_lines_iso) in TLI order."""
    rng = np.random.default_rng(seed)
    counts = np.floor(np.asarray(fractions) * nlines).astype(np.int64)
    counts[0] += nlines - counts.sum()
    wn = np.empty(nlines, np.double)
    iso = np.empty(nlines, np.int16)
    pos = 0
    for i, n in enumerate(counts):
        w = rng.uniform(wn_low, wn_high, int(n))
        w.sort()
        wn[pos:pos + n] = w
        iso[pos:pos + n] = i
        pos += n
    elow = rng.uniform(10.0, 8000.0, nlines)
    gf = 10.0 ** rng.uniform(-12.0, -4.0, nlines)
    return wn, elow, gf, iso, counts


def synthetic_h2o_database():
    temp, z = h2o_partition_table()
    return Database('Synthetic H2O', 'H2O', temp, H2O_ISOTOPES['names'],
                    H2O_ISOTOPES['mass'], H2O_ISOTOPES['ratio'], z)


def make_synthetic_tli(tli_file, nlines, wn_low, wn_high, seed=0):
    """Write a synthetic single-database H2O TLI file; returns the Database header."""
    db = synthetic_h2o_database()
    wn, elow, gf, iso, counts = synthetic_lines(nlines, wn_low, wn_high, seed=seed)
    write_tli(tli_file, [db],
              [{'wn': wn, 'elow': elow, 'gf': gf, 'iso_id': iso, 'n_lines_iso': counts}],
              wn_low, wn_high)
    return db
