"""make_tli: line-list databases -> TLI file.

Drop-in for the reference's `pyratbay.opacity.make_tli` (pyratbay/opacity/lread.py:36-318): same
arguments, database grouping, isotope filtering, (isotope, wavenumber) ordering and file
layout, so the output is byte-identical (tests/test_linelist.py).  The records come from the
column-wise readers of linelist.py and are written by tli.write_tli in whole-array writes (the
reference packs Python lists through struct).
"""
import os

import numpy as np

from . import constants as pc
from . import linelist
from . import tli as ptli


def _as_list(value, n=None):
    """A scalar string becomes a one-element list; a one-element list is repeated n times."""
    items = [value] if isinstance(value, str) else list(value)
    if n is not None and len(items) == 1:
        items = items * n
    return items


def _collect(databases, name, wn_low, wn_high, verb):
    """Concatenated (wn, gf, elow, iso_id) of every input file that belongs to database `name`,
    plus the reader object that describes the database."""
    columns = ([], [], [], [])
    reader = None
    for db in databases:
        if db.name != name:
            continue
        reader = db
        found = db.dbread(wn_low, wn_high, verb)
        if found is not None:
            for column, values in zip(columns, found):
                column.append(values)
    wn, gf, elow, iso_id = (np.concatenate(c) for c in columns)
    return reader, wn, gf, elow, iso_id


def _tli_order(wn, iso_id, counts):
    """Permutation that sorts by isotope and, inside every isotope block, by wavenumber
    (the two-stage argsort of lread.py:186-195, reproduced call for call so that ties land
    where the reference puts them)."""
    order = np.argsort(iso_id)
    stop = 0
    for n in counts:
        start, stop = stop, stop + n
        block = order[start:stop]
        order[start:stop] = block[np.argsort(wn[block])]
    return order


def make_tli(dblist, pflist, dbtype, tlifile, wl_low, wl_high, wl_units='um', log=None):
    """Create a TLI file.

    dblist, pflist, dbtype: (lists of) database files, partition-function sources ('tips' or
    a tabulated file) and database types ('hitran', 'exomol', 'repack').
    wl_low, wl_high: wavelength boundaries in `wl_units`.
    Returns the list of tli.Database headers written.
    """
    log = log if log is not None else linelist._NullLog()
    verb = getattr(log, 'verb', 0)
    required = [
        (tlifile, 'Undefined TLI file (tlifile).'),
        (wl_low, 'Undefined low wavelength boundary (wl_low)'),
        (wl_high, 'Undefined high wavelength boundary (wl_high)'),
        (dblist, 'There are no input database files (dblist)'),
        (dbtype, 'There are no input database types (dbtype)'),
        (pflist, 'There are no partition-function inputs (pflist)'),
    ]
    for value, message in required:
        if value is None:
            log.error(message)

    files = [os.path.realpath(path) for path in _as_list(dblist)]
    partitions = _as_list(pflist, len(files))
    kinds = _as_list(dbtype, len(files))
    if not (len(files) == len(partitions) == len(kinds)):
        log.error(
            f'The number of Line-transition files ({len(files)}) does not match '
            f'the number of partition-function files ({len(partitions)}) or '
            f'database-type files ({len(kinds)})')

    log.head('\nReading input database files:')
    readers = []
    for path, pf, kind in zip(files, partitions, kinds):
        if kind not in linelist.DB_READERS:
            log.error(f"Unknown type '{kind}' for database '{path}'.  "
                      f"Select from: {sorted(linelist.DB_READERS)}")
        log.head(path, indent=2)
        readers.append(linelist.DB_READERS[kind](path, pf, log))
    names = list(dict.fromkeys(db.name for db in readers))   # unique, first-seen order
    log.msg(f'There are {len(files)} input database file(s).\n\n')

    # wavelength window -> wavenumber window in cm-1 (lread.py:129-131)
    wn_low = 1.0 / wl_high / pc.u(wl_units)
    wn_high = 1.0 / wl_low / pc.u(wl_units)

    headers, blocks = [], []
    for name in names:
        db, wn, gf, elow, iso_id = _collect(readers, name, wn_low, wn_high, verb)
        # iso_id indexes db.isotopes; local ids run 0..N-1 over the isotopes that have lines
        present, local_id, counts = np.unique(iso_id, return_inverse=True, return_counts=True)
        order = _tli_order(wn, iso_id, counts)

        iso_names = np.array(db.isotopes)[present]
        temp, partition, pf_names = db.getpf(verb)
        missing = ~np.isin(iso_names, pf_names)
        if np.any(missing):
            log.error('No partition functions found for these isotopes of the '
                      f'{db.molecule} line list: {iso_names[missing]}')
        rows = [pf_names.index(iso) for iso in iso_names]
        headers.append(ptli.Database(
            db.name, db.molecule, temp, iso_names, np.array(db.mass)[present],
            np.array(db.isoratio)[present], np.asarray(partition)[rows]))
        blocks.append({'wn': wn[order], 'elow': elow[order], 'gf': gf[order],
                       'iso_id': local_id[order], 'n_lines_iso': counts})
        log.msg(f"Database '{db.name}' ({db.molecule}): {len(wn):,d} line transitions, "
                f"{len(iso_names)} isotopes.", indent=2)

    ptli.write_tli(tlifile, headers, blocks, wn_low, wn_high)
    log.head(f"Generated TLI file: '{tlifile}'.")
    return headers
