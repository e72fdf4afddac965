"""make_tli: line-list databases -> TLI file (pyratbay/opacity/lread.py:36-318).

Same arguments, database grouping, isotope filtering, (isotope, wavenumber) ordering and
file layout as the reference; the records are read by the vectorised readers of linelist.py
and written column-wise by tli.write_tli (the reference packs Python lists with struct).
"""
import os

import numpy as np

from . import constants as pc
from . import linelist
from . import tli as ptli


def make_tli(dblist, pflist, dbtype, tlifile, wl_low, wl_high, wl_units='um', log=None):
    """Create a TLI file.

    dblist, pflist, dbtype: (lists of) database files, partition-function sources ('tips' or
    a tabulated file) and database types ('hitran', 'exomol', 'repack').
    wl_low, wl_high: wavelength boundaries in `wl_units`.
    Returns the list of tli.Database headers written.
    """
    if log is None:
        log = linelist._NullLog()
    if tlifile is None:
        log.error('Undefined TLI file (tlifile).')
    if wl_low is None:
        log.error('Undefined low wavelength boundary (wl_low)')
    if wl_high is None:
        log.error('Undefined high wavelength boundary (wl_high)')
    if dblist is None:
        log.error('There are no input database files (dblist)')
    if dbtype is None:
        log.error('There are no input database types (dbtype)')
    if pflist is None:
        log.error('There are no partition-function inputs (pflist)')

    if isinstance(dblist, str):
        dblist = [dblist]
    nfiles = len(dblist)
    if isinstance(pflist, str):
        pflist = [pflist]
    if len(pflist) == 1:
        pflist = [pflist[0] for _ in range(nfiles)]
    if isinstance(dbtype, str):
        dbtype = [dbtype]
    if len(dbtype) == 1:
        dbtype = [dbtype[0] for _ in range(nfiles)]
    if nfiles != len(pflist) or nfiles != len(dbtype):
        log.error(
            f'The number of Line-transition files ({nfiles}) does not match '
            f'the number of partition-function files ({len(pflist)}) or '
            f'database-type files ({len(dbtype)})')

    dblist = [os.path.realpath(dbase) for dbase in dblist]
    databases, unique_dbs = [], []
    log.head('\nReading input database files:')
    for dbase, pf, dtype in zip(dblist, pflist, dbtype):
        if dtype not in linelist.DB_READERS:
            log.error(f"Unknown type '{dtype}' for database '{dbase}'.  "
                      f"Select from: {sorted(linelist.DB_READERS)}")
        log.head(dbase, indent=2)
        db = linelist.DB_READERS[dtype](dbase, pf, log)
        databases.append(db)
        if db.name not in unique_dbs:
            unique_dbs.append(db.name)
    log.msg(f'There are {nfiles} input database file(s).\n\n')

    # Boundaries in wavenumber space (cm-1), lread.py:129-131
    wn_low = 1.0 / wl_high / pc.u(wl_units)
    wn_high = 1.0 / wl_low / pc.u(wl_units)

    headers, lines = [], []
    for db_name in unique_dbs:
        wn, gf, elow, iso_id = [], [], [], []
        this_db = None
        for db in databases:
            if db.name != db_name:
                continue
            this_db = db
            transitions = db.dbread(wn_low, wn_high, getattr(log, 'verb', 0))
            if transitions is None:
                continue
            wn.append(transitions[0])
            gf.append(transitions[1])
            elow.append(transitions[2])
            iso_id.append(transitions[3])
        db = this_db
        wn = np.concatenate(wn)
        gf = np.concatenate(gf)
        elow = np.concatenate(elow)
        iso_id = np.concatenate(iso_id)

        # iso_id indexes db.isotopes; iso_idx runs 0..N-1 over the isotopes that have lines
        unique_iso, iso_idx, ntrans_iso = np.unique(
            iso_id, return_inverse=True, return_counts=True)
        # Sort by isotope, then each isotope by wavenumber (lread.py:186-203): argsort of the
        # isotope ids first, an argsort of the wavenumbers inside every block second.
        isort = np.argsort(iso_id)
        ihi = 0
        for ntrans in ntrans_iso:
            ilo = ihi
            ihi += ntrans
            block = isort[ilo:ihi]
            isort[ilo:ihi] = block[np.argsort(wn[block])]
        wn, gf, elow, iso_idx = wn[isort], gf[isort], elow[isort], iso_idx[isort]

        iso_names = np.array(db.isotopes)[unique_iso]
        iso_mass = np.array(db.mass)[unique_iso]
        iso_ratio = np.array(db.isoratio)[unique_iso]
        temp, partition, pf_iso = db.getpf(getattr(log, 'verb', 0))
        iso_match = np.isin(iso_names, pf_iso)
        if np.any(~iso_match):
            log.error('No partition functions found for these isotopes of the '
                      f'{db.molecule} line list: {iso_names[~iso_match]}')
        pf_idx = [pf_iso.index(iso) for iso in iso_names]
        headers.append(ptli.Database(db.name, db.molecule, temp, iso_names, iso_mass,
                                     iso_ratio, np.asarray(partition)[pf_idx]))
        lines.append({'wn': wn, 'elow': elow, 'gf': gf, 'iso_id': iso_idx,
                      'n_lines_iso': ntrans_iso})
        log.msg(f"Database '{db.name}' ({db.molecule}): {len(wn):,d} line transitions, "
                f"{len(iso_names)} isotopes.", indent=2)

    ptli.write_tli(tlifile, headers, lines, wn_low, wn_high)
    log.head(f"Generated TLI file: '{tlifile}'.")
    return headers
