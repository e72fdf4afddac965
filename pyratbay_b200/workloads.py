"""Synthetic workloads of BASELINE.json, shared by bench.py, the
CPU-baseline script and the tests so that every arm sees identical inputs."""
from types import SimpleNamespace

import numpy as np

from . import atmosphere as pa
from . import constants as pc
from . import tli as ptli
from .spectrum import Spectrum

# Uniform composition of the reference's tests/inputs/atmosphere_uniform_test.atm
UNIFORM_SPECIES = ['H2', 'He', 'H', 'Na', 'K', 'H2O', 'CH4', 'CO', 'CO2']
UNIFORM_VMR = [8.5e-01, 1.49e-01, 1.0e-06, 3.0e-06, 5.0e-08, 4.0e-04, 1.0e-04, 5.0e-04, 1.0e-07]


def layer_temperatures(nlayers, realization=0, ttop=800.0, tbottom=2200.0):
    """Smooth T(p): `ttop` at the top to `tbottom` at the bottom, linear in log p, plus a small
    realization-dependent offset (a retrieval evaluates a new profile every step)."""
    base = np.linspace(ttop, tbottom, nlayers)
    return base + 0.37 * (realization % 97)


def forward_model_workload(nlines=1_000_000, nlayers=81, wl_low_um=0.5, wl_high_um=5.0,
                           wnstep=1.0, wnosamp=2160, seed=0, ptop=1e-6, pbottom=100.0):
    """configs[1]: synthetic H2O line list, `nlayers`-layer atmosphere 1e-6..100 bar,
    forward-model extinction (add=1) on a constant-step wavenumber grid."""
    spec = Spectrum(wl_low=wl_low_um * pc.um, wl_high=wl_high_um * pc.um, wnstep=wnstep,
                    wnosamp=wnosamp)
    press = pa.pressure(ptop, pbottom, nlayers)
    vmr = np.tile(np.asarray(UNIFORM_VMR), (nlayers, 1))
    atm = pa.Atmosphere(press, layer_temperatures(nlayers), vmr, UNIFORM_SPECIES)
    db = ptli.synthetic_h2o_database()
    wn, elow, gf, iso, counts = ptli.synthetic_lines(nlines, spec.wnlow, spec.wnhigh, seed=seed)
    return SimpleNamespace(
        spec=spec, atm=atm, db=db, wn=wn, elow=elow, gf=gf, isoid=iso.astype(int),
        iso_atm_index=np.full(db.niso, UNIFORM_SPECIES.index('H2O'), int),
        iso_mol_index=np.zeros(db.niso, int), nlines=nlines, nlayers=nlayers,
        name=(f"synthetic {nlines:.0e}-line H2O, {nlayers}-layer atmosphere, "
              f"{wl_low_um}-{wl_high_um} um forward-model extinction"))


def table_workload(nlines=100_000_000, ntemp=20, nlayers=51, nwave=100_000, wl_low_um=0.3,
                   wl_high_um=30.0, tmin=300.0, tstep=150.0, ptop=1e-6, pbottom=100.0,
                   window=None):
    """configs[2]: cross-section table of a synthetic ExoMol-scale H2O list over
    ntemp x nlayers (T,p) units and a constant-step grid of `nwave` samples (add=0).

    The grid is defined by round numbers so that every arm builds the identical one:
    wnlow = 1/wl_high, wnstep = (1/wl_low - 1/wl_high)/nwave rounded to 2 digits, wnosamp from
    the reference's 4e-4 rule (spectrum.py:187-190).  `window` = (first, count) restricts the
    workload to that slice of the output grid with the same line density (the bounded CPU
    sample of bench.py): lines ~ U over the sub-window, nlines scaled by its share.

    Returns a namespace with the `Pyrat` input keys (`inputs`), the atmosphere, the database
    header and `make_lines()` -> (wn, elow, gf, iso_id, counts)."""
    from .spectrum import _HCN
    wnlow = 1.0 / (wl_high_um * pc.um)
    wnstep = float(f"{(1.0 / (wl_low_um * pc.um) - wnlow) / nwave:.2g}")
    wnosamp = int(_HCN[wnstep / _HCN <= 0.0004][0])
    first, count = (0, nwave) if window is None else window
    lo = wnlow + first * wnstep
    hi = lo + (count - 0.5) * wnstep                 # int((hi-lo)/wnstep) + 1 == count
    nlines_w = int(round(nlines * count / nwave))
    tmax = tmin + (ntemp - 1) * tstep
    press = pa.pressure(ptop, pbottom, nlayers)
    vmr = np.tile(np.asarray(UNIFORM_VMR), (nlayers, 1))
    atm = pa.Atmosphere(press, np.full(nlayers, 1000.0), vmr, UNIFORM_SPECIES)
    db = ptli.synthetic_h2o_database()
    inputs = dict(wnlow=lo, wnhigh=hi, wnstep=wnstep, wnosamp=wnosamp, tmin=tmin, tmax=tmax,
                  tstep=tstep, verb=0)
    return SimpleNamespace(
        inputs=inputs, atm=atm, db=db, nlines=nlines_w, ntemp=ntemp, nlayers=nlayers,
        nwave=count, wnstep=wnstep, wnosamp=wnosamp, temps=tmin + tstep * np.arange(ntemp),
        iso_atm_index=np.full(db.niso, UNIFORM_SPECIES.index('H2O'), int),
        iso_mol_index=np.zeros(db.niso, int),
        make_lines=lambda seed=0: ptli.synthetic_lines(nlines_w, lo, lo + (count - 1) * wnstep,
                                                       seed=seed),
        name=(f"synthetic {nlines:.0e}-line H2O, {ntemp} T ({tmin:g}-{tmax:g} K) x {nlayers} p "
              f"({ptop:g}-{pbottom:g} bar) x {nwave} wn ({wl_low_um:g}-{wl_high_um:g} um, "
              f"wnstep {wnstep:g} cm-1, wnosamp {wnosamp}) cross-section table, add=0"))


def partition(db, temps):
    """Z_i(T) [len(T), niso] with the reference's interpolant (scipy slinear)."""
    import scipy.interpolate as sip
    temps = np.atleast_1d(temps)
    return np.array([sip.interp1d(db.temp, db.iso_pf[j], kind='slinear')(temps)
                     for j in range(db.niso)]).T
