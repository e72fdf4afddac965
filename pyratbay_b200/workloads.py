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


def partition(db, temps):
    """Z_i(T) [len(T), niso] with the reference's interpolant (scipy slinear)."""
    import scipy.interpolate as sip
    temps = np.atleast_1d(temps)
    return np.array([sip.interp1d(db.temp, db.iso_pf[j], kind='slinear')(temps)
                     for j in range(db.niso)]).T
