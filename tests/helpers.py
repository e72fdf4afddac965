"""Shared set-up for the tests: builds the array bundle that one reference
ec.extinction(...) call takes (pyratbay/pyrat/extinction.py:197-208), from either the
golden mock-HITRAN case or a seeded synthetic line list."""
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pyratbay_b200 import atmosphere as pa  # noqa: E402
from pyratbay_b200 import constants as pc  # noqa: E402
from pyratbay_b200 import tli as ptli  # noqa: E402
from pyratbay_b200.spectrum import Spectrum  # noqa: E402
from pyratbay_b200.voigt import Voigt  # noqa: E402


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=True)


def mock_atmosphere():
    a = golden("mock_atmosphere.npz")
    return pa.Atmosphere(a["press"], a["temp"], a["vmr"], [str(s) for s in a["species"]])


def oracle_module():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle  # noqa: E402
    return oracle


class Case(SimpleNamespace):
    """Static inputs of the hot path for one configuration."""

    def unit_args(self, temp, density, isoz, iso_iext=None):
        """Positional arguments 2..24 of ec.extinction for one (T,p) unit."""
        if iso_iext is None:
            iso_iext = self.iso_mol_index
        return (self.profile, self.size, self.index, self.lorentz, self.doppler,
                self.spec.wn, self.spec.own, self.spec.odivisors,
                density, self.atm.mol_radius, self.atm.mol_mass,
                self.iso_atm_index, self.iso_mass, self.iso_ratio, isoz, iso_iext,
                self.lwn, self.elow, self.gf, self.isoid,
                self.cutoff, self.ethresh, temp)


def build_case(spec, atm, db, lwn, elow, gf, isoid, extent=300.0, cutoff=25.0,
               tmin=None, tmax=None, ethresh=1e-30, with_profile=True, nlor=100, ndop=50,
               dlratio=0.1):
    import scipy.interpolate as sip
    species = list(atm.species)
    niso = db.niso
    iso_atm_index = np.full(niso, species.index(db.molname), int)
    v = Voigt(spec, atm, iso_atm_index, None, extent=extent, cutoff=cutoff, tmin=tmin,
              tmax=tmax, nlor=nlor, ndop=ndop, dlratio=dlratio)
    case = Case(
        spec=spec, atm=atm, db=db, voigt=v, lorentz=v.lorentz, doppler=v.doppler,
        size_in=v.size.copy(), size=v.size, index=v.index, cutoff=cutoff, ethresh=ethresh,
        iso_atm_index=iso_atm_index, iso_mass=db.iso_mass.copy(),
        iso_ratio=db.iso_ratio.copy(), iso_mol_index=np.zeros(niso, int),
        lwn=np.asarray(lwn, np.double), elow=np.asarray(elow, np.double),
        gf=np.asarray(gf, np.double), isoid=np.asarray(isoid, int), nspec=1,
        pf_interp=[sip.interp1d(db.temp, db.iso_pf[j], kind='slinear') for j in range(niso)],
        profile=None)
    if with_profile:
        orc = oracle_module()
        case.profile = np.zeros(v.profile_len, np.double)
        orc.grid(case.profile, case.size, case.index, case.lorentz, case.doppler,
                 spec.ownstep)
    return case


def partition(case, temps):
    temps = np.atleast_1d(temps)
    return np.array([f(temps) for f in case.pf_interp])  # [niso, ntemp]


_CACHE = {}


def mock_case(resolution=None, with_profile=True):
    """The reference-runnable configuration (BASELINE.json configs[0]): mock HITRAN H2O,
    1.00-1.01 um, wnstep 1, wnosamp 2160, extent 100, T = 300..3000 K."""
    key = ("mock", resolution, with_profile)
    if key in _CACHE:
        return _CACHE[key]
    spec = Spectrum(wl_low=1.00 * pc.um, wl_high=1.01 * pc.um, wnstep=1.0, wnosamp=2160,
                    resolution=resolution)
    atm = mock_atmosphere()
    tlifile = os.path.join(GOLDEN, "mock_hitran_h2o.tli")
    dbs, wn, gf, elow, iso = ptli.read_tli_file(tlifile, spec.wnlow, spec.wnhigh)
    case = build_case(spec, atm, dbs[0], wn, elow, gf, iso, extent=100.0, cutoff=25.0,
                      tmin=300.0, tmax=3000.0, with_profile=with_profile)
    case.tlifile = tlifile
    _CACHE[key] = case
    return case


def synthetic_case(nlines=20000, wnlow=9000.0, wnhigh=9200.0, wnstep=1.0, wnosamp=720,
                   seed=0, extent=50.0, cutoff=25.0, nlayers=9, resolution=None,
                   nlor=30, ndop=12, with_profile=True, ethresh=1e-30):
    """Seeded synthetic H2O line list on a small grid (profile table of a few MB)."""
    key = ("syn", nlines, wnlow, wnhigh, wnstep, wnosamp, seed, extent, cutoff, nlayers,
           resolution, nlor, ndop, with_profile, ethresh)
    if key in _CACHE:
        return _CACHE[key]
    spec = Spectrum(wnlow=wnlow, wnhigh=wnhigh, wnstep=wnstep, wnosamp=wnosamp,
                    resolution=resolution)
    a = golden("mock_atmosphere.npz")
    press = pa.pressure(1e-6, 100.0, nlayers)
    temp = np.linspace(400.0, 2600.0, nlayers)
    vmr = np.tile(a["vmr"][0], (nlayers, 1))
    atm = pa.Atmosphere(press, temp, vmr, [str(s) for s in a["species"]])
    db = ptli.synthetic_h2o_database()
    lwn, elow, gf, iso, _ = ptli.synthetic_lines(nlines, wnlow - 5.0, wnhigh + 5.0, seed=seed)
    case = build_case(spec, atm, db, lwn, elow, gf, iso, extent=extent, cutoff=cutoff,
                      tmin=300.0, tmax=3000.0, nlor=nlor, ndop=ndop,
                      with_profile=with_profile, ethresh=ethresh)
    _CACHE[key] = case
    return case
