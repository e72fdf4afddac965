"""A light `Pyrat` container exposing the attributes the reference's hot path reads
(pyratbay/pyrat/pyrat_obj.py:31-126): spec, atm, ex, opacity (models / models_type),
voigt, log, inputs, ncpu -- built around one GPU engine.  It covers runmode=opacity and the
LBL extinction stage; everything else of the reference's Pyrat object is out of scope.
"""
from types import SimpleNamespace

import numpy as np

from . import atmosphere as pa
from . import extinction as ex_mod
from . import tools as pt
from .engine import Engine
from .line_by_line import Line_By_Line
from .spectrum import Spectrum
from .voigt import Voigt

_INPUT_DEFAULTS = dict(
    wnlow=None, wnhigh=None, wl_low=None, wl_high=None, wnstep=None, wnosamp=None,
    resolution=None, wlstep=None, tmin=None, tmax=None, tstep=None, ethresh=1e-30,
    voigt_extent=300.0, voigt_cutoff=25.0, voigt_ndop=50, voigt_nlor=100, voigt_dmin=None,
    voigt_dmax=None, voigt_lmin=None, voigt_lmax=None, voigt_dlratio=0.1, tlifile=None,
    sampled_cs=None, single_isotope=None, atmfile=None, ptop=None, pbottom=None,
    nlayers=None, ncpu=1, verb=2, logfile=None, runmode='opacity',
)


class Pyrat:
    def __init__(self, inputs, atm=None, device=0, log=None):
        """inputs: a config-file path, a dict or a namespace with the `[pyrat]` keys of
        SURVEY.md section 5.  atm: an atmosphere.Atmosphere (else read from inputs.atmfile)."""
        if isinstance(inputs, str):
            inputs = pt.parse(inputs)
        elif isinstance(inputs, dict):
            inputs = SimpleNamespace(**inputs)
        for key, val in _INPUT_DEFAULTS.items():
            if not hasattr(inputs, key):
                setattr(inputs, key, val)
        if isinstance(inputs.tlifile, str):
            inputs.tlifile = [inputs.tlifile]
        if isinstance(inputs.sampled_cs, str):
            inputs.sampled_cs = [inputs.sampled_cs]
        self.inputs = inputs
        self.log = log if log is not None else pt.Log(None, verb=inputs.verb)
        self.ncpu = inputs.ncpu
        self.device = device
        self.last_timing = None

        self.spec = Spectrum(
            wnlow=inputs.wnlow, wnhigh=inputs.wnhigh, wl_low=inputs.wl_low,
            wl_high=inputs.wl_high, wnstep=inputs.wnstep, wnosamp=inputs.wnosamp,
            resolution=inputs.resolution, wlstep=inputs.wlstep, log=self.log)

        if atm is None:
            if inputs.atmfile is None:
                raise ValueError("an atmosphere (atm=...) or inputs.atmfile is required")
            species, press, temp, vmr = pa.read_atm(inputs.atmfile)
            atm = pa.Atmosphere(press, temp, vmr, species)
        self.atm = atm

        self.ex = SimpleNamespace(
            tmin=inputs.tmin, tmax=inputs.tmax, tstep=inputs.tstep,
            sampled_cs=inputs.sampled_cs, ntemp=None, temp=None, z=None, etable=None)

        self.engine = Engine(device)
        self.engine.set_grid(self.spec.wn, self.spec.own, self.spec.odivisors)

        self.opacity = SimpleNamespace(models=[], models_type=[])
        self.lbl = None
        self.voigt = None
        if inputs.tlifile is not None:
            self.lbl = Line_By_Line(
                inputs.tlifile, self.atm.species, self.spec.wnlow, self.spec.wnhigh, self,
                ethresh=inputs.ethresh, single_isotope=inputs.single_isotope, log=self.log)
            self.opacity.models.append(self.lbl)
            self.opacity.models_type.append('lbl')
            self.engine.set_species(
                self.atm.mol_radius, self.atm.mol_mass, self.lbl.iso_atm_index,
                self.lbl.iso_mass, self.lbl.iso_ratio)
            self.engine.set_lines(self.lbl.wn, self.lbl.elow, self.lbl.gf, self.lbl.isoid)
            self.voigt = Voigt(
                self.spec, self.atm, self.lbl.iso_atm_index, self.engine,
                extent=inputs.voigt_extent, cutoff=inputs.voigt_cutoff,
                dlratio=inputs.voigt_dlratio, ndop=inputs.voigt_ndop, nlor=inputs.voigt_nlor,
                dmin=inputs.voigt_dmin, dmax=inputs.voigt_dmax, lmin=inputs.voigt_lmin,
                lmax=inputs.voigt_lmax, tmin=inputs.tmin, tmax=inputs.tmax, log=self.log)

    def compute_opacity(self, **kwargs):
        """Cross-section table build (pyrat_obj.py:121-126); keyword options of
        extinction.compute_opacity (write, host, nchunks)."""
        ex_mod.compute_opacity(self, **kwargs)

    def calc_lbl_extinction(self, temp=None, vmr=None, skip_mol=[]):
        """The LBL part of Pyrat.run's extinction stage (pyrat_obj.py:203-206): update the
        atmosphere, then extinction coefficient (cm-1) [nlayers, nwave]."""
        self.atm.calc_profiles(temp, vmr)
        return self.lbl.calc_extinction_coefficient(
            self.atm.temp, self.atm.d[:, self.lbl.mol_index], skip_mol=skip_mol)

    def get_ec(self, layer):
        """Per-species LBL extinction at one layer (pyrat_obj.py get_ec -> opacity.get_ec)."""
        density = self.atm.d[:, self.lbl.mol_index]
        ec = self.lbl.calc_extinction_coefficient(self.atm.temp, density, layer=layer)
        return ec, list(self.lbl.species)


def run(cfile, device=0):
    """Driver for runmode = tli (driver.py:35-46) and runmode = opacity (driver.py:59-61)."""
    inputs = pt.parse(cfile)
    if inputs.runmode == 'tli':
        from . import constants as pc
        from . import lread
        log = pt.Log(inputs.logfile, verb=inputs.verb if inputs.verb is not None else 2)
        if inputs.tlifile is None:
            log.error('Undefined TLI file (tlifile)')
        # wavelengths back to their original units, as the reference passes them
        wl_low = inputs.wl_low / pc.u(inputs.wlunits) if inputs.wl_low is not None else None
        wl_high = inputs.wl_high / pc.u(inputs.wlunits) if inputs.wl_high is not None else None
        lread.make_tli(inputs.dblist, inputs.pflist, inputs.dbtype, inputs.tlifile[0],
                       wl_low, wl_high, inputs.wlunits, log)
        return None
    pyrat = Pyrat(cfile, device=device)
    if pyrat.inputs.runmode == 'opacity':
        pyrat.compute_opacity()
    return pyrat
