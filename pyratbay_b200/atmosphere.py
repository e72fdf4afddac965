"""Minimal atmosphere container: the attributes of pyrat.atm the hot path reads
(press [bar], temp, vmr, d, species, mol_mass [amu], mol_radius [cm]).

Species masses / collision radii are the entries of the reference's
pyratbay/data/molecules.dat (lines 32-237) for the species used in its test
atmospheres; radii in Angstrom, converted to cm like pyrat/atmosphere.py:328.
"""
import numpy as np

from . import constants as pc

# name: (mass g/mol, radius Angstrom)
MOLECULES = {
    'H': (1.008000, 1.10), 'H2': (2.016000, 1.44), 'He': (4.002602, 1.40),
    'H2O': (18.015000, 1.60), 'CH4': (16.043000, 2.00), 'CO': (28.010000, 1.69),
    'CO2': (44.009000, 1.90), 'NH3': (17.031000, 1.80), 'HCN': (27.026000, 2.50),
    'N2': (28.014000, 1.82), 'O2': (31.998800, 1.73), 'C2H2': (26.037300, 2.63),
    'Na': (22.989769, 2.20), 'K': (39.098300, 2.80), 'TiO': (63.866000, 2.43),
    'VO': (66.940500, 2.38), 'OH': (17.007000, 0.97),
}


def pressure(ptop, pbottom, nlayers):
    """Log-spaced pressure profile in bar (atmosphere/atmosphere.py:33-94)."""
    return np.logspace(np.log10(ptop), np.log10(pbottom), nlayers)


def ideal_gas_density(vmr, press_bar, temp):
    """Number density (molecules cm-3), same operation order as the reference's
    atmosphere/atmosphere.py:662-664 so that atm.d is bit-identical."""
    vmr = np.asarray(vmr, np.double)
    press_bar = np.asarray(press_bar, np.double)
    temp = np.asarray(temp, np.double)
    if np.shape(vmr) == np.shape(press_bar):
        return vmr * (press_bar * pc.bar) / (temp * pc.k)
    return vmr * np.expand_dims(press_bar / temp, axis=1) * pc.bar / pc.k


class Atmosphere:
    def __init__(self, press, temp, vmr, species, molecules=None):
        self.press = np.asarray(press, np.double)          # bar
        self.temp = np.asarray(temp, np.double)
        self.vmr = np.asarray(vmr, np.double)              # [nlayers, nmol]
        self.species = list(species)
        self.nlayers = len(self.press)
        self.nmol = len(self.species)
        if self.vmr.shape != (self.nlayers, self.nmol):
            raise ValueError("vmr must have shape [nlayers, nmol]")
        table = dict(MOLECULES)
        if molecules:
            table.update(molecules)
        absent = [s for s in self.species if s not in table]
        if absent:
            raise ValueError(f"These species: {absent} are not listed in the molecules table")
        self.mol_mass = np.array([table[s][0] for s in self.species])
        self.mol_radius = np.array([table[s][1] * pc.A for s in self.species])
        self.d = ideal_gas_density(self.vmr, self.press, self.temp)

    def calc_profiles(self, temp=None, vmr=None):
        """Update temperature / abundances and the densities (subset of
        pyrat/atmosphere.py calc_profiles)."""
        if temp is not None:
            self.temp = np.asarray(temp, np.double)
        if vmr is not None:
            self.vmr = np.asarray(vmr, np.double)
        self.d = ideal_gas_density(self.vmr, self.press, self.temp)


def read_atm(atmfile):
    """Read a Pyrat Bay .atm file (io/io.py read_atm format): returns
    (species, press [bar], temp [K], vmr [nlayers, nmol])."""
    units = {'pressure': 'bar', 'temperature': 'kelvin'}
    species, data = None, []
    section = None
    with open(atmfile) as f:
        for raw in f:
            line = raw.strip()
            if not line or line.startswith('#'):
                continue
            if line.startswith('@'):
                section = line[1:].upper()
                continue
            if section == 'PRESSURE':
                units['pressure'] = line
            elif section == 'TEMPERATURE':
                units['temperature'] = line
            elif section == 'SPECIES':
                species = line.split()
            elif section == 'DATA':
                data.append([float(v) for v in line.split()])
    if species is None or not data:
        raise ValueError(f"'{atmfile}' is not a valid atmospheric file")
    data = np.array(data)
    nmol = len(species)
    vmr = data[:, -nmol:]
    press = data[:, 0] * pc.u(units['pressure']) / pc.bar
    temp = data[:, 1]
    return species, press, temp, vmr
