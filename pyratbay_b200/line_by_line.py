"""Line-by-line data and on-the-fly extinction, mirroring
pyratbay/pyrat/line_by_line.py:71-248 on top of the GPU engine."""
import numpy as np
import scipy.interpolate as sip

from . import extinction as ex
from .tli import read_tli_file
from ._mem import pinned_zeros


class Line_By_Line:
    """All line-by-line data.  Attributes follow the reference: wn, elow, gf, isoid,
    iso_name/iso_mass/iso_ratio/iso_atm_index/iso_mol_index/iso_pf_interp, species, nspec,
    niso, ntransitions, tmin, tmax, ethresh, ec."""

    def __init__(self, tlifile, species, wn_low, wn_high, pyrat, ethresh=1e-30,
                 single_isotope=None, log=None):
        self.name = 'line by line'
        self.pyrat = pyrat
        self.ethresh = ethresh
        self.tlifile = [tlifile] if isinstance(tlifile, str) else list(tlifile)
        self.db = []
        self.wn = np.array([], np.double)
        self.elow = np.array([], np.double)
        self.gf = np.array([], np.double)
        self.isoid = np.array([], int)
        self.nwave = pyrat.spec.nwave
        self.nlayers = pyrat.atm.nlayers
        self.ec = pinned_zeros((self.nlayers, self.nwave))

        # Collect all databases; isotope ids are offset per *file* (line_by_line.py:112-125)
        # (one concatenation at the end, none for a single file: at 1e8 lines every avoidable
        # copy of a column is 0.8 GB)
        cols = {'wn': [], 'gf': [], 'elow': [], 'isoid': []}
        for tli_file in self.tlifile:
            databases, wn, gf, elow, iso_id = read_tli_file(tli_file, wn_low, wn_high, log)
            niso = int(np.sum([db.niso for db in self.db]))
            iso_id = iso_id.astype(int)
            if niso:
                iso_id += niso
            self.db += databases
            cols['wn'].append(wn)
            cols['gf'].append(gf)
            cols['elow'].append(elow)
            cols['isoid'].append(iso_id)
        for key, parts in cols.items():
            if len(parts) == 1:
                setattr(self, key, parts[0])
            elif parts:
                setattr(self, key, np.concatenate(parts))
        self.isoid = np.asarray(self.isoid, int)

        self.tmin = np.amax([np.amin(db.temp) for db in self.db])
        self.tmax = np.amin([np.amax(db.temp) for db in self.db])
        self.ndb = len(self.db)

        # Isotopic info (line_by_line.py:134-159)
        niso = self.niso = int(np.sum([db.niso for db in self.db]))
        species = list(species)
        self.iso_name = []
        self.iso_mass = np.zeros(niso)
        self.iso_ratio = np.zeros(niso)
        self.iso_atm_index = np.zeros(niso, int)
        self.iso_pf_interp = []
        self._db_pf_interp = []
        mol_names = []
        total = 0
        for db in self.db:
            sl = slice(total, total + db.niso)
            self.iso_name += list(db.iso_name)
            self.iso_mass[sl] = db.iso_mass
            self.iso_ratio[sl] = db.iso_ratio
            if db.molname not in species:
                raise ValueError(
                    f"The species '{db.molname}' is not present in the "
                    "atmosphere, required for LBL calculation")
            mol_names.append(db.molname)
            self.iso_atm_index[sl] = species.index(db.molname)
            for j in range(db.niso):
                self.iso_pf_interp.append(
                    sip.interp1d(db.temp, db.iso_pf[j], kind='slinear'))
            # the same interpolant over all isotopes of the database at once (one SciPy call
            # per database instead of one per isotope in `partition`; identical values)
            self._db_pf_interp.append(
                (sl, sip.interp1d(db.temp, db.iso_pf, kind='slinear', axis=-1)))
            total += db.niso

        # Single out an isotope if requested (line_by_line.py:161-175)
        if single_isotope is not None:
            if single_isotope not in self.iso_name:
                raise ValueError(
                    f'Single-isotope {repr(single_isotope)} not found in TLI file')
            k = list(self.iso_name).index(single_isotope)
            mask = self.isoid == k
            self.wn, self.gf = self.wn[mask], self.gf[mask]
            self.elow, self.isoid = self.elow[mask], self.isoid[mask]
            self.iso_ratio[:] = 0.0
            self.iso_ratio[k] = 1.0

        self.ntransitions = len(self.wn)
        self.species = np.unique(mol_names)
        self.nspec = len(self.species)
        self.mol_index = [species.index(mol) for mol in self.species]
        self.iso_mol_index = np.array(
            [list(self.species).index(species[i]) for i in self.iso_atm_index])
        self.iso_name = np.array(self.iso_name)
        self.iso_pf = None

    def partition(self, temperature):
        """Z_i(T) [niso, len(T)] (line_by_line.py:219-222 and pyrat/extinction.py:90-92)."""
        temperature = np.atleast_1d(temperature)
        z = np.zeros((self.niso, len(temperature)), np.double)
        for sl, interp in self._db_pf_interp:
            z[sl] = interp(temperature)
        return z

    def calc_extinction_coefficient(self, temperature, density, layer=None, skip_mol=[]):
        """On-the-fly LBL extinction (line_by_line.py:200-248).

        Quirk kept from the reference: `temperature` only feeds Z(T); the layer
        temperatures and densities come from pyrat.atm (pyrat/extinction.py:186-188).
        With `layer`, returns per-species cross sections times density[layer]; otherwise
        fills and returns self.ec [nlayers, nwave] (cm-1)."""
        self.iso_pf = self.partition(temperature)
        if layer is not None:
            ec = ex.extinction(self.pyrat, [layer], grid=False, add=False)
            dens = np.atleast_1d(density[layer])
            for i in range(self.nspec):
                # the reference multiplies by density[layer] (one entry per LBL species)
                ec[i] *= dens[i] if len(dens) == self.nspec else dens
            return ec
        # The reference zeroes self.ec and lets the C code accumulate into it; the batched
        # call overwrites every element of self.ec (zeros where nothing contributes), so the
        # 8*nlayers*nwave-byte host memset is not needed.
        ex.extinction(self.pyrat, np.arange(self.nlayers), grid=False, add=True,
                      skip_mol=skip_mol)
        return self.ec

    def __str__(self):
        """Same text as the reference's Line_By_Line.__str__ (line_by_line.py:251-295)."""
        from .tools import Formatted_Write
        fw = Formatted_Write()
        fw.write('Line-transition information:')
        fw.write('Input TLI files (tlifile): {}', self.tlifile)
        fw.write('Number of databases (ndb): {:d}', self.ndb)
        for db in self.db:
            fw.write('\n' + str(db))
        fw.write(
            '\nTotal number of line transitions (ntransitions): {:,d}\n'
            'Minimum and maximum temperatures (tmin, tmax): [{:.1f}, {:.1f}] K',
            self.ntransitions, self.tmin, self.tmax)
        fw.write('Line-transition isotope IDs (isoid):\n    {}', self.isoid, edge=7)
        fw.write('Line-transition wavenumbers (wn, cm-1):\n    {}', self.wn,
                 fmt={'float': '{:.3f}'.format}, edge=3)
        fw.write('Line-transition lower-state energy (elow, cm-1):\n    {}', self.elow,
                 fmt={'float': '{: .3e}'.format}, edge=3)
        fw.write('Line-transition gf (gf, cm-1):\n    {}', self.gf,
                 fmt={'float': '{: .3e}'.format}, edge=3)
        fw.write('Line-transition strength threshold (ethresh): {:.2e}', self.ethresh)
        fw.write('Isotopes information:')
        fw.write('Number of isotopes (niso): {:d}', self.niso)
        fw.write(
            '\nIsotope  Molecule      Mass    Isotopic   Database'
            '\n            index     g/mol       ratio'
            '\n (name)    (imol)    (mass)     (ratio)')
        iso_index = db_index = 0
        for i in range(self.niso):
            fw.write('{:>7s}  {:8d}  {:8.4f}   {:.3e}   {}', self.iso_name[i],
                     self.iso_atm_index[i], self.iso_mass[i], self.iso_ratio[i],
                     self.db[db_index].name)
            iso_index += 1
            if iso_index == self.db[db_index].niso:
                db_index += 1
                iso_index = 0
        return fw.text
