"""Host orchestration of the LBL extinction, mirroring pyratbay/pyrat/extinction.py:14-215.

The reference forks `ncpu` processes and calls the C kernel once per (T,p) index; here all
requested indices go to the GPU engine in ONE batched call (fork and CUDA do not mix), and
across GPUs the indices are sharded one process per GPU (parallel.py).
"""
import numpy as np

from . import constants as pc
from . import io
from . import parallel
from ._mem import pinned_zeros


def compute_opacity(pyrat):
    """Compute the cross-section table (cm2 molecule-1) over the (T, p, wn) grid and write
    it to `ex.sampled_cs[0]` (pyrat/extinction.py:14-126)."""
    ex = pyrat.ex
    spec = pyrat.spec
    log = pyrat.log

    if ex.sampled_cs is None:
        log.error('Undefined output cross_section file (sampled_cross_sec) needed to '
                  'compute opacity table')
    if len(ex.sampled_cs) > 1:
        log.error('Computing opacity table, but there was more than one '
                  'output opacity file (sampled_cross_sec)')
    if ex.tmin is None:
        log.error('Undefined lower temperature boundary (tmin) needed to '
                  'compute opacity table')
    if ex.tmax is None:
        log.error('Undefined upper temperature boundary (tmax) needed to '
                  'compute opacity table')
    if ex.tstep is None:
        log.error('Undefined temperature sampling step (tstep) needed to '
                  'compute opacity table')
    if pyrat.inputs.tlifile is None:
        log.error('Undefined input TLI files (tlifile) needed to compute opacity table')

    i_lbl = pyrat.opacity.models_type.index('lbl')
    lbl = pyrat.opacity.models[i_lbl]
    if len(lbl.species) > 1:
        log.error('Cross-section files must be for a single species only, but '
                  'line-by-line data include transitions for multiple ones: '
                  f'{lbl.species}')

    cs_file = ex.sampled_cs[0]
    log.head(f"\nGenerating new cross-section table file:\n  '{cs_file}'")
    if ex.tmin < lbl.tmin:
        log.error('Requested cross-section table temperature '
                  f'(tmin={ex.tmin:.1f} K) below the lowest available TLI '
                  f'temperature ({lbl.tmin:.1f} K)')
    if ex.tmax > lbl.tmax:
        log.error('Requested cross-section table temperature '
                  f'(tmax={ex.tmax:.1f} K) above the highest available TLI '
                  f'temperature ({lbl.tmax:.1f} K)')

    # Temperature array and partition functions (:80-92)
    ex.ntemp = int((ex.tmax - ex.tmin) / ex.tstep) + 1
    ex.temp = np.linspace(ex.tmin, ex.tmin + (ex.ntemp - 1) * ex.tstep, ex.ntemp)
    ex.species = lbl.species[0]
    with np.printoptions(formatter={'float': '{:.1f}'.format}):
        log.msg(f"Temperature sample (K):\n {ex.temp}", indent=2)
    log.msg("Interpolate partition function.", indent=2)
    ex.z = lbl.partition(ex.temp)

    ex.wn = spec.wn
    ex.nwave = spec.nwave
    ex.press = pyrat.atm.press
    ex.nlayers = pyrat.atm.nlayers

    log.msg("Calculate cross-sections.", indent=2)
    ex.etable = pinned_zeros((ex.ntemp, ex.nlayers, ex.nwave))

    # One batched GPU call per rank over its share of the (T,p) indices (:108-119)
    n_units = ex.ntemp * ex.nlayers
    rank, world = parallel.rank_world()
    if world == 1:
        extinction(pyrat, np.arange(n_units), grid=True, add=False)
    else:
        # deal units to ranks by estimated cost (wide, high-pressure profiles cost more)
        idx = np.arange(n_units)
        cost = parallel.unit_costs(
            pyrat.voigt, spec, pyrat.atm, lbl.iso_atm_index, lbl.iso_mass,
            ex.temp[idx // ex.nlayers], ex.press[idx % ex.nlayers],
            pyrat.atm.vmr[idx % ex.nlayers])
        mine = parallel.partition_units(n_units, rank, world, cost)
        extinction(pyrat, mine, grid=True, add=False)
        flat = ex.etable.reshape(n_units, ex.nwave)
        parallel.assemble_rows(flat, mine, device=pyrat.device, cost=cost)

    if rank == 0:
        io.write_opacity(cs_file, ex.species, ex.temp, ex.press, ex.wn, ex.etable)
        log.head(f"Cross-section table written to file: '{cs_file}'.", indent=2)


def extinction(pyrat, indices, grid=False, add=False, skip_mol=[]):
    """Extinction coefficient for atmospheric layers or table indices
    (pyrat/extinction.py:129-215).

    grid=True : index = itemp*nlayers + ilayer; stores into pyrat.ex.etable[itemp, ilayer].
    add=True  : co-added extinction (cm-1) stored into lbl.ec[ilayer].
    otherwise : returns the per-species cross section (cm2 molecule-1) of the FIRST index
                only, as the reference does (:214-215).
    """
    atm = pyrat.atm
    spec = pyrat.spec
    i_lbl = pyrat.opacity.models_type.index('lbl')
    lbl = pyrat.opacity.models[i_lbl]
    voigt = pyrat.voigt
    log = pyrat.log
    indices = np.asarray(indices, int)
    if not grid and not add:
        indices = indices[:1]
    if len(indices) == 0:
        return None

    interpolate = spec.resolution is not None or spec.wlstep is not None
    iso_mol_indices = np.copy(lbl.iso_mol_index)
    for mol in np.intersect1d(skip_mol, lbl.species):
        mol_index = list(lbl.species).index(mol)
        iso_mol_indices[iso_mol_indices == mol_index] = -1

    ilayer = indices % atm.nlayers
    pressure = atm.press[ilayer]
    if grid:
        itemp = indices // atm.nlayers
        temp = pyrat.ex.temp[itemp]
        density = (atm.vmr[ilayer] * pressure[:, None] * pc.bar / (pc.k * temp[:, None]))
        iso_pf = pyrat.ex.z[:, itemp].T
        log.msg(f"Extinction-coefficient table: {len(indices)} (T,p) units in one batch.",
                indent=2)
    else:
        temp = atm.temp[ilayer]
        density = atm.d[ilayer]
        iso_pf = lbl.iso_pf[:, ilayer].T
        log.msg(f"Calculating extinction at {len(indices)} layer(s).", indent=2)

    nrows = 1 if add else lbl.nspec
    # Write straight into the caller-visible array when the batch covers it in order.
    out = None
    in_order = np.array_equal(indices, np.arange(len(indices)))
    if grid and in_order and nrows == 1 \
            and len(indices) == pyrat.ex.etable.shape[0] * pyrat.ex.etable.shape[1]:
        out = pyrat.ex.etable.reshape(len(indices), 1, spec.nwave)
    elif add and not grid and in_order and len(indices) == atm.nlayers:
        out = lbl.ec.reshape(atm.nlayers, 1, spec.nwave)
    result = pyrat.engine.extinction_batch(
        temp, density, iso_pf, iso_mol_indices, lbl.nspec, lbl.ethresh, add, interpolate,
        out=out)
    pyrat.last_timing = pyrat.engine.last_timing()

    if grid:
        if out is None:
            pyrat.ex.etable[itemp, ilayer] = result[:, 0]
    elif add:
        if out is None:
            lbl.ec[ilayer] = result[:, 0]
    else:
        return result[0]
