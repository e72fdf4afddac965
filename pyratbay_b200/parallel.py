"""Sharding of (T,p) units over one-process-per-GPU ranks and table assembly.

The units are independent (the reference forks over them, pyrat/extinction.py:108-119), so
there is no data-path collective; a single all-gather assembles the table when every rank
needs it.  Works with `nccl` (device tensors) and `gloo` (CPU tensors, used by the tests).
"""
import numpy as np


def rank_world():
    try:
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def partition_units(n_units, rank, world, cost=None):
    """Indices of the units owned by `rank`.

    Without a cost model the split is the reference's round-robin (index % world ==
    rank), which interleaves pressures and temperatures.  With `cost` (one weight per
    unit) units are dealt greedily, heaviest first, to the least-loaded rank."""
    if world <= 1:
        return np.arange(n_units)
    if cost is None:
        return np.arange(rank, n_units, world)
    cost = np.asarray(cost, np.double)
    order = np.argsort(-cost, kind='stable')
    load = np.zeros(world)
    owner = np.empty(n_units, int)
    for u in order:
        r = int(np.argmin(load))
        owner[u] = r
        load[r] += cost[u]
    return np.where(owner == rank)[0]


def unit_costs(voigt, spec, atm, iso_atm_index, iso_mass, temps, press, vmr):
    """Relative cost of each (T,p) unit of a table for load balancing: the accumulate kernel's
    work per line grows with the number of output samples a line reaches, i.e. with the Voigt
    half-size selected by the unit's Lorentz width (constants of src_c/include/constants.h,
    formulas of _extcoeff.c:138-183).  temps/press/vmr are per unit."""
    kb, amu, c = 1.380658e-16, 1.66053886e-24, 2.99792458e10
    temps = np.asarray(temps, np.double)
    dens = np.asarray(vmr, np.double) * (np.asarray(press) * 1e6 / (1.380649e-16 * temps))[:, None]
    imol = int(np.atleast_1d(iso_atm_index)[0])
    mass = float(np.atleast_1d(iso_mass)[0])
    diam = atm.mol_radius[imol] + atm.mol_radius
    alor = (np.sqrt(2 * kb * temps / np.pi / amu) / c
            * np.sum(dens * diam**2 * np.sqrt(1 / mass + 1 / atm.mol_mass), axis=1))
    ilor = np.clip(np.searchsorted(voigt.lorentz, alor), 0, len(voigt.lorentz) - 1)
    half = np.array([np.mean(voigt.size[i][voigt.size[i] > 0]) for i in ilor])
    if voigt.cutoff > 0:
        half = np.minimum(half, voigt.cutoff / spec.ownstep)
    wnstep = spec.wn[1] - spec.wn[0] if len(spec.wn) > 1 else 1.0
    footprint = 2.0 * half * spec.ownstep / wnstep       # output samples per line
    return 15.0 + 8.0 * (1.0 + footprint / 32.0)          # warp-instruction model, DESIGN.md


def assemble_rows(table, mine, device=None, cost=None):
    """All-gather the rows each rank computed into every rank's `table` [n_units, nwave].

    `mine` are this rank's unit indices (as from partition_units with the same arguments
    on every rank).  Rows travel as one padded tensor per rank."""
    import torch
    import torch.distributed as dist
    rank, world = rank_world()
    if world == 1:
        return table
    n_units, nwave = table.shape
    counts = [len(partition_units(n_units, r, world, cost)) for r in range(world)]
    pad = max(counts)
    backend = dist.get_backend()
    dev = torch.device('cpu')
    if backend == 'nccl':
        dev = torch.device('cuda', device if device is not None else torch.cuda.current_device())
    local = torch.zeros((pad, nwave), dtype=torch.float64, device=dev)
    if len(mine):
        local[:len(mine)] = torch.from_numpy(np.ascontiguousarray(table[mine])).to(dev)
    gathered = torch.empty((world, pad, nwave), dtype=torch.float64, device=dev)
    if backend == 'nccl':
        dist.all_gather_into_tensor(gathered.view(world * pad, nwave), local)
    else:
        dist.all_gather([gathered[r] for r in range(world)], local)
    gathered = gathered.cpu().numpy()
    for r in range(world):
        idx = partition_units(n_units, r, world, cost)
        table[idx] = gathered[r, :len(idx)]
    return table
