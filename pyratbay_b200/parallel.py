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


def partition_units(n_units, rank, world, cost=None, equal_counts=False):
    """Indices of the units owned by `rank`.

    Without a cost model the split is the reference's round-robin (index % world ==
    rank), which interleaves pressures and temperatures.  With `cost` (one weight per
    unit) units are dealt greedily, heaviest first, to the least-loaded rank; with
    `equal_counts` no rank takes more than ceil(n_units / world) units, so that the rows
    of every rank fit one fixed-size slot of an all-gather."""
    return unit_owners(n_units, world, cost, equal_counts)[rank]


def unit_owners(n_units, world, cost=None, equal_counts=False):
    """The unit indices of every rank (list of `world` ascending index arrays)."""
    if world <= 1:
        return [np.arange(n_units)]
    if cost is None:
        return [np.arange(r, n_units, world) for r in range(world)]
    cost = np.asarray(cost, np.double)
    order = np.argsort(-cost, kind='stable')
    load = np.zeros(world)
    count = np.zeros(world, int)
    cap = -(-n_units // world) if equal_counts else n_units
    owner = np.empty(n_units, int)
    for u in order:
        r = int(np.argmin(np.where(count < cap, load, np.inf)))
        owner[u] = r
        load[r] += cost[u]
        count[r] += 1
    return [np.where(owner == r)[0] for r in range(world)]


def unit_owners_by_temperature(ntemp, nlayers, world, cost=None):
    """Partition of a table's units (index = itemp*nlayers + ilayer) that keeps the units of a
    temperature together: the temperatures are taken in snake order (coldest, hottest, second
    coldest, ...: a monotonic cost trend with T averages out), their units concatenated, and the
    sequence cut into `world` contiguous pieces of equal estimated cost.  A rank then needs the
    line strengths (and the dense-path set-up) of ~ntemp/world + 1 temperatures instead of all of
    them.  Returns the ascending unit indices of every rank (counts may differ by a few units)."""
    n_units = ntemp * nlayers
    if world <= 1:
        return [np.arange(n_units)]
    cost = np.ones(n_units) if cost is None else np.asarray(cost, np.double)
    lo, hi, order = 0, ntemp - 1, []
    while lo <= hi:
        order.append(lo)
        if hi != lo:
            order.append(hi)
        lo, hi = lo + 1, hi - 1
    seq = np.concatenate([t * nlayers + np.arange(nlayers) for t in order])
    cum = np.cumsum(cost[seq])
    cuts = np.searchsorted(cum, cum[-1] * np.arange(1, world) / world, side='left') + 1
    cuts = np.clip(cuts, 0, n_units)
    return [np.sort(piece) for piece in np.split(seq, cuts)]


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


class TableAssembler:
    """Device-resident assembly of a (T,p) table whose rows are computed by several ranks.

    Every rank owns `owners[rank]` (unit indices, ascending) and computes them chunk by chunk
    into its slot buffer `local` [pad, nwave] on the device; a finished chunk is all-gathered
    (NCCL over NVLink, asynchronously on a side stream) while the next chunk is computed, and
    `finish()` returns the complete table [n_units, nwave] on the device of EVERY rank.  No
    host bounce and no padding copy: the engine writes straight into the buffer the collective
    reads; slots are padded to a common row count (`equal_counts` partitions differ by at
    most one row).  With one rank the slot buffer is the table (chunks then only serve the
    caller's own pipelining: finished rows are written out while the next ones are computed).

    Reference analogue: the shared `mp.Array` that the forked workers of
    pyratbay/pyrat/extinction.py:100-122 fill (and line_sampling.py:253-275 for the consumer).
    """

    def __init__(self, n_units, nwave, owners, rank, device=None, nchunks=4, align=1):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.n_units, self.nwave = int(n_units), int(nwave)
        self.owners = [np.asarray(o, int) for o in owners]
        self.world, self.rank = len(self.owners), rank
        self.mine = self.owners[rank]
        self.backend = dist.get_backend() if self.world > 1 else None
        if device is None or (self.world > 1 and self.backend != 'nccl'):
            self.dev = torch.device('cpu')
        else:
            self.dev = torch.device('cuda', device)
        pad = max(len(o) for o in self.owners)
        nchunks = max(1, min(int(nchunks), pad))
        # chunk boundaries in slot rows, optionally on multiples of `align` (one rank: whole
        # temperatures per chunk, so that no strengths pass is computed twice)
        bounds = np.linspace(0, pad, nchunks + 1)
        bounds = np.round(bounds / align) * align if align > 1 else np.round(bounds)
        bounds[0], bounds[-1] = 0, pad
        self.bounds = sorted(set(int(b) for b in np.clip(bounds, 0, pad)))
        self.local = torch.empty((pad, self.nwave), dtype=torch.float64, device=self.dev)
        self.pending = []
        if self.world == 1:
            self.table = self.local
            return
        self.table = torch.empty((self.n_units, self.nwave), dtype=torch.float64, device=self.dev)
        self.stage, self.src, self.dst = [], [], []
        for c in range(len(self.bounds) - 1):
            lo, hi = self.bounds[c], self.bounds[c + 1]
            rows = hi - lo
            self.stage.append(torch.empty((self.world, rows, self.nwave), dtype=torch.float64,
                                          device=self.dev))
            src, dst = [], []
            for r, own in enumerate(self.owners):
                n = max(0, min(hi, len(own)) - lo)
                src.append(r * rows + np.arange(n))
                dst.append(own[lo:lo + n])
            self.src.append(torch.from_numpy(np.concatenate(src)).to(self.dev))
            self.dst.append(torch.from_numpy(np.concatenate(dst)).to(self.dev))
        if self.dev.type == 'cuda':
            self.side = torch.cuda.Stream(device=self.dev)

    def chunks(self):
        """(chunk id, my unit indices of the chunk, device address of their first row)."""
        for c in range(len(self.bounds) - 1):
            lo, hi = self.bounds[c], self.bounds[c + 1]
            units = self.mine[lo:min(hi, len(self.mine))]
            yield c, units, self.local[lo:].data_ptr() if lo < len(self.local) else 0

    def publish(self, c):
        """Chunk `c` of the slot buffer is complete (the engine call has returned): start its
        all-gather; returns immediately on the NCCL path."""
        if self.world == 1:
            return
        torch, dist = self.torch, self.dist
        lo, hi = self.bounds[c], self.bounds[c + 1]
        part = self.local[lo:hi]
        if self.backend == 'nccl':
            with torch.cuda.stream(self.side):
                work = dist.all_gather_into_tensor(
                    self.stage[c].view(self.world * (hi - lo), self.nwave), part, async_op=True)
        else:
            work = dist.all_gather(list(self.stage[c].unbind(0)), part.contiguous(),
                                   async_op=True)
        self.pending.append((c, work))

    def finish(self):
        """Wait for the collectives and scatter the gathered slots to their table rows."""
        if self.world == 1:
            return self.table
        torch = self.torch
        for c, work in self.pending:
            work.wait()
            flat = self.stage[c].view(-1, self.nwave)
            if len(self.src[c]) == flat.shape[0]:
                self.table.index_copy_(0, self.dst[c], flat)
            else:
                self.table.index_copy_(0, self.dst[c], flat.index_select(0, self.src[c]))
        self.pending = []
        if self.dev.type == 'cuda':
            torch.cuda.current_stream(self.dev).wait_stream(self.side)
        return self.table


def assemble_rows(table, mine, device=None, cost=None):
    """All-gather the rows each rank computed into every rank's host `table` [n_units, nwave]
    (`mine` = this rank's unit indices, as from partition_units with the same arguments on
    every rank).  Convenience for host arrays; the product path keeps the rows on the device
    and uses TableAssembler directly (extinction.compute_opacity)."""
    import torch
    rank, world = rank_world()
    if world == 1:
        return table
    n_units, nwave = table.shape
    owners = unit_owners(n_units, world, cost)
    asm = TableAssembler(n_units, nwave, owners, rank, device=device, nchunks=1)
    asm.local[:len(mine)] = torch.from_numpy(np.ascontiguousarray(table[mine])).to(asm.dev)
    asm.publish(0)
    table[:] = asm.finish().cpu().numpy()
    return table
