"""Host orchestration of the LBL extinction, mirroring pyratbay/pyrat/extinction.py:14-215.

The reference forks `ncpu` processes and calls the C kernel once per (T,p) index; here all
requested indices go to the GPU engine in ONE batched call (fork and CUDA do not mix), and
across GPUs the indices are sharded one process per GPU (parallel.py).
"""
import numpy as np

from . import constants as pc
from . import io
from . import parallel
from ._mem import pinned_empty


def compute_opacity(pyrat, write=True, host='rank0', nchunks=4):
    """Compute the cross-section table (cm2 molecule-1) over the (T, p, wn) grid and write
    it to `ex.sampled_cs[0]` (pyrat/extinction.py:14-126).

    write  : write the .npz file (rank 0); False leaves the table in memory only.
    host   : which ranks copy the table to host memory (`ex.etable`): 'rank0', 'all', 'none'.
    nchunks: pieces a rank's rows are computed and all-gathered in (world > 1)."""
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
    # One process per GPU.  Every rank computes its share of the (T,p) units (:108-119, where
    # the reference forks) straight into device memory; finished chunks are all-gathered over
    # NCCL while the next chunk is computed, so the whole table ends up in the HBM of every
    # rank (`ex.etable_dev`, what a device-side consumer such as Line_Sample reads) and is
    # copied to the host (`ex.etable`, the array the reference fills) only where it is written.
    import torch
    n_units = ex.ntemp * ex.nlayers
    rank, world = parallel.rank_world()
    cost = None
    if world > 1:
        # deal units to ranks by estimated cost (wide, high-pressure profiles cost more)
        idx = np.arange(n_units)
        cost = parallel.unit_costs(
            pyrat.voigt, spec, pyrat.atm, lbl.iso_atm_index, lbl.iso_mass,
            ex.temp[idx // ex.nlayers], ex.press[idx % ex.nlayers],
            pyrat.atm.vmr[idx % ex.nlayers])
    # whole temperatures per rank where the balance allows it: strengths and the dense path's
    # per-temperature set-up are then computed for ~ntemp/world + 1 temperatures per rank
    owners = parallel.unit_owners_by_temperature(ex.ntemp, ex.nlayers, world, cost)
    if write and host == 'none':
        host = 'rank0'
    to_host = host == 'all' or (host == 'rank0' and rank == 0)
    stream_out = world == 1 and to_host             # rows leave the device chunk by chunk
    want_chunks = nchunks if (world > 1 or stream_out) else 1
    asm = getattr(ex, '_assembler', None)
    if asm is None or (asm.n_units, asm.nwave, asm.world, asm.want_chunks) != \
            (n_units, ex.nwave, world, want_chunks) \
            or any(not np.array_equal(a, b) for a, b in zip(asm.owners, owners)):
        ex._assembler = None                        # release the old buffers first
        asm = ex._assembler = parallel.TableAssembler(
            n_units, ex.nwave, owners, rank, device=pyrat.device, nchunks=want_chunks,
            align=ex.nlayers if world == 1 else 1)
        asm.want_chunks = want_chunks
    ex.timing = {'strengths_ms': 0.0, 'accumulate_ms': 0.0, 'total_ms': 0.0, 'dense_ms': 0.0,
                 'dense_units': 0}
    need_etable = to_host and (getattr(ex, 'etable', None) is None
                               or ex.etable.shape != (ex.ntemp, ex.nlayers, ex.nwave))
    if need_etable and not stream_out:
        ex.etable = pinned_empty((ex.ntemp, ex.nlayers, ex.nwave))
    elif need_etable:
        ex.etable = None            # allocated by the drain thread while the first chunk computes

    # With one rank the rows are final as soon as a chunk is computed: a helper thread copies
    # them to the host and appends them to the .npz while the next chunk is on the GPU (the
    # engine call and the file write both release the GIL).
    drain = _RowDrain(pyrat, cs_file if write else None) if stream_out else None
    try:
        for c, units, ptr in asm.chunks():
            if len(units):
                extinction(pyrat, units, grid=True, add=False, out_device_ptr=ptr)
                for key in ex.timing:
                    ex.timing[key] += pyrat.last_timing[key]
            asm.publish(c)
            if drain is not None and len(units):
                drain.put(asm.local, asm.bounds[c], asm.bounds[c] + len(units))
        table = asm.finish()
    finally:
        if drain is not None:
            drain.close()
    ex.etable_dev = table.view(ex.ntemp, ex.nlayers, ex.nwave)

    if to_host and not stream_out:
        torch.from_numpy(ex.etable).copy_(ex.etable_dev)     # one D2H into pinned memory
    if write and rank == 0:
        if not stream_out:
            io.write_opacity(cs_file, ex.species, ex.temp, ex.press, ex.wn, ex.etable)
        log.head(f"Cross-section table written to file: '{cs_file}'.", indent=2)


class _RowDrain:
    """Helper thread of compute_opacity: device rows -> pinned `ex.etable` -> .npz member."""

    def __init__(self, pyrat, cs_file):
        import queue
        import threading
        import torch
        ex = pyrat.ex
        self.ex = ex
        self.cs_file = cs_file
        self.device = torch.device('cuda', pyrat.device)
        self.writer = None
        self.queue = queue.Queue()
        self.error = None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def put(self, rows_dev, lo, hi):
        self.queue.put((rows_dev, lo, hi))

    def _run(self):
        import torch
        try:
            # page-locking 0.8 GB takes a few 0.1 s: done here, behind the first chunk's compute
            ex = self.ex
            torch.cuda.set_device(self.device)
            if ex.etable is None:
                ex.etable = pinned_empty((ex.ntemp, ex.nlayers, ex.nwave))
            self.flat = torch.from_numpy(ex.etable).view(-1, ex.nwave)
            self.stream = torch.cuda.Stream(device=self.device)
            if self.cs_file is not None:
                self.writer = io.OpacityWriter(self.cs_file, ex.species, ex.temp, ex.press, ex.wn)
            while True:
                item = self.queue.get()
                if item is None:
                    return
                rows_dev, lo, hi = item
                with torch.cuda.stream(self.stream):
                    self.flat[lo:hi].copy_(rows_dev[lo:hi], non_blocking=True)
                self.stream.synchronize()
                if self.writer is not None:
                    self.writer.write(self.flat[lo:hi].numpy())
        except Exception as exc:   # surfaced by close()
            self.error = exc

    def close(self):
        self.queue.put(None)
        self.thread.join()
        if self.writer is not None:
            if self.error is None:
                self.writer.close()
            else:
                self.writer.__exit__(type(self.error), self.error, None)
        if self.error is not None:
            raise self.error


def extinction(pyrat, indices, grid=False, add=False, skip_mol=[], out_device_ptr=None):
    """Extinction coefficient for atmospheric layers or table indices
    (pyrat/extinction.py:129-215).

    grid=True : index = itemp*nlayers + ilayer; stores into pyrat.ex.etable[itemp, ilayer].
    add=True  : co-added extinction (cm-1) stored into lbl.ec[ilayer].
    out_device_ptr : device address of a [len(indices), nrows, nwave] float64 buffer; the rows
                are left there (in `indices` order) and nothing is stored on the host.
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
    if out_device_ptr is not None:
        pyrat.engine.extinction_batch(
            temp, density, iso_pf, iso_mol_indices, lbl.nspec, lbl.ethresh, add, interpolate,
            out_device_ptr=out_device_ptr)
        pyrat.last_timing = pyrat.engine.last_timing()
        return None
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
