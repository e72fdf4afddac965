"""Worker of tests/test_gpu_multi.py: builds a small cross-section table with
`Pyrat.compute_opacity` on WORLD_SIZE ranks (one per GPU, NCCL) and checks that every rank ends
with the complete table on its device, equal bit for bit to the rows a single engine computes.
Launched with torch.distributed.run; prints 'OK rank <r>' per rank."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from pyratbay_b200 import constants as pc, tli as ptli, workloads
    from pyratbay_b200.pyrat import Pyrat
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tmp = sys.argv[1]
    w = workloads.table_workload(200_000, ntemp=5, nlayers=9, nwave=3000, wl_low_um=1.0,
                                 wl_high_um=1.3)
    path = os.path.join(tmp, "lines.tli")
    if rank == 0:
        wn, elow, gf, iso, counts = w.make_lines()
        ptli.write_tli(path, [w.db], [{"wn": wn, "elow": elow, "gf": gf, "iso_id": iso,
                                       "n_lines_iso": counts}],
                       w.inputs["wnlow"], w.inputs["wnhigh"])
    dist.barrier()
    cs = os.path.join(tmp, "table.npz")
    pyrat = Pyrat(dict(w.inputs, tlifile=[path], sampled_cs=[cs]), atm=w.atm, device=local)
    pyrat.compute_opacity(host="all", nchunks=3)
    ex = pyrat.ex
    n_units = 45
    itemp, ilayer = np.arange(n_units) // 9, np.arange(n_units) % 9
    temps = ex.temp[itemp]
    dens = w.atm.vmr[ilayer] * w.atm.press[ilayer, None] * pc.bar / (pc.k * temps[:, None])
    direct = pyrat.engine.extinction_batch(temps, dens, ex.z[:, itemp].T, w.iso_mol_index, 1,
                                           1e-30, 0, 0)
    assert ex.etable_dev.is_cuda and ex.etable_dev.device.index == local
    assert np.array_equal(ex.etable_dev.cpu().numpy().reshape(n_units, -1), direct[:, 0])
    assert np.array_equal(ex.etable.reshape(n_units, -1), direct[:, 0])
    assert len(ex._assembler.mine) < n_units      # really sharded
    dist.barrier()
    if rank == 0:
        from pyratbay_b200 import io
        assert np.array_equal(io.read_opacity(cs, extract="opacity"), ex.etable)
    print(f"OK rank {rank}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
