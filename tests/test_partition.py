"""Multi-GPU host logic on CPU: partitioner properties and a world_size-2 gloo run of the
N>1 path (each rank computes its band with an injected engine; rank 0 assembles)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

from rfsi_b200 import partition as part

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("nrow,ncol,ndays,world", [(10000, 10000, 365, 8), (487, 490, 366, 4), (7, 5, 30, 4),
                                                   (3, 3, 2, 8), (100, 1, 1, 3)])
def test_partition_covers_everything_exactly_once(nrow, ncol, ndays, world):
    for mode in ("auto", "rows", "days"):
        units = part.partition(nrow, ncol, ndays, world, mode)
        assert [u.rank for u in units] == list(range(world))
        # coverage on the (raster row, day) lattice: every cell exactly once
        cover = np.zeros((ndays, nrow), dtype=np.int32)
        for u in units:
            assert u.pix_begin % ncol == 0 and u.npix % ncol == 0      # whole raster rows
            cover[u.day_begin:u.day_begin + u.ndays, u.pix_begin // ncol:(u.pix_begin + u.npix) // ncol] += 1
        assert (cover == 1).all()
        sizes = [u.npix * u.ndays for u in units]
        assert max(sizes) - min(sizes) <= max(ncol * ndays, nrow * ncol)  # balanced to one row / one day


def test_auto_mode_prefers_row_bands_for_big_rasters():
    assert all(u.ndays == 365 for u in part.partition(10000, 10000, 365, 8))
    assert all(u.npix == 5 * 7 for u in part.partition(7, 5, 30, 4))       # tiny raster: split days


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    import oracle
    from rfsi_b200 import synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = synth.make_case("c1", ndays=2)
    case.obs_z[1, ::9] = np.nan
    X, y = synth.training_table(case, lambda l, o, z, n: oracle.near_obs(l, o, z, n)[:2])
    flat = synth.grow_forest_sklearn(X, y, case.feature_names, num_trees=8, n_jobs=1)

    def compute(u):  # the engine is injected: here the CPU oracle stands in for the CUDA call
        r = oracle.predict_fused(flat.as_dict(), case.feature_names, loc_xy=case.locations(u.pix_begin, u.npix),
                                 obs_xy=case.obs_xy, obs_z=case.obs_z[u.day_begin:u.day_begin + u.ndays],
                                 n_obs=case.n_obs)
        return r["mean"]

    unit, local = part.run_partitioned(compute, case.nrow, case.ncol, 2, rank, world)
    full = part.gather_to_root(unit, local, case.nrow, case.ncol, 2)
    if rank == 0:
        ref = compute(part.WorkUnit(0, 0, case.npix, 0, 2))
        q.put(bool(np.array_equal(full, ref)))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_bands_reassemble_bit_exactly():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
