"""world_size-2 gloo tests (CPU) of the multi-GPU exchange steps: frame sharding + histogram
all-reduce for the RDF, atom sharding + series all-reduce for MSD/ACF.  The per-rank partial
results come from the oracle (the CUDA kernels cannot run here); what is under test is the
sharding arithmetic and the reduction in lammps_analysis_b200.distributed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    from lammps_analysis_b200 import distributed as D
    from oracle import dynamics as od
    from oracle import rdf as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)            # same data on every rank
        box = np.array([12.0, 12.0, 12.0])
        pos = {"Na": (rng.random((30, 5, 3)) * 12).astype(np.float32),
               "Cl": (rng.random((26, 5, 3)) * 12).astype(np.float32)}
        cutoff = orc.default_cutoff(box)
        nbins = orc.default_number_of_bins(cutoff)
        frames = np.arange(5)
        mine = D.shard_frames(frames)
        assert D.world_size() == world and D.rank() == rank
        part = orc.rdf_counts(pos, ["Na", "Cl"], box, mine, cutoff, nbins, 10, 1) if len(mine) \
            else {k: np.zeros(nbins, dtype=np.int64) for k in ("Na_Na", "Na_Cl", "Cl_Cl")}
        hist = torch.from_numpy(np.stack([part[k] for k in ("Na_Na", "Na_Cl", "Cl_Cl")]))
        D.all_reduce_sum_([hist])
        full = orc.rdf_counts(pos, ["Na", "Cl"], box, frames, cutoff, nbins, 10, 1)
        assert np.array_equal(hist.numpy(), np.stack([full[k] for k in ("Na_Na", "Na_Cl", "Cl_Cl")]))

        A, T, N = 11, 120, 30
        x = np.cumsum(rng.normal(size=(A, T, 3)), axis=1).astype(np.float32)
        plan = dict(batch_size=T, n_batches=1, remainder=0, minibatch=False)
        lo, hi = D.shard_atoms(0, A)
        msd_part, _ = od.einstein_msd(x[lo:hi], plan, N, 1, np.arange(N))
        msd = torch.from_numpy(np.array(msd_part))
        D.all_reduce_sum_([msd, None])
        msd_full, _ = od.einstein_msd(x, plan, N, 1, np.arange(N))
        np.testing.assert_allclose(msd.numpy(), msd_full, rtol=1e-12)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_shard_helpers_cover_everything_once():
    from lammps_analysis_b200 import distributed as D

    for n, w in [(10, 3), (7, 8), (1000, 8), (0, 2)]:
        blocks = [D.shard_atoms(5, 5 + n, r, w) for r in range(w)]
        assert blocks[0][0] == 5 and blocks[-1][1] == 5 + n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1
        fr = np.arange(n)
        got = np.sort(np.concatenate([D.shard_frames(fr, r, w) for r in range(w)]))
        assert np.array_equal(got, fr)
    assert D.world_size() == 1 and D.rank() == 0


@pytest.mark.timeout(300)
def test_two_rank_reduction_equals_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
