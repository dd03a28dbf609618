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


# ---------------------------------------------------------------------------------------------
# atom-sharded store, frame exchange, rank-0 gating of persistent state
# ---------------------------------------------------------------------------------------------
def _sharded_worker(rank, world, port, out_dir):
    import torch.distributed as dist

    from lammps_analysis_b200 import distributed as D
    from lammps_analysis_b200.file_io import ScriptInput
    from lammps_analysis_b200.project import Project

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)            # the same source data on every rank
        T = 9
        data = {"Na": {"Positions": rng.random((11, T, 3)).astype(np.float32)},
                "Cl": {"Positions": rng.random((1, T, 3)).astype(np.float32)}}
        for persist in (False, True):
            project = Project("shard", storage_path=os.path.join(out_dir, "proj"),
                              persist=persist)
            assert project.sharded
            exp = project.add_experiment("X", timestep=1.0, temperature=1.0, units="metal")
            exp.add_data(ScriptInput(data, [4.0, 4.0, 4.0], atom_major=True))
            st = exp.store
            for sp, n in (("Na", 11), ("Cl", 1)):
                path = f"{sp}/Positions"
                assert st.shape(path) == (n, T, 3)
                lo, hi = st.owned_rows(path)
                assert (lo, hi) == D.shard_atoms(0, n, rank, world)
                assert np.array_equal(st.host(path), data[sp]["Positions"][lo:hi])
                # host reads of a sharded dataset are collective and return the whole array
                assert np.array_equal(st.load_data(path), data[sp]["Positions"].astype(float))
                # the frame exchange hands every rank all atoms of its frames
                frames = np.array([0, 3, 4, 8, 2])
                local = torch.from_numpy(np.ascontiguousarray(st.host(path)[:, frames]))
                full = D.exchange_frames(local, [b - a for a, b in st.rows_per_rank(path)],
                                         len(frames))
                mine = D.shard_frames(frames)
                assert np.array_equal(full.numpy(), data[sp]["Positions"][:, mine])
            # rows of another rank are refused, not silently read
            other = D.shard_atoms(0, 11, 1 - rank, world)
            if not persist:
                with pytest.raises(Exception):
                    st._local("Na/Positions", *other)
            # replicated observable; written once for a shared directory
            st.put("Observables/J", np.full((1, T, 3), 2.5))
            assert st.owned_rows("Observables/J") == (0, 1)
            assert np.array_equal(st.host("Observables/J"), np.full((1, T, 3), 2.5, np.float32))
            # appended data extends every rank's block (resize keeps the rows it owns)
            more = {s: {"Positions": rng.random((n, 4, 3)).astype(np.float32)}
                    for s, n in (("Na", 11), ("Cl", 1))}
            exp.add_data(ScriptInput(more, [4.0, 4.0, 4.0], atom_major=True, name="more"))
            lo, hi = st.owned_rows("Na/Positions")
            want = np.concatenate([data["Na"]["Positions"], more["Na"]["Positions"]], axis=1)
            assert st.shape("Na/Positions") == (11, T + 4, 3)
            assert np.array_equal(st.host("Na/Positions"), want[lo:hi])
            # result cache: rank 0 writes a shared database, every rank gets the same hit
            params = {"a": 1, "version": exp.version}
            assert project.find_computation("calc", "X", params) is None
            project.store_computation("calc", "X", params,
                                      {"Na": {"x": [float(i) for i in range(40)], "v": 1.5}},
                                      metadata={"ties": 0})
            hit = project.find_computation("calc", "X", params)
            assert hit is not None and hit["Na"]["x"][39] == 39.0 and hit.metadata == {"ties": 0}
            if persist:
                assert (project._db is not None) == (rank == 0)
                dist.barrier()
                files = os.listdir(os.path.join(out_dir, "proj", "MDSuite_Project_shard", "X",
                                                "database"))
                assert "Na__Positions.npy" in files and "index.json" in files
        open(os.path.join(out_dir, f"sharded_ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_store_exchange_and_gating(tmp_path):
    port = _free_port()
    mp.spawn(_sharded_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "sharded_ok0") and os.path.exists(tmp_path / "sharded_ok1")
