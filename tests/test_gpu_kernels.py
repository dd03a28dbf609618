"""Parity of the sm_100a kernels (through the C ABI) against the NumPy oracle.

Bars (BASELINE.json north_star): RDF bin counts bit-exact; MSD / ACF series within 1e-5
relative; unwrapped positions bit-exact (fp32 store); ionic current within fp32 storage
rounding of the fp64 sum.
"""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north_star tolerance for floating-point series


def _nacl(n_atoms, n_frames, box, seed):
    from lammps_analysis_b200.synthetic import nacl_trajectory

    return nacl_trajectory(n_atoms, n_frames, box, seed)


def _oracle_counts(data, species, box, frames, cutoff, nbins, minibatch, n_batches):
    from oracle import rdf as orc

    pos = {s: data[s]["Positions"] for s in species}
    return orc.rdf_counts(pos, species, box, np.asarray(frames), cutoff, nbins, minibatch,
                          n_batches)


@pytest.mark.parametrize("n_atoms,n_frames,box", [(216, 6, 20.0), (1000, 4, 32.0)])
def test_rdf_counts_bit_exact(cuda, n_atoms, n_frames, box):
    import torch
    from lammps_analysis_b200.engine import RdfEngine, to_device_f32
    from oracle import rdf as orc

    data, box_arr = _nacl(n_atoms, n_frames, box, seed=1)
    species = ["Na", "Cl"]
    cutoff = orc.default_cutoff(box_arr)
    nbins = orc.default_number_of_bins(cutoff)
    frames = np.arange(n_frames)
    ref = _oracle_counts(data, species, box_arr, frames, cutoff, nbins, 100, n_frames)

    eng = RdfEngine([data[s]["Positions"].shape[0] for s in species], box_arr, cutoff, nbins,
                    device=cuda)
    trajs = [to_device_f32(data[s]["Positions"], cuda) for s in species]
    eng.add_frames(trajs, frames)
    got = eng.counts()
    keys = [f"{species[a]}_{species[b]}"
            for a, b in itertools.combinations_with_replacement(range(2), 2)]
    for p, key in enumerate(keys):
        mism = int(np.count_nonzero(got[p] != ref[key]))
        assert mism == 0, f"{key}: {mism} bins differ"
        assert got[p].sum() == ref[key].sum()
    # Q1: pair totals per frame are bounded by (n-1)(n-2)/2 and (n_a-1)(n_b-1)
    assert got.sum() > 0


@pytest.mark.parametrize("tuning", [0x100, 0x200, 0x300, 0x400, 0x3100, 0x3400, 0x5100, 0x6100,
                                    0x6400, 0x8400, 0x9400])
def test_rdf_kernel_variants_agree(cuda, tuning):
    """Every tile configuration / atomic mode yields the same integers."""
    from lammps_analysis_b200.engine import RdfEngine, to_device_f32
    from oracle import rdf as orc

    data, box_arr = _nacl(1000, 2, 32.0, seed=7)
    species = ["Na", "Cl"]
    cutoff = orc.default_cutoff(box_arr)
    nbins = orc.default_number_of_bins(cutoff)
    trajs = [to_device_f32(data[s]["Positions"], cuda) for s in species]
    base = RdfEngine([500, 500], box_arr, cutoff, nbins, device=cuda)
    base.add_frames(trajs, [0, 1])
    var = RdfEngine([500, 500], box_arr, cutoff, nbins, device=cuda)
    var.add_frames(trajs, [0, 1], tuning=tuning)
    assert np.array_equal(base.counts(), var.counts())


def test_rdf_tie_census_matches_numpy(cuda):
    """mdk_rdf_tie_count: in-cutoff pairs of the sampled rows and the number whose reference
    bin (double-step rule) differs from plain fp32 binning, against the same two rules in
    NumPy.  The engine's packed frame drops the first atom of each species (Q1)."""
    import torch
    from lammps_analysis_b200 import kernels as K
    from lammps_analysis_b200.engine import RdfEngine, to_device_f32
    from oracle import rdf as orc

    F32 = np.float32
    data, box_arr = _nacl(1000, 1, 32.0, seed=11)
    species = ["Na", "Cl"]
    cutoff = orc.default_cutoff(box_arr)
    nbins = 9000                      # fine bins: a few dozen ties among ~2e5 sampled pairs
    eng = RdfEngine([500, 500], box_arr, cutoff, nbins, device=cuda)
    eng.tie_rows = 300
    eng.add_frames([to_device_f32(data[s]["Positions"], cuda) for s in species], [0])
    rep = eng.tie_report()
    # the same census on the host: packed order = Na[1:], Cl[1:]
    P = np.concatenate([data[s]["Positions"][1:, 0, :] for s in species]).astype(F32)
    box = np.asarray(box_arr, dtype=F32)
    i, j = np.triu_indices(len(P), k=1)
    keep = i < 300
    i, j = i[keep], j[keep]
    r = P[j] - P[i]
    r = r - np.rint(r / box) * box
    sq = r * r
    d = np.sqrt((sq[:, 0] + sq[:, 1]) + sq[:, 2])
    d = d[d < F32(cutoff)]
    step = float(F32(cutoff)) / float(nbins)
    k_ref = np.minimum(d.astype(np.float64) / step, nbins - 1).astype(np.int32)
    inv_step = F32(float(nbins) / float(F32(cutoff)))
    k_f32 = np.minimum(np.floor(d * inv_step), nbins - 1).astype(np.int32)
    assert rep["pairs_checked"] == len(d)
    assert rep["ties"] == int(np.count_nonzero(k_ref != k_f32))
    assert 0 < rep["ties"] < len(d) // 100
    # a second batch does not repeat the census
    eng.add_frames([to_device_f32(data[s]["Positions"], cuda) for s in species], [0])
    assert eng.tie_report()["pairs_checked"] == len(d)


def test_rdf_schedule_stress(cuda):
    """The default kernel on sorted frames (AM 7) hands column tiles through a two-stage
    mbarrier ring in which the last warp to leave a stage refills it; compute-sanitizer's
    racecheck is not available on this GPU pool, so the schedule is stressed instead: 200
    launches on one 60,000-atom frame with random work-item sizes (column chunk) and persistent
    grid sizes -- every interleaving must give the same integers, which the first launch has
    in common with the plain all-pairs kernel."""
    import torch
    from lammps_analysis_b200.engine import RdfEngine
    from lammps_analysis_b200.synthetic import device_fluid

    n, L = 60_000, 106.0
    traj = device_fluid(n, 1, L, 21, cuda)
    cutoff = L / 2 - 0.1
    nbins = int(cutoff / 0.01)
    plain = RdfEngine([n], [L] * 3, cutoff, nbins, device=cuda, spatial_sort=False)
    plain.add_frames([traj], [0])
    want = plain.counts()
    eng = RdfEngine([n], [L] * 3, cutoff, nbins, device=cuda, spatial_sort=True)
    eng.tie_rows = 0
    buf = eng._buffer(1)
    eng.pack_sorted([traj], [0], buf)
    bbox = eng.boxes(buf, 1)
    rng = np.random.default_rng(99)
    for it in range(200):
        tuning = 0x8400 | (int(rng.integers(1, 8)) << 16) | (int(rng.integers(1, 16)) << 20)
        if it == 0:
            tuning = 0x8400
        eng.hist.zero_()
        eng.add_packed(buf, 1, tuning=tuning, bbox=bbox)
        got = eng.hist.view(1, nbins)
        if it == 0:
            assert np.array_equal(got.cpu().numpy(), want)
            first = got.clone()
        else:
            assert torch.equal(got, first), f"launch {it} (tuning {tuning:#x}) changed the counts"


def test_rdf_exact_division_mode(cuda):
    """cutoff >= L/2 forces the true-division minimum image; still bit-exact."""
    from lammps_analysis_b200.engine import RdfEngine, to_device_f32

    data, box_arr = _nacl(216, 3, 20.0, seed=3)
    species = ["Na", "Cl"]
    cutoff, nbins = 12.5, 500
    ref = _oracle_counts(data, species, box_arr, np.arange(3), cutoff, nbins, 50, 3)
    eng = RdfEngine([108, 108], box_arr, cutoff, nbins, device=cuda)
    assert eng.exact_div
    eng.add_frames([to_device_f32(data[s]["Positions"], cuda) for s in species], np.arange(3))
    got = eng.counts()
    for p, key in enumerate(["Na_Na", "Na_Cl", "Cl_Cl"]):
        assert np.array_equal(got[p], ref[key]), key


def test_rdf_bin_count_beyond_shared_memory(cuda):
    """40,000 bins do not fit the shared-memory tables: the global-memory fallback gives the
    oracle's counts (and the same kernel forced on a small bin count agrees with the default)."""
    from lammps_analysis_b200.engine import RdfEngine, to_device_f32
    from oracle import rdf as orc

    rng = np.random.default_rng(61)
    pos = {"A": (rng.random((400, 2, 3)) * 30.0).astype(np.float32),
           "B": (rng.random((300, 2, 3)) * 30.0).astype(np.float32)}
    box = np.array([30.0, 30.0, 30.0])
    cutoff, nbins = 14.9, 40_000
    ref = orc.rdf_counts(pos, ["A", "B"], box, np.arange(2), cutoff, nbins, 100, 2)
    eng = RdfEngine([400, 300], box, cutoff, nbins, device=cuda)
    trajs = [to_device_f32(pos[s], cuda) for s in ("A", "B")]
    eng.add_frames(trajs, np.arange(2))
    got = eng.counts()
    for p, key in enumerate(["A_A", "A_B", "B_B"]):
        assert np.array_equal(got[p], ref[key]), key
    small = RdfEngine([400, 300], box, cutoff, 1490, device=cuda)
    small.add_frames(trajs, np.arange(2))
    forced = RdfEngine([400, 300], box, cutoff, 1490, device=cuda)
    forced.add_frames(trajs, np.arange(2), tuning=0x5000)
    assert np.array_equal(small.counts(), forced.counts())


def test_rdf_single_species_and_empty(cuda):
    from lammps_analysis_b200.engine import RdfEngine, to_device_f32
    from oracle import rdf as orc

    rng = np.random.default_rng(5)
    pos = (rng.random((300, 2, 3)) * 18.0).astype(np.float32)
    box = np.array([18.0, 18.0, 18.0])
    cutoff, nbins = 8.9, 890
    ref = orc.rdf_counts({"1": pos}, ["1"], box, np.arange(2), cutoff, nbins, 64, 2)
    eng = RdfEngine([300], box, cutoff, nbins, device=cuda)
    eng.add_frames([to_device_f32(pos, cuda)], np.arange(2))
    assert np.array_equal(eng.counts()[0], ref["1_1"])
    # no frames -> all-zero histogram, no launch error
    eng2 = RdfEngine([300], box, cutoff, nbins, device=cuda)
    eng2.add_frames([to_device_f32(pos, cuda)], np.arange(0))
    assert eng2.counts().sum() == 0


@pytest.mark.parametrize("tuning", [0, 0x3000, 0x6000, 0x8400])
def test_rdf_sorted_culled_matches_oracle(cuda, tuning):
    """Morton-ordered pack + block culling: same integers as the oracle (two species, so the
    cross-species tiles and the diagonal tiles are both exercised)."""
    from lammps_analysis_b200.engine import RdfEngine, to_device_f32
    from oracle import rdf as orc

    rng = np.random.default_rng(31)
    n, L = 5000, 58.0
    pos = {"A": (rng.random((n, 1, 3)) * L).astype(np.float32),
           "B": (rng.random((n + 37, 1, 3)) * L).astype(np.float32)}
    box = np.array([L, L, L])
    cutoff, nbins = 14.0, 1400          # cutoff << L/2: many blocks are out of range
    ref = orc.rdf_counts(pos, ["A", "B"], box, np.arange(1), cutoff, nbins, 100, 1)
    eng = RdfEngine([n, n + 37], box, cutoff, nbins, device=cuda, spatial_sort=True)
    assert eng.spatial_sort
    eng.add_frames([to_device_f32(pos[s], cuda) for s in ("A", "B")], np.arange(1), tuning=tuning)
    got = eng.counts()
    for p, key in enumerate(["A_A", "A_B", "B_B"]):
        assert np.array_equal(got[p], ref[key]), key


@pytest.mark.parametrize("tuning", [0, 0x3000, 0x6000, 0x8400])
@pytest.mark.parametrize("cutoff_frac", [0.12, 0.3, 0.4999])
def test_rdf_culling_is_exact_at_scale(cuda, cutoff_frac, tuning):
    """Size-independent property: the culled, spatially sorted pass returns exactly the
    histogram of the plain all-pairs pass (100,000 atoms, 5e9 pairs per frame)."""
    import torch
    from lammps_analysis_b200.engine import RdfEngine
    from lammps_analysis_b200.synthetic import device_fluid

    n, L = 100_000, 170.0
    traj = device_fluid(n, 2, L, 41, cuda)
    cutoff = cutoff_frac * L
    nbins = 2000
    plain = RdfEngine([n], [L] * 3, cutoff, nbins, device=cuda, spatial_sort=False)
    plain.add_frames([traj], np.arange(2))
    culled = RdfEngine([n], [L] * 3, cutoff, nbins, device=cuda, spatial_sort=True)
    culled.add_frames([traj], np.arange(2), tuning=tuning)
    a, b = plain.counts(), culled.counts()
    assert a.sum() > 0 and np.array_equal(a, b)


@pytest.mark.parametrize("wrapped", [True, False])
def test_rdf_uniform_image_two_species_orthorhombic(cuda, wrapped):
    """Size-independent property at a scale where the uniform-image kernel (AM 7) is the
    default: two species of unequal, ragged size in a non-cubic box -- the sorted, culled pass
    returns exactly the histogram of the plain all-pairs pass (all three species pairs, diagonal
    tiles included), for wrapped coordinates and for coordinates spread over two box images."""
    import torch
    from lammps_analysis_b200.engine import RdfEngine

    gen = torch.Generator(device=cuda)
    gen.manual_seed(77)
    box = np.array([150.0, 131.0, 118.5])
    counts = [70_001, 40_963]
    lo, span = (0.0, 1.0) if wrapped else (-0.7, 2.1)
    trajs = [((torch.rand(n, 1, 3, device=cuda, generator=gen) * span + lo)
              * torch.tensor(box, dtype=torch.float32, device=cuda)).contiguous() for n in counts]
    cutoff, nbins = 58.0, 5800
    plain = RdfEngine(counts, box, cutoff, nbins, device=cuda, spatial_sort=False)
    plain.add_frames(trajs, np.arange(1))
    culled = RdfEngine(counts, box, cutoff, nbins, device=cuda, spatial_sort=True)
    culled.add_frames(trajs, np.arange(1))
    a, b = plain.counts(), culled.counts()
    assert a.shape[0] == 3 and all(a[p].sum() > 0 for p in range(3))
    assert np.array_equal(a, b)


@pytest.mark.parametrize("tuning", [0, 0x3000, 0x8400])
def test_rdf_culling_with_unwrapped_coordinates(cuda, tuning):
    """Coordinates spread over several box images: the box test must stay conservative, and the
    uniform-image blocks (0x8400) must pick up shifts of one and two box lengths."""
    from lammps_analysis_b200.engine import RdfEngine, to_device_f32

    rng = np.random.default_rng(32)
    n, L = 20000, 40.0
    pos = ((rng.random((n, 1, 3)) * 2.3 - 0.6) * L).astype(np.float32)  # spans 2.3 box lengths
    cutoff, nbins = 9.0, 900
    plain = RdfEngine([n], [L] * 3, cutoff, nbins, device=cuda, spatial_sort=False)
    plain.add_frames([to_device_f32(pos, cuda)], np.arange(1))
    culled = RdfEngine([n], [L] * 3, cutoff, nbins, device=cuda, spatial_sort=True)
    culled.add_frames([to_device_f32(pos, cuda)], np.arange(1), tuning=tuning)
    assert np.array_equal(plain.counts(), culled.counts())


def _plan(A, T, data_range, ct, memory=60e9, scale=150):
    from oracle.planner import ArrayDatabase, plan_trajectory_calculator

    class _Shape:
        def __init__(self, shape):
            self.shape = shape

    db = ArrayDatabase({"x": _Shape((A, T, 3))})
    return plan_trajectory_calculator(db, ["x"], data_range, ct,
                                      {"linear": {"scale_factor": scale}}, memory)


@pytest.mark.parametrize("A,T,N,ct,memory", [
    (64, 400, 50, 1, 60e9),       # one batch
    (64, 400, 50, 3, 60e9),       # correlation_time > 1
    (60, 1000, 100, 1, 2.0e5 * 60 / 64),  # several batches + remainder
    (9, 1500, 700, 1, 60e9),      # more lags than one lag block of the dense kernel (576)
    (5, 700, 577, 1, 60e9),       # one lag into the second lag block; few windows
    (11, 900, 300, 1, 60e9),      # two-atom kernel with 5 lags per thread (odd atom count)
    (7, 1000, 400, 1, 60e9),      # ... with 7 lags per thread
])
def test_msd_matches_oracle(cuda, A, T, N, ct, memory):
    from lammps_analysis_b200.engine import msd_series, plan_windows, to_device_f32
    from oracle import dynamics as od

    rng = np.random.default_rng(11)
    x = np.cumsum(rng.normal(0, 0.1, size=(A, T, 3)), axis=1).astype(np.float32)
    plan = _plan(A, T, N, ct, memory)
    tau = np.arange(N)
    ref, ref_count = od.einstein_msd(x, plan, N, ct, tau)
    launches = plan_windows(plan, N, ct, A)
    got, count = msd_series(to_device_f32(x, cuda), launches, N, ct, tau)
    assert count == ref_count
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=RTOL, atol=0)


@pytest.mark.parametrize("A,T,N,memory", [
    (37, 1000, 1, 60e9), (37, 1000, 5, 60e9), (21, 1301, 16, 60e9), (21, 1300, 32, 60e9),
    (9, 130, 32, 60e9),                      # fewer windows than one 128-frame chunk
    (30, 999, 8, 2.0e5 * 30 / 64),           # several batches (t0 != 0, unaligned rows)
])
def test_msd_short_lag_streaming_kernel(cuda, A, T, N, memory):
    """n_lags <= 32 takes the HBM-streaming kernel (one warp per atom, 128-frame chunks)."""
    from lammps_analysis_b200.engine import msd_series, plan_windows, to_device_f32
    from oracle import dynamics as od

    rng = np.random.default_rng(15)
    x = np.cumsum(rng.normal(0, 0.1, size=(A, T, 3)), axis=1).astype(np.float32)
    plan = _plan(A, T, N, 1, memory)
    tau = np.arange(N)
    ref, ref_count = od.einstein_msd(x, plan, N, 1, tau)
    got, count = msd_series(to_device_f32(x, cuda), plan_windows(plan, N, 1, A), N, 1, tau)
    assert count == ref_count
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=RTOL, atol=0)


def test_msd_sparse_tau(cuda):
    from lammps_analysis_b200.engine import msd_series, plan_windows, to_device_f32
    from oracle import dynamics as od

    rng = np.random.default_rng(12)
    A, T, N = 32, 300, 80
    x = np.cumsum(rng.normal(0, 0.1, size=(A, T, 3)), axis=1).astype(np.float32)
    plan = _plan(A, T, N, 1)
    tau = np.linspace(0, N - 1, 17, dtype=int)
    ref, ref_count = od.einstein_msd(x, plan, N, 1, tau)
    got, count = msd_series(to_device_f32(x, cuda), plan_windows(plan, N, 1, A), N, 1, tau)
    assert count == ref_count
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=RTOL, atol=0)


@pytest.mark.parametrize("A,T,N,ct", [(48, 300, 40, 1), (48, 300, 40, 4), (20, 700, 130, 1)])
def test_acf_matches_oracle(cuda, A, T, N, ct):
    from lammps_analysis_b200.engine import acf_series, plan_windows, to_device_f32
    from oracle import dynamics as od

    rng = np.random.default_rng(13)
    v = rng.normal(0, 1.0, size=(A, T, 3))
    for t in range(1, T):
        v[:, t] = 0.9 * v[:, t - 1] + 0.4 * v[:, t]
    v = v.astype(np.float32)
    plan = _plan(A, T, N, ct)
    time = np.arange(N) * 0.002
    ref_sum, ref_count, ref_sig = od.gk_diffusion_acf(v, plan, N, ct, time, 1.0, 1.0)
    launches = plan_windows(plan, N, ct, A)
    got, count, wins, sizes = acf_series(to_device_f32(v, cuda), launches, N, ct)
    assert count == ref_count
    scale = np.abs(ref_sum).max()
    np.testing.assert_allclose(got.cpu().numpy(), ref_sum, rtol=RTOL, atol=RTOL * 1e-2 * scale)
    # per-window series -> sigmas (cumulative trapezoid of the atom-mean ACF)
    from scipy.integrate import cumulative_trapezoid

    win = np.concatenate([w.cpu().numpy() for w in wins], axis=0)
    sig = cumulative_trapezoid(win / sizes[0], x=time, axis=1)
    assert sig.shape == ref_sig.shape
    np.testing.assert_allclose(sig, ref_sig, rtol=RTOL, atol=RTOL * np.abs(ref_sig).max())


@pytest.mark.parametrize("A,T,N,ct,memory", [
    (33, 500, 1, 1, 60e9), (33, 500, 3, 1, 60e9), (17, 1001, 8, 2, 60e9), (17, 1000, 16, 1, 60e9),
    (600, 140, 16, 1, 60e9),                  # more atoms than one fp32 fold run, short series
    (30, 999, 8, 1, 2.0e5 * 30 / 64),         # several batches (t0 != 0, unaligned rows)
])
def test_acf_short_lag_streaming_kernel(cuda, A, T, N, ct, memory):
    """data_range <= 16 takes the HBM-streaming lag-product kernel."""
    from lammps_analysis_b200.engine import acf_series, plan_windows, to_device_f32
    from oracle import dynamics as od

    rng = np.random.default_rng(16)
    v = rng.normal(0, 1.0, size=(A, T, 3))
    for t in range(1, T):
        v[:, t] = 0.8 * v[:, t - 1] + 0.6 * v[:, t]
    v = v.astype(np.float32)
    plan = _plan(A, T, N, ct, memory)
    time = np.arange(N) * 0.002
    ref_sum, ref_count, ref_sig = od.gk_diffusion_acf(v, plan, N, ct, time, 1.0, 1.0)
    got, count, wins, sizes = acf_series(to_device_f32(v, cuda), plan_windows(plan, N, ct, A), N, ct)
    assert count == ref_count
    scale = np.abs(ref_sum).max()
    np.testing.assert_allclose(got.cpu().numpy(), ref_sum, rtol=RTOL, atol=RTOL * 1e-2 * scale)
    if N > 1:
        from scipy.integrate import cumulative_trapezoid

        win = np.concatenate([w.cpu().numpy() for w in wins], axis=0)
        sig = cumulative_trapezoid(win / sizes[0], x=time, axis=1)
        np.testing.assert_allclose(sig, ref_sig, rtol=RTOL, atol=RTOL * np.abs(ref_sig).max())


@pytest.mark.parametrize("box", [[10.0, 11.5, 9.25],            # exactly representable in fp32
                                 [10.1, 11.37, 9.2123456789]])    # fp64 path
def test_unwrap_bit_exact_with_carry(cuda, box):
    import torch
    from lammps_analysis_b200 import kernels as K
    from lammps_analysis_b200.engine import to_device_f32
    from oracle import transformations as ot

    rng = np.random.default_rng(21)
    A, T = 70, 333
    box = np.array(box)
    walk = np.cumsum(rng.normal(0, 1.5, size=(A, T, 3)), axis=1) + rng.random((A, 1, 3)) * box
    pos = np.mod(walk, box).astype(np.float32)
    ref = ot.run_unwrap(pos, box, batch_size=100)  # 3 batches + remainder, carry-over
    dev = to_device_f32(pos, cuda)
    out = torch.empty_like(dev)
    carry_pos = torch.zeros(A, 3, dtype=torch.float32, device=cuda)
    carry_img = torch.zeros(A, 3, dtype=torch.float64, device=cuda)
    have = False
    for lo in range(0, T, 100):
        hi = min(T, lo + 100)
        chunk = dev[:, lo:hi].contiguous()
        o = torch.empty_like(chunk)
        K.unwrap(chunk, box, carry_pos, carry_img, have, o)
        out[:, lo:hi] = o
        have = True
    got = out.cpu().numpy()
    assert np.array_equal(got, ref)
    # one batch over the whole series gives the same answer
    out2 = torch.empty_like(dev)
    carry_img.zero_()
    K.unwrap(dev, box, carry_pos, carry_img, False, out2)
    assert np.array_equal(out2.cpu().numpy(), ref)


@pytest.mark.parametrize("T", [4, 128, 200, 260])
def test_unwrap_vector_path(cuda, T):
    """T % 4 == 0 takes the 16-byte vector path; chunk boundaries at 128 frames."""
    import torch
    from lammps_analysis_b200 import kernels as K
    from lammps_analysis_b200.engine import to_device_f32
    from oracle import transformations as ot

    rng = np.random.default_rng(24)
    A = 37
    box = np.array([7.7, 8.0, 9.31])
    walk = np.cumsum(rng.normal(0, 2.5, size=(A, T, 3)), axis=1) + rng.random((A, 1, 3)) * box
    pos = np.mod(walk, box).astype(np.float32)
    ref = ot.run_unwrap(pos, box, batch_size=T)
    dev = to_device_f32(pos, cuda)
    out = torch.empty_like(dev)
    K.unwrap(dev, box, None, torch.zeros(A, 3, dtype=torch.float64, device=cuda), False, out)
    assert np.array_equal(out.cpu().numpy(), ref)


def test_unwrap_reference_known_answer(cuda):
    """CI/unit_tests/transformations/test_transformations.py:147-189 (unwrap with carry)."""
    import json
    import os
    import torch
    from lammps_analysis_b200 import kernels as K
    from lammps_analysis_b200.engine import to_device_f32

    path = os.path.join(os.path.dirname(__file__), "golden", "unwrap_carry.json")
    g = json.load(open(path))
    pos = np.asarray(g["pos"], dtype=np.float32)
    box = np.asarray(g["box"], dtype=np.float64)
    dev = to_device_f32(pos, cuda)
    out = torch.empty_like(dev)
    carry_pos = to_device_f32(np.asarray(g["last_pos"], dtype=np.float32), cuda)
    carry_img = torch.as_tensor(np.asarray(g["last_image_box"], dtype=np.float64), device=cuda)
    K.unwrap(dev, box, carry_pos, carry_img, True, out)
    np.testing.assert_allclose(out.cpu().numpy(), np.asarray(g["expected"]), rtol=0, atol=1e-6)
    np.testing.assert_allclose(carry_img.cpu().numpy(), np.asarray(g["expected_image_box"]))


def test_unwrap_single_frame_excursions(cuda):
    """An atom that leaves through a face and is back the next frame: the two jumps cancel in the
    running sum of the lane that owns both frames (and in the chunk total), but the frame in
    between carries another image.  Excursions at every frame offset of a lane, across lane and
    chunk boundaries, in both directions, for an fp32-exact and an inexact box length."""
    import torch
    from lammps_analysis_b200 import kernels as K
    from lammps_analysis_b200.engine import to_device_f32
    from oracle import transformations as ot

    T = 300
    for L in (32.0, 31.7):
        box = np.array([L, L * 1.25, L * 0.75])
        rng = np.random.default_rng(9)
        pos = np.empty((40, T, 3), dtype=np.float32)
        pos[:] = (rng.random((40, 1, 3)) * 0.2 + 0.05) * box       # just inside the lower faces
        pos += (rng.random((40, T, 3)) * 0.01 * box).astype(np.float32)
        for a in range(40):
            t = 1 + 3 * a + (a % 7)                                 # every offset mod 4, 127/128, ...
            d = a % 3
            sign = 1 if a % 2 else -1
            pos[a, t, d] = (pos[a, t, d] - sign * 0.1 * box[d]) % box[d] if sign > 0 \
                else pos[a, t, d]                                   # out through the lower face
            if sign < 0:
                pos[a, :, d] = box[d] - pos[a, :, d]                # sit near the upper face ...
                pos[a, t, d] = 0.03 * box[d]                        # ... and hop over it once
            if a % 5 == 0 and t + 130 < T:
                pos[a, t + 127:t + 129, d] = pos[a, t, d]           # a two-frame stay across a chunk edge
        want = ot.run_unwrap(pos, box, batch_size=T)
        dev = to_device_f32(pos, cuda)
        out = torch.empty_like(dev)
        K.unwrap(dev, box, None, torch.zeros(40, 3, dtype=torch.float64, device=cuda), False, out)
        assert np.array_equal(out.cpu().numpy(), want), f"box {L}"


def test_unwrap_indices(cuda):
    import torch
    from lammps_analysis_b200 import kernels as K
    from lammps_analysis_b200.engine import to_device_f32
    from oracle import transformations as ot

    rng = np.random.default_rng(22)
    pos = (rng.random((40, 50, 3)) * 7.3).astype(np.float32)
    img = rng.integers(-3, 4, size=(40, 50, 3)).astype(np.float32)
    box = np.array([7.3, 7.3, 8.1])
    ref = ot.unwrap_via_indices_transform_batch(pos, img, box).astype(np.float32)
    out = torch.empty(40, 50, 3, dtype=torch.float32, device=cuda)
    K.unwrap_indices(to_device_f32(pos, cuda), to_device_f32(img, cuda), box, out)
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("mode", ["scalar", "per_atom", "per_atom_frame"])
def test_ionic_current(cuda, mode):
    import torch
    from lammps_analysis_b200 import kernels as K
    from lammps_analysis_b200.engine import to_device_f32
    from oracle import transformations as ot

    rng = np.random.default_rng(23)
    A, T = 123, 257
    v_na = rng.normal(size=(A, T, 3)).astype(np.float32)
    v_cl = rng.normal(size=(A + 7, T, 3)).astype(np.float32)
    if mode == "scalar":
        q_na, q_cl = 1.0, -1.0
        d_na, d_cl = q_na, q_cl
    elif mode == "per_atom":
        q_na = rng.normal(size=(A, 1, 1)).astype(np.float32)
        q_cl = rng.normal(size=(A + 7, 1, 1)).astype(np.float32)
        d_na, d_cl = to_device_f32(q_na.ravel(), cuda), to_device_f32(q_cl.ravel(), cuda)
    else:
        q_na = rng.normal(size=(A, T, 1)).astype(np.float32)
        q_cl = rng.normal(size=(A + 7, T, 1)).astype(np.float32)
        d_na, d_cl = to_device_f32(q_na[..., 0], cuda), to_device_f32(q_cl[..., 0], cuda)
    ref = ot.ionic_current_transform_batch({
        "Na": {"Velocities": v_na, "Charge": np.asarray(q_na, dtype=np.float64).reshape(
            (1, 1, 1) if mode == "scalar" else q_na.shape)},
        "Cl": {"Velocities": v_cl, "Charge": np.asarray(q_cl, dtype=np.float64).reshape(
            (1, 1, 1) if mode == "scalar" else q_cl.shape)},
    })
    J = torch.zeros(T, 3, dtype=torch.float64, device=cuda)
    K.ionic_current(to_device_f32(v_na, cuda), d_na, J)
    K.ionic_current(to_device_f32(v_cl, cuda), d_cl, J)
    np.testing.assert_allclose(J.cpu().numpy(), ref, rtol=1e-12, atol=1e-11)


# ---------------------------------------------------------------------------------------------
# kernels vs vectors produced by executing the reference's own Python source
# (tests/golden/reference_run.json, generator: tests/golden/make_reference_goldens.py)
# ---------------------------------------------------------------------------------------------
def _reference_run():
    import json
    import os

    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_run.json")) as fh:
        return json.load(fh)


def test_reference_run_rdf_counts_on_gpu(cuda):
    from lammps_analysis_b200.engine import RdfEngine, to_device_f32

    g = _reference_run()["rdf"]
    pos = [np.asarray(g["positions"][s], dtype=np.float32) for s in g["species"]]
    n_frames = pos[0].shape[1]
    eng = RdfEngine([p.shape[0] for p in pos], g["box"], g["cutoff"], g["nbins"], device=cuda)
    eng.add_frames([to_device_f32(p, cuda) for p in pos], np.arange(n_frames))
    got = eng.counts()
    for p, key in enumerate(["Na_Na", "Na_Cl", "Cl_Cl"]):
        assert np.array_equal(got[p], np.asarray(g["counts"][key])), key


def test_reference_run_flux_kernels_on_gpu(cuda):
    """mdk_flux_sum / mdk_thermal_flux against vectors from the reference's own source."""
    import json
    import os

    import torch
    from lammps_analysis_b200 import kernels as K
    from lammps_analysis_b200.engine import to_device_f32

    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_flux.json")) as fh:
        g = json.load(fh)
    T = len(g["momentum_flux"])
    Jm = torch.zeros(T, 3, dtype=torch.float64, device=cuda)
    Jt = torch.zeros_like(Jm)
    Jh = torch.zeros_like(Jm)
    for sp, d in g["inputs"].items():
        dev = {k: to_device_f32(np.asarray(v, dtype=np.float32), cuda) for k, v in d.items()}
        A = dev["Stress"].shape[0]
        ke = dev["Kinetic_Energy"].reshape(A, T).contiguous()
        pe = dev["Potential_Energy"].reshape(A, T).contiguous()
        K.flux_sum(dev["Stress"], Jm, comp0=3)
        K.thermal_flux(dev["Stress"], dev["Velocities"], ke, pe, Jt)
        K.flux_sum(dev["Unwrapped_Positions"], Jh, comp0=0, w1=ke, w2=pe)
    for got, key in ((Jm, "momentum_flux"), (Jt, "thermal_flux"), (Jh, "integrated_heat_current")):
        want = np.asarray(g[key])
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-13,
                                   atol=1e-14 * np.abs(want).max())


def test_reference_run_transformations_on_gpu(cuda):
    import torch
    from lammps_analysis_b200 import kernels as K
    from lammps_analysis_b200.engine import to_device_f32

    ref = _reference_run()
    g = ref["unwrap"]
    pos = np.asarray(g["positions"], dtype=np.float32)
    box = np.asarray(g["box"])
    dev = to_device_f32(pos, cuda)
    A = pos.shape[0]
    carry_pos = torch.zeros(A, 3, dtype=torch.float32, device=cuda)
    carry_img = torch.zeros(A, 3, dtype=torch.float64, device=cuda)
    out = torch.empty_like(dev)
    have = False
    for lo, hi in g["batches"]:                      # same batches, carry-over threaded through
        chunk = dev[:, lo:hi].contiguous()
        o = torch.empty_like(chunk)
        K.unwrap(chunk, box, carry_pos, carry_img, have, o)
        out[:, lo:hi] = o
        have = True
    want = np.asarray(g["unwrapped"]).astype(np.float32)   # the reference persists float32
    assert np.array_equal(out.cpu().numpy(), want)
    gi = ref["unwrap_indices"]
    o = torch.empty_like(dev)
    K.unwrap_indices(dev, to_device_f32(np.asarray(gi["images"], dtype=np.float32), cuda), box, o)
    assert np.array_equal(o.cpu().numpy(), np.asarray(gi["unwrapped"]).astype(np.float32))
    gc = ref["ionic_current"]
    T = len(gc["current"])
    J = torch.zeros(T, 3, dtype=torch.float64, device=cuda)
    M = torch.zeros(T, 3, dtype=torch.float64, device=cuda)
    for s, q in gc["charge"].items():
        v = to_device_f32(np.asarray(gc["velocities"][s], dtype=np.float32), cuda)
        K.ionic_current(v, q, J)
        K.ionic_current(v, q, M)      # the dipole moment is the same reduction on positions
    np.testing.assert_allclose(J.cpu().numpy(), np.asarray(gc["current"]), rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(M.cpu().numpy(), np.asarray(ref["dipole_moment"]["moment"]),
                               rtol=1e-13, atol=1e-14)
