"""Parity at BASELINE.json's full sizes.

* Against the ORACLE where it can be made to finish: config C2 at full size (5,000 frames,
  data_range 500 -- the 8-band-tile shape of the ACF kernel and the two-atom MSD kernel the
  bench times) against the committed oracle output tests/golden/c2_full.json; config C3's full
  Green-Kubo window count (W = 9,500); a 512-row x 10^6-column slab of the C5 pair matrix on the
  default (Hilbert-sorted, uniform-image) RDF path.
* Through size-independent properties where it cannot: analytic trajectories for MSD / ACF,
  round trips for the unwrap, invariances of the RDF histogram.  C5 shard: 125,000 atoms x
  2,000 frames, data_range 500; C4: 100,000 atoms per frame."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

A, T, N = 125_000, 2000, 500


@pytest.fixture(scope="module")
def ballistic(cuda):
    """x_a(t) = x0_a + v_a * t with fp32-exact values: every displacement is exact."""
    import torch

    gen = torch.Generator(device=cuda)
    gen.manual_seed(7)
    v = torch.randint(-8, 9, (A, 1, 3), device=cuda, generator=gen).float() / 16.0
    x0 = torch.randint(0, 1024, (A, 1, 3), device=cuda, generator=gen).float() / 8.0
    t = torch.arange(T, device=cuda, dtype=torch.float32).view(1, T, 1)
    return (x0 + v * t).contiguous(), v


def test_msd_of_ballistic_motion_is_exact(cuda, ballistic):
    """msd_sum[k] = W * sum_a |v_a|^2 * k^2 (all terms exactly representable)."""
    from lammps_analysis_b200.engine import msd_series, plan_windows

    x, v = ballistic
    plan = dict(batch_size=T, n_batches=1, remainder=0, minibatch=False)
    launches = plan_windows(plan, N, 1, A)
    got, count = msd_series(x, launches, N, 1, np.arange(N))
    W = T - N
    assert count == W * (A + 1)
    v2 = float((v.double() ** 2).sum())
    want = W * v2 * np.arange(N, dtype=np.float64) ** 2
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-6, atol=0)


def test_acf_of_constant_velocities(cuda, ballistic):
    """v_a(t) = v_a: every window's unbiased ACF equals sum_a |v_a|^2 at every lag."""
    from lammps_analysis_b200.engine import acf_series, plan_windows

    _, v = ballistic
    vel = v.expand(A, T, 3).contiguous()
    plan = dict(batch_size=T, n_batches=1, remainder=0, minibatch=False)
    launches = plan_windows(plan, N, 1, A)
    acf, count, wins, sizes = acf_series(vel, launches, N, 1)
    W = T - N
    v2 = float((v.double() ** 2).sum())
    np.testing.assert_allclose(acf.cpu().numpy(), np.full(N, W * v2), rtol=1e-6)
    np.testing.assert_allclose(wins[0].cpu().numpy(), np.full((W, N), v2), rtol=1e-6)


def test_acf_is_linear_in_atoms(cuda):
    """ACF(atoms 0..A) = ACF(0..A/2) + ACF(A/2..A) (the sharding identity used across GPUs)."""
    import torch
    from lammps_analysis_b200.engine import acf_series, plan_windows

    gen = torch.Generator(device=cuda)
    gen.manual_seed(8)
    A2 = 20_000
    vel = torch.randn(A2, T, 3, device=cuda, generator=gen)
    plan = dict(batch_size=T, n_batches=1, remainder=0, minibatch=False)
    launches = plan_windows(plan, N, 1, A2)
    full, _, _, _ = acf_series(vel, launches, N, 1, per_window=False)
    lo, _, _, _ = acf_series(vel, launches, N, 1, per_window=False, a_shard=(0, A2 // 2))
    hi, _, _, _ = acf_series(vel, launches, N, 1, per_window=False, a_shard=(A2 // 2, A2))
    f = full.cpu().numpy()
    np.testing.assert_allclose((lo + hi).cpu().numpy(), f, rtol=5e-6, atol=1e-7 * np.abs(f).max())
    assert f[0] > 0 and np.abs(f[1:]).max() < 0.01 * f[0]   # white noise: delta-correlated


def test_unwrap_round_trip_and_idempotence(cuda):
    """wrap(unwrap(wrap(x))) == wrap(x) bit for bit on a power-of-two box, unwrap of continuous
    motion recovers it exactly, and unwrapping an already unwrapped row changes nothing."""
    import torch
    from lammps_analysis_b200 import kernels as K

    gen = torch.Generator(device=cuda)
    gen.manual_seed(9)
    L = 64.0
    steps = torch.randint(-64, 65, (A, T, 3), device=cuda, generator=gen).float() / 32.0
    start = torch.randint(0, 2048, (A, 1, 3), device=cuda, generator=gen).float() / 32.0
    walk = start + torch.cumsum(steps, dim=1) - steps[:, :1]     # exact in fp32 (multiples of 1/32)
    wrapped = torch.remainder(walk, L).contiguous()
    out = torch.empty_like(wrapped)
    img = torch.zeros(A, 3, dtype=torch.float64, device=cuda)
    K.unwrap(wrapped, [L] * 3, None, img, False, out)
    assert torch.equal(out, walk)                                  # |step| < L/2: exact recovery
    assert torch.equal(torch.remainder(out, L), wrapped)
    out2 = torch.empty_like(out)
    img.zero_()
    K.unwrap(out, [L] * 3, None, img, False, out2)
    assert torch.equal(out2, out)
    assert float(img.abs().max()) == 0.0


def test_rdf_histogram_invariances_c4(cuda):
    """100,000-atom frame: (i) the histogram does not depend on the order of the atoms,
    (ii) exchanging the x and y axes (cubic box) leaves it unchanged, (iii) two frames give the
    sum of the single-frame histograms, (iv) the total count equals the number of pairs inside
    the cutoff sphere (compared with the expectation for a uniform fluid)."""
    import torch
    from lammps_analysis_b200.engine import RdfEngine
    from lammps_analysis_b200.synthetic import device_fluid

    n, L = 100_000, 170.0
    traj = device_fluid(n, 2, L, 4, cuda)
    cutoff = L / 2 - 0.1
    nbins = int(cutoff / 0.01)

    def hist(t, frames, sort=None):
        eng = RdfEngine([n], [L] * 3, cutoff, nbins, drop_first=False, device=cuda,
                        spatial_sort=sort)
        eng.add_frames([t], frames)
        return eng.counts()[0]

    h0, h1, h01 = hist(traj, [0]), hist(traj, [1]), hist(traj, [0, 1])
    assert np.array_equal(h0 + h1, h01)
    perm = torch.randperm(n, device=cuda)
    assert np.array_equal(hist(traj[perm].contiguous(), [0]), h0)
    assert np.array_equal(hist(traj, [0], sort=True), h0)
    # d^2 = (x*x + y*y) + z*z in fp32 is symmetric under x <-> y (not under a swap with z, whose
    # square is added last): the cubic box makes the swapped system an equivalent one
    swapped = traj[:, :, [1, 0, 2]].contiguous()
    assert np.array_equal(hist(swapped, [0]), h0)
    expected = n * (n - 1) / 2 * (4 / 3 * np.pi * cutoff**3) / L**3
    assert abs(h0.sum() - expected) < 0.01 * expected


# ---------------------------------------------------------------------------------------------
# oracle parity at the timed shapes
# ---------------------------------------------------------------------------------------------
RTOL = 1e-5
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c2_full.json")


@pytest.fixture(scope="module")
def c2_full(tmp_path_factory, cuda):
    from lammps_analysis_b200.config import config
    from lammps_analysis_b200.file_io import ScriptInput
    from lammps_analysis_b200.project import Project
    from lammps_analysis_b200.synthetic import nacl_trajectory

    gold = json.load(open(GOLDEN))
    c = gold["config"]
    config.planner_memory_bytes = c["planner_memory"]
    data, box = nacl_trajectory(c["n_atoms"], c["n_frames"], c["box"], seed=c["seed"],
                                sigma_step=c["sigma_step"])
    project = Project("c2full", storage_path=str(tmp_path_factory.mktemp("c2full")))
    exp = project.add_experiment("NaCl", timestep=c["time_step"], temperature=1400.0,
                                 units=c["units"])
    exp.add_data(ScriptInput(data, box, sample_rate=c["sample_rate"], atom_major=True))
    return exp, gold


def test_c2_full_size_einstein_matches_oracle_golden(c2_full):
    """BASELINE configs[1] at full size: 1,000-atom NaCl, 5,000 frames, data_range 500 ->
    W = 4,500 windows; unwrap + msd_dense2_kernel + fit through the public call, against the
    per-window oracle (tests/golden/make_full_config_goldens.py)."""
    exp, gold = c2_full
    res = exp.run.EinsteinDiffusionCoefficients(data_range=gold["config"]["data_range"], plot=False)
    for sp in ("Na", "Cl"):
        ref = gold["einstein"][sp]
        assert ref["count"] == 4500 * 501
        unw = exp.store.host(f"{sp}/Unwrapped_Positions")
        assert float(np.asarray(unw, dtype=np.float64).sum()) == ref["unwrapped_checksum"]
        np.testing.assert_allclose(res[sp]["msd"], ref["msd"], rtol=RTOL)
        np.testing.assert_allclose(res[sp]["time"], ref["time"], rtol=1e-12)
        assert res[sp]["diffusion_coefficient"] == pytest.approx(ref["diffusion_coefficient"],
                                                                 rel=RTOL)
        assert res[sp]["uncertainty"] == pytest.approx(ref["uncertainty"], rel=1e-3)


def test_c2_full_size_green_kubo_matches_oracle_golden(c2_full):
    """Same config: acf_band_kernel on 8 band tiles (N = 500, T = 5,000) + prefix / window
    kernels, per-window integrals for the SEM, against the per-window complex128 FFT oracle."""
    exp, gold = c2_full
    res = exp.run.GreenKuboDiffusionCoefficients(data_range=gold["config"]["data_range"], plot=False)
    for sp in ("Na", "Cl"):
        ref = gold["green_kubo"][sp]
        assert ref["count"] == 4500 * 501 and ref["n_windows"] == 4500
        scale = np.abs(ref["acf"]).max()
        np.testing.assert_allclose(res[sp]["acf"], ref["acf"], rtol=RTOL, atol=1e-7 * scale)
        np.testing.assert_allclose(res[sp]["integral"], ref["integral"], rtol=RTOL,
                                   atol=1e-7 * np.abs(ref["integral"]).max())
        np.testing.assert_allclose(res[sp]["integral_uncertainty"], ref["integral_uncertainty"],
                                   rtol=1e-4)
        assert res[sp]["diffusion_coefficient"][0] == pytest.approx(ref["diffusion_coefficient"],
                                                                    rel=RTOL)
        assert res[sp]["uncertainty"][0] == pytest.approx(ref["uncertainty"], rel=1e-4)


def test_c3_full_window_count_green_kubo_ionic(tmp_path, cuda):
    """BASELINE configs[2]'s Green-Kubo leg at its full window count: a 10,000-frame ionic
    current, data_range 500 -> W = 9,500 windows, against the per-window oracle (live: one
    row, ~10 s).  The current is an Ornstein-Uhlenbeck series stored as the observable."""
    from lammps_analysis_b200.config import config
    from lammps_analysis_b200.file_io import ScriptInput
    from lammps_analysis_b200.project import Project
    from oracle import dynamics as od
    from oracle.planner import ArrayDatabase, plan_trajectory_calculator

    config.planner_memory_bytes = 60e9
    T, N = 10_000, 500
    rng = np.random.default_rng(33)
    J = np.empty((T, 3))
    J[0] = rng.normal(size=3)
    noise = rng.normal(0, np.sqrt(1 - 0.98**2), size=(T, 3))
    for t in range(1, T):
        J[t] = 0.98 * J[t - 1] + noise[t]
    J32 = J.astype(np.float32)[None]
    box = [69.0, 69.0, 69.0]
    project = Project("c3full", storage_path=str(tmp_path))
    exp = project.add_experiment("NaCl", timestep=0.002, temperature=1400.0, units="metal")
    exp.add_data(ScriptInput({"Na": {"Velocities": np.zeros((2, T, 3), np.float32)}}, box,
                             sample_rate=10, atom_major=True))
    exp.store.put("Observables/Ionic_Current", J32)
    res = exp.run.GreenKuboIonicConductivity(data_range=N, plot=False)

    class _S:
        shape = (1, T, 3)

    plan = plan_trajectory_calculator(ArrayDatabase({"x": _S()}), ["x"], N, 1,
                                      {"linear": {"scale_factor": 5}}, 60e9)
    tau, _, _, times = od.handle_tau_values(np.s_[:], N, 0.002, 10)
    acf_sum, count, sigmas = od.gk_ionic_acf(J32.astype(np.float64), plan, N, 1, tau, times)
    assert count == 9500
    pref = od.gk_ionic_prefactor(1e-10, 1e-12, 1400.0, float(np.prod(box)))
    ref = od.gk_ionic_finish(acf_sum, count, sigmas, times, pref, N - 1)
    scale = np.abs(ref["acf"]).max()
    np.testing.assert_allclose(res["System"]["acf"], ref["acf"], rtol=RTOL, atol=1e-7 * scale)
    np.testing.assert_allclose(res["System"]["integral_uncertainty"], ref["integral_uncertainty"],
                               rtol=1e-4)
    assert res["System"]["ionic_conductivity"][0] == pytest.approx(ref["ionic_conductivity"][0],
                                                                   rel=RTOL)
    assert res["System"]["uncertainty"][0] == pytest.approx(ref["uncertainty"][0], rel=1e-4)


@pytest.mark.timeout(900)
def test_rdf_million_atom_frame_against_oracle_slab(cuda):
    """C5-size frame on the DEFAULT path (Hilbert-sorted pack, block culling, uniform-image
    blocks: no tuning flag): species A = 513 atoms clustered in a corner of the box (so that
    its 32-row groups are compact and pairs cross the periodic boundary with image shifts),
    species B = 10^6 atoms.  The A x B histogram is a 512 x 999,999 slab of the pair matrix the
    oracle can evaluate (rdf_counts_slab, 5e8 distances); A x A is checked as well.  Bit-exact."""
    import torch
    from lammps_analysis_b200.engine import RdfEngine, to_device_f32
    from oracle import rdf as orc

    rng = np.random.default_rng(55)
    n_b, L = 1_000_000, (1_000_000 / 0.05) ** (1 / 3)
    box = np.array([L, L, L])
    cutoff = orc.default_cutoff(box)
    nbins = orc.default_number_of_bins(cutoff)
    assert nbins == 13562
    pos_a = (rng.random((513, 1, 3)) * 22.0).astype(np.float32)        # corner cluster
    pos_b = (rng.random((n_b, 1, 3)) * L).astype(np.float32)
    pos_b[pos_b >= np.float32(L)] = 0.0
    eng = RdfEngine([513, n_b], box, cutoff, nbins, device=cuda)       # drop_first (Q1) default
    assert eng.spatial_sort and not eng.exact_div
    eng.add_frames([to_device_f32(pos_a, cuda), to_device_f32(pos_b, cuda)], [0])
    got = eng.counts()
    ref_ab = orc.rdf_counts_slab(pos_a[1:, 0], pos_b[1:, 0], box, cutoff, nbins)
    assert ref_ab.sum() > 2e8
    assert np.array_equal(got[1], ref_ab), \
        f"{np.count_nonzero(got[1] != ref_ab)} of {nbins} A_B bins differ from the oracle"
    ref_aa = orc.rdf_counts_direct(pos_a, [0], [513], box, cutoff, nbins)[(0, 0)]
    assert np.array_equal(got[0], ref_aa)
    rep = eng.tie_report()
    assert rep["pairs_checked"] > 0 and rep["ties"] < rep["pairs_checked"] // 100
