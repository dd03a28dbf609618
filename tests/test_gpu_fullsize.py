"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot
finish these sizes): analytic trajectories for MSD / ACF, round trips for the unwrap, and
invariances of the RDF histogram.  C5 shard: 125,000 atoms x 2,000 frames, data_range 500;
C4: 100,000 atoms per frame."""
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
