"""End-to-end parity of the calculator API (Project -> Experiment -> run.X) against the oracle
on the BASELINE parity configs (scaled where the oracle would take minutes).

C1: 1,000-atom NaCl LAMMPS text dump, 100 frames -> RadialDistributionFunction (+CN)
C2: same system, 1,500 frames, data_range 200    -> Einstein + Green-Kubo diffusion
C3: 1,728-atom NaCl, 1,200 frames, data_range 150 -> unwrap, ionic current, GK ionic conductivity
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-5  # BASELINE.json north_star: series and transport coefficients within 1e-5 relative
MEM = 60e9   # planner memory pinned so that the oracle and the product share one plan


@pytest.fixture(scope="module")
def nacl_c1(tmp_path_factory, cuda):
    from lammps_analysis_b200.config import config
    from lammps_analysis_b200.file_io import write_lammps_dump
    from lammps_analysis_b200.project import Project
    from lammps_analysis_b200.synthetic import nacl_trajectory

    config.planner_memory_bytes = MEM
    tmp = tmp_path_factory.mktemp("c1")
    data, box = nacl_trajectory(1000, 100, 32.0, seed=1)
    dump = str(tmp / "nacl.lammpstraj")
    write_lammps_dump(dump, data, box, step_stride=10)
    project = Project("c1", storage_path=str(tmp))
    exp = project.add_experiment("NaCl", timestep=0.002, temperature=1400.0, units="metal",
                                 simulation_data=dump)
    return project, exp, data, box


def test_ingest_matches_source(nacl_c1):
    project, exp, data, box = nacl_c1
    assert list(exp.species) == ["Na", "Cl"]
    assert exp.number_of_configurations == 100 and exp.sample_rate == 10
    assert exp.box_array == [32.0, 32.0, 32.0]
    for sp in ("Na", "Cl"):
        for prop in ("Positions", "Velocities"):
            assert np.array_equal(exp.store.host(f"{sp}/{prop}"), data[sp][prop])


def test_c1_rdf_bit_exact_and_normalised(nacl_c1):
    from oracle import rdf as orc

    project, exp, data, box = nacl_c1
    res = exp.run.RadialDistributionFunction(number_of_configurations=100, plot=False)
    species = ["Na", "Cl"]
    cutoff = orc.default_cutoff(box)
    nbins = orc.default_number_of_bins(cutoff)
    assert nbins == 1590
    frames = orc.sample_configurations(0, 99, 100)
    # reference plan at 60 GB: 100 single-frame batches x 10 atom minibatches
    bs, nb = orc.rdf_plan({"Na": 500, "Cl": 500}, 100, 100, MEM)
    assert (bs, nb) == (1, 100)
    ref_counts = orc.rdf_counts({s: data[s]["Positions"] for s in species}, species, box, frames,
                                cutoff, nbins, 100, nb)
    ref = orc.rdf_normalise(ref_counts, {"Na": 500, "Cl": 500}, box, cutoff, nbins, 100, 1e-10)
    assert res.keys() == ["Na_Na", "Na_Cl", "Cl_Cl"]
    # integer counts: recover them from the calculator that produced the stored result
    calc = exp.run.RadialDistributionFunction
    calc.__class__.__call__.__wrapped__(calc, number_of_configurations=100, plot=False)
    calc.check_input()
    counts = calc.compute_counts()
    ties = 0
    for p, key in enumerate(res.keys()):
        ties += int(np.count_nonzero(counts[p] != ref_counts[key]))
        np.testing.assert_allclose(res[key]["x"], ref[key]["x"], rtol=1e-12)
        y, yr = np.array(res[key]["y"]), np.array(ref[key]["y"])
        assert np.isnan(y[0]) or np.isinf(y[0])          # Q2: r[0] = 0
        np.testing.assert_allclose(y[1:], yr[1:], rtol=RTOL)
    assert ties == 0, f"{ties} bins differ from the oracle"
    # north_star "tie count reported": the census of pairs on an fp32 bin edge is stored with
    # the result (outside data_dict, whose keys stay the reference's) and survives the cache
    rep = res.metadata["tie_report"]
    assert rep["pairs_checked"] > 100_000 and 0 <= rep["ties"] < rep["pairs_checked"] // 1000
    again = exp.run.RadialDistributionFunction(number_of_configurations=100, plot=False)
    assert again.id == res.id and again.metadata["tie_report"] == rep
    assert set(res.data_dict) == {"Na_Na", "Na_Cl", "Cl_Cl"}


def test_cache_returns_stored_computation(nacl_c1):
    project, exp, data, box = nacl_c1
    a = exp.run.RadialDistributionFunction(number_of_configurations=100, plot=False)
    b = exp.run.RadialDistributionFunction(number_of_configurations=100, plot=False)
    assert a.id == b.id
    assert a.computation_parameter["number_of_configurations"] == 100
    assert a.computation_parameter["cutoff"] is None      # stored with unresolved defaults
    d = project.run.RadialDistributionFunction(number_of_configurations=100, plot=False)
    assert list(d) == ["NaCl"] and d["NaCl"].id == a.id
    c = exp.run.RadialDistributionFunction(number_of_configurations=50, plot=False)
    assert c.id != a.id


def test_c1_coordination_numbers(nacl_c1):
    from oracle import coordination as oc

    project, exp, data, box = nacl_c1
    rdf = exp.run.RadialDistributionFunction(number_of_configurations=100, plot=False)
    cn = exp.run.CoordinationNumbers(rdf_data=rdf, plot=False, savgol_window_length=31)
    vol_nm3 = np.prod(box) * (1e-10) ** 3 / 1e-27
    ref = oc.coordination_numbers(rdf.data_dict, {"Na": 500, "Cl": 500}, vol_nm3,
                                  savgol_window_length=31)
    for key in ref:
        np.testing.assert_allclose(cn[key]["cn"], ref[key]["cn"], rtol=1e-9)
        assert cn[key]["CN_1"] == pytest.approx(ref[key]["CN_1"], rel=1e-9)
        assert cn[key]["CN_1_error"] == pytest.approx(ref[key]["CN_1_error"], rel=1e-6, abs=1e-12)


@pytest.fixture(scope="module")
def nacl_c2(tmp_path_factory, cuda):
    from lammps_analysis_b200.config import config
    from lammps_analysis_b200.file_io import ScriptInput
    from lammps_analysis_b200.project import Project
    from lammps_analysis_b200.synthetic import nacl_trajectory

    config.planner_memory_bytes = MEM
    tmp = tmp_path_factory.mktemp("c2")
    data, box = nacl_trajectory(1000, 1500, 32.0, seed=2, sigma_step=0.3)
    project = Project("c2", storage_path=str(tmp))
    exp = project.add_experiment("NaCl", timestep=0.002, temperature=1400.0, units="metal")
    exp.add_data(ScriptInput(data, box, sample_rate=10, charges={"Na": 1.0, "Cl": -1.0},
                             atom_major=True))
    return project, exp, data, box


def _oracle_plan(A, T, N, ct, scale):
    from oracle.planner import ArrayDatabase, plan_trajectory_calculator

    class _S:
        shape = (A, T, 3)

    return plan_trajectory_calculator(ArrayDatabase({"x": _S()}), ["x"], N, ct,
                                      {"linear": {"scale_factor": scale}}, MEM)


def test_c2_einstein(nacl_c2):
    from oracle import dynamics as od
    from oracle import transformations as ot

    project, exp, data, box = nacl_c2
    N = 200
    res = exp.run.EinsteinDiffusionCoefficients(data_range=N, plot=False)
    plan = _oracle_plan(500, 1500, N, 1, 150)
    for sp in ("Na", "Cl"):
        unw = ot.run_unwrap(data[sp]["Positions"], box, batch_size=1500)
        assert np.array_equal(exp.store.host(f"{sp}/Unwrapped_Positions"), unw)
        tau, _, _, times = od.handle_tau_values(np.s_[:], N, 0.002, 10)
        msd_sum, count = od.einstein_msd(unw, plan, N, 1, tau)
        assert count == 1300 * 501
        ref = od.einstein_finish(msd_sum, count, times, 1e-10, 1e-12, N - 1)
        np.testing.assert_allclose(res[sp]["msd"], ref["msd"], rtol=RTOL)
        np.testing.assert_allclose(res[sp]["time"], ref["time"], rtol=1e-12)
        assert res[sp]["diffusion_coefficient"] == pytest.approx(ref["diffusion_coefficient"],
                                                                 rel=RTOL)
        assert res[sp]["uncertainty"] == pytest.approx(ref["uncertainty"], rel=1e-3)


def test_c2_green_kubo_diffusion(nacl_c2):
    from oracle import dynamics as od

    project, exp, data, box = nacl_c2
    N = 200
    res = exp.run.GreenKuboDiffusionCoefficients(data_range=N, plot=False)
    plan = _oracle_plan(500, 1500, N, 1, 150)
    for sp in ("Na", "Cl"):
        _, _, _, times = od.handle_tau_values(np.s_[:], N, 0.002, 10)
        time = times * 1e-12
        acf_sum, count, sigmas = od.gk_diffusion_acf(data[sp]["Velocities"], plan, N, 1, time,
                                                     1e-10, 1e-12)
        ref = od.gk_diffusion_finish(acf_sum, count, sigmas, time, N - 1)
        scale = np.abs(ref["acf"]).max()
        np.testing.assert_allclose(res[sp]["acf"], ref["acf"], rtol=RTOL, atol=1e-7 * scale)
        np.testing.assert_allclose(res[sp]["integral"], ref["integral"], rtol=RTOL)
        assert res[sp]["diffusion_coefficient"][0] == pytest.approx(
            ref["diffusion_coefficient"][0], rel=RTOL)
        assert res[sp]["uncertainty"][0] == pytest.approx(ref["uncertainty"][0], rel=1e-4)


def test_c3_ionic_conductivity(tmp_path, cuda):
    from lammps_analysis_b200.config import config
    from lammps_analysis_b200.file_io import ScriptInput
    from lammps_analysis_b200.project import Project
    from lammps_analysis_b200.synthetic import nacl_trajectory
    from oracle import dynamics as od
    from oracle import transformations as ot
    from oracle.planner import ArrayDatabase, plan_trajectory_calculator

    config.planner_memory_bytes = MEM
    data, box = nacl_trajectory(1728, 1200, 38.0, seed=3)
    project = Project("c3", storage_path=str(tmp_path))
    exp = project.add_experiment("NaCl", timestep=0.002, temperature=1400.0, units="metal")
    exp.add_data(ScriptInput(data, box, sample_rate=10, atom_major=True))
    exp.species["Na"].charge = 1.0
    exp.species["Cl"].charge = -1.0
    N = 150
    res = exp.run.GreenKuboIonicConductivity(data_range=N, plot=False)
    J_ref = ot.run_ionic_current({s: data[s]["Velocities"] for s in ("Na", "Cl")},
                                 {"Na": 1.0, "Cl": -1.0})
    J = exp.store.host("Observables/Ionic_Current")
    assert J.shape == (1, 1200, 3)
    # fp64 atom sum rounded once to float32: allow 1 ulp for the summation order
    np.testing.assert_allclose(J, J_ref, rtol=2e-7, atol=1e-6)

    class _S:
        shape = (1, 1200, 3)

    plan = plan_trajectory_calculator(ArrayDatabase({"x": _S()}), ["x"], N, 1,
                                      {"linear": {"scale_factor": 5}}, MEM)
    tau, _, _, times = od.handle_tau_values(np.s_[:], N, 0.002, 10)
    acf_sum, count, sigmas = od.gk_ionic_acf(J.astype(np.float64), plan, N, 1, tau, times)
    pref = od.gk_ionic_prefactor(1e-10, 1e-12, 1400.0, float(np.prod(box)))
    ref = od.gk_ionic_finish(acf_sum, count, sigmas, times, pref, N - 1)
    assert count == 1050
    scale = np.abs(ref["acf"]).max()
    np.testing.assert_allclose(res["System"]["acf"], ref["acf"], rtol=RTOL, atol=1e-7 * scale)
    assert res["System"]["ionic_conductivity"][0] == pytest.approx(
        ref["ionic_conductivity"][0], rel=RTOL)
    assert res["System"]["uncertainty"][0] == pytest.approx(ref["uncertainty"][0], rel=1e-4)


def test_atom_minibatch_plan_matches_oracle(tmp_path, cuda):
    """Forces the planner's atom mini-batch path (memory too small for one data_range window),
    as the reference's tests do with change_memory_fraction."""
    from lammps_analysis_b200.config import config
    from lammps_analysis_b200.file_io import ScriptInput
    from lammps_analysis_b200.project import Project
    from oracle import dynamics as od
    from oracle.planner import ArrayDatabase, plan_trajectory_calculator

    rng = np.random.default_rng(9)
    A, T, N = 200, 400, 120
    x = np.cumsum(rng.normal(0, 0.1, size=(A, T, 3)), axis=1).astype(np.float32)
    mem = 150 * (A * 12) * 100 / 0.5     # room for 100 frames of all atoms -> minibatch
    config.planner_memory_bytes = mem
    try:
        project = Project("mb", storage_path=str(tmp_path))
        exp = project.add_experiment("X", timestep=0.001, temperature=300.0, units="real")
        exp.add_data(ScriptInput({"Ar": {"Unwrapped_Positions": x}}, [50.0] * 3, atom_major=True))
        res = exp.run.EinsteinDiffusionCoefficients(data_range=N, plot=False)

        class _S:
            shape = (A, T, 3)

        plan = plan_trajectory_calculator(ArrayDatabase({"x": _S()}), ["x"], N, 1,
                                          {"linear": {"scale_factor": 150}}, mem)
        assert plan["minibatch"] and plan["n_atom_batches"] == 2
        msd_sum, count = od.einstein_msd(x, plan, N, 1, np.arange(N))
        ref = np.array(msd_sum) / count * (1e-10) ** 2
        np.testing.assert_allclose(res["Ar"]["msd"], ref, rtol=RTOL)
    finally:
        config.planner_memory_bytes = MEM


def test_two_rank_sharding_matches_single_rank(cuda):
    """Runs tests/multigpu_check.py under torchrun with two ranks: one per GPU over NCCL when
    the box has two GPUs, otherwise both on the one GPU with the collectives going through the
    host (gloo) -- the calculators' sharding, frame exchange and reductions are the same code."""
    import os
    import subprocess
    import sys

    import torch

    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ)
    env["MDK_MG_BACKEND"] = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29611",
           os.path.join(here, "multigpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "multigpu_check ok on 2 ranks" in res.stdout


# ---------------------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------------------
def _small_project(tmp_path, name, data, box, **kw):
    from lammps_analysis_b200.config import config
    from lammps_analysis_b200.file_io import ScriptInput
    from lammps_analysis_b200.project import Project

    config.planner_memory_bytes = MEM
    project = Project(name, storage_path=str(tmp_path))
    exp = project.add_experiment("X", timestep=0.001, temperature=300.0, units="real")
    exp.add_data(ScriptInput(data, box, atom_major=True, **kw))
    return project, exp


def test_rdf_three_species_noncubic_box_and_tiny_species(tmp_path, cuda):
    """3 species (6 pairs), orthorhombic box, one species with a single atom (after the
    reference's first-atom drop it contributes nothing), explicit cutoff / bins / frame range."""
    from oracle import rdf as orc

    rng = np.random.default_rng(51)
    box = np.array([21.0, 17.5, 25.0])
    counts = {"A": 150, "B": 1, "C": 77}
    data = {s: {"Positions": (rng.random((n, 9, 3)) * box).astype(np.float32)}
            for s, n in counts.items()}
    project, exp = _small_project(tmp_path, "rdf3", data, box)
    res = exp.run.RadialDistributionFunction(number_of_configurations=4, start=1, stop=7,
                                             cutoff=8.0, number_of_bins=123, plot=False)
    frames = orc.sample_configurations(1, 7, 4)
    ref_counts = orc.rdf_counts({s: data[s]["Positions"] for s in counts}, list(counts), box,
                                frames, 8.0, 123, 4, 4)
    ref = orc.rdf_normalise(ref_counts, counts, box, 8.0, 123, 4, 1e-10)
    assert res.keys() == ["A_A", "A_B", "A_C", "B_B", "B_C", "C_C"]
    for key in res.keys():
        y, yr = np.array(res[key]["y"])[1:], np.array(ref[key]["y"])[1:]
        np.testing.assert_allclose(y, yr, rtol=RTOL)
    assert np.all(np.array(res["A_B"]["y"])[1:] == 0) and np.all(np.array(res["B_B"]["y"])[1:] == 0)


def test_rdf_atom_selection_and_species_subset(tmp_path, cuda):
    from oracle import rdf as orc

    rng = np.random.default_rng(52)
    box = np.array([15.0, 15.0, 15.0])
    data = {s: {"Positions": (rng.random((90, 5, 3)) * 15).astype(np.float32)} for s in "AB"}
    project, exp = _small_project(tmp_path, "rdfsel", data, box)
    sel = {"A": [3, 5, 8, 13, 21, 34, 55, 89], "B": list(range(10, 40))}
    res = exp.run.RadialDistributionFunction(number_of_configurations=5, atom_selection=sel,
                                             plot=False)
    picked = {s: data[s]["Positions"][sel[s]] for s in "AB"}
    cutoff = orc.default_cutoff(box)
    nbins = orc.default_number_of_bins(cutoff)
    ref_counts = orc.rdf_counts(picked, ["A", "B"], box, orc.sample_configurations(0, 4, 5),
                                cutoff, nbins, 5, 5)
    ref = orc.rdf_normalise(ref_counts, {"A": 8, "B": 30}, box, cutoff, nbins, 5, 1e-10)
    for key in res.keys():
        np.testing.assert_allclose(np.array(res[key]["y"])[1:], np.array(ref[key]["y"])[1:],
                                   rtol=RTOL)
    only_b = exp.run.RadialDistributionFunction(number_of_configurations=5, species=["B"],
                                                plot=False)
    assert only_b.keys() == ["B_B"]


def test_rdf_duplicate_sample_frames_rejected(tmp_path, cuda):
    rng = np.random.default_rng(53)
    data = {"A": {"Positions": (rng.random((20, 10, 3)) * 9).astype(np.float32)}}
    project, exp = _small_project(tmp_path, "rdfdup", data, [9.0] * 3)
    with pytest.raises(ValueError):
        # default number_of_configurations=500 > 10 frames (the reference fails in h5py here)
        exp.run.RadialDistributionFunction(plot=False)


def test_einstein_tau_resolution_and_correlation_time(tmp_path, cuda):
    """tau_values as an int (sub-sampled lags) and correlation_time > 1 go through the generic
    kernel; results follow the oracle."""
    from oracle import dynamics as od

    rng = np.random.default_rng(54)
    A, T, N = 40, 500, 90
    x = np.cumsum(rng.normal(0, 0.2, size=(A, T, 3)), axis=1).astype(np.float32)
    project, exp = _small_project(tmp_path, "eintau", {"Ar": {"Unwrapped_Positions": x}},
                                  [40.0] * 3)
    # with 16 lags the default fit_range (data_range - 1) is never reached -- the reference
    # crashes on that too -- so the fit end index is given explicitly
    res = exp.run.EinsteinDiffusionCoefficients(data_range=N, correlation_time=3, tau_values=16,
                                                fit_range=15, plot=False)
    plan = _oracle_plan(A, T, N, 3, 150)
    tau, dr, _, times = od.handle_tau_values(16, N, 0.001, 1)
    msd_sum, count = od.einstein_msd(x, plan, dr, 3, tau)
    np.testing.assert_allclose(res["Ar"]["msd"], np.array(msd_sum) / count * 1e-20, rtol=RTOL)
    np.testing.assert_allclose(res["Ar"]["time"], times * 1e-15, rtol=1e-12)


def test_missing_data_errors(tmp_path, cuda):
    from lammps_analysis_b200.transformations import CannotFindPropertyError

    rng = np.random.default_rng(55)
    data = {"A": {"Positions": (rng.random((20, 30, 3)) * 9).astype(np.float32)}}
    project, exp = _small_project(tmp_path, "missing", data, [9.0] * 3)
    with pytest.raises(CannotFindPropertyError):
        exp.run.GreenKuboIonicConductivity(data_range=10, plot=False)   # no velocities
    with pytest.raises(KeyError):
        exp.run.GreenKuboDiffusionCoefficients(data_range=10, plot=False)
    # unwrapping works and is skipped the second time
    exp.run.CoordinateUnwrapper()
    first = exp.store.host("A/Unwrapped_Positions").copy()
    exp.run.CoordinateUnwrapper()
    assert np.array_equal(first, exp.store.host("A/Unwrapped_Positions"))


def test_unwrap_via_indices_is_chosen_when_box_images_exist(tmp_path, cuda):
    rng = np.random.default_rng(56)
    pos = (rng.random((15, 40, 3)) * 7).astype(np.float32)
    img = rng.integers(-2, 3, size=(15, 40, 3)).astype(np.float32)
    project, exp = _small_project(tmp_path, "uvi", {"A": {"Positions": pos, "Box_Images": img}},
                                  [7.0, 7.5, 8.0])
    calc = exp.run.EinsteinDiffusionCoefficients
    type(calc).__call__.__wrapped__(calc, data_range=10, plot=False)
    calc.check_input()          # dependency resolution: trajectory_calculator.py:117-194
    want = (pos.astype(np.float64) + img.astype(np.float64) * np.array([7.0, 7.5, 8.0]))
    assert np.array_equal(exp.store.host("A/Unwrapped_Positions"), want.astype(np.float32))


# ---------------------------------------------------------------------------------------------
# "next" rows: consumers of the MSD kernel and of the RDF result
# ---------------------------------------------------------------------------------------------
def test_c1_pmf_and_kirkwood_buff(nacl_c1):
    from oracle import coordination as oc

    project, exp, data, box = nacl_c1
    rdf = exp.run.RadialDistributionFunction(number_of_configurations=100, plot=False)
    kb = exp.run.KirkwoodBuffIntegral(rdf_data=rdf, plot=False, savgol_window_length=31)
    ref = oc.kirkwood_buff_integral(rdf.data_dict, savgol_window_length=31)
    for key in ref:
        np.testing.assert_allclose(kb[key]["kb_integral"], ref[key]["kb_integral"], rtol=1e-9)
        np.testing.assert_allclose(kb[key]["r"], ref[key]["r"], rtol=1e-12)
    # the measured g(r) is exactly zero inside the core, where ln g diverges (scipy's savgol
    # filter rejects the infinities, as it does for the reference): the PMF is checked on a
    # strictly positive model g(r) handed over as a Computation
    from collections import OrderedDict

    from lammps_analysis_b200.project import Computation

    r = np.linspace(0.0, 1.5, 600)
    g = 1.0 + 1.8 * np.exp(-3.0 * r) * np.cos(14.0 * r - 1.0)
    model = Computation("Radial_Distribution_Function", "NaCl",
                        {"number_of_bins": 600, "cutoff": 15.0, "number_of_configurations": 7},
                        OrderedDict([("Na_Cl", {"x": r.tolist(), "y": g.tolist()})]))
    pmf = exp.run.PotentialOfMeanForce(rdf_data=model, plot=False)
    ref = oc.potential_of_mean_force(model.data_dict, 1400.0)
    np.testing.assert_allclose(pmf["Na_Cl"]["pomf"], ref["Na_Cl"]["pomf"], rtol=1e-12)
    assert pmf["Na_Cl"]["POMF_1"] == pytest.approx(ref["Na_Cl"]["POMF_1"], rel=1e-12)
    assert pmf["Na_Cl"]["POMF_1_error"] == pytest.approx(ref["Na_Cl"]["POMF_1_error"], rel=1e-9,
                                                          abs=1e-30)


def test_einstein_helfand_ionic_conductivity(tmp_path, cuda):
    from lammps_analysis_b200.config import config
    from lammps_analysis_b200.file_io import ScriptInput
    from lammps_analysis_b200.project import Project
    from lammps_analysis_b200.synthetic import nacl_trajectory
    from oracle import dynamics as od
    from oracle import transformations as ot
    from oracle.planner import ArrayDatabase, plan_trajectory_calculator

    config.planner_memory_bytes = MEM
    data, box = nacl_trajectory(512, 900, 26.0, seed=6, sigma_step=0.3)
    project = Project("eh", storage_path=str(tmp_path))
    exp = project.add_experiment("NaCl", timestep=0.002, temperature=1400.0, units="metal")
    exp.add_data(ScriptInput({s: {"Positions": data[s]["Positions"]} for s in data}, box,
                             sample_rate=10, atom_major=True))
    exp.species["Na"].charge = 1.0
    exp.species["Cl"].charge = -1.0
    N = 120
    res = exp.run.EinsteinHelfandIonicConductivity(data_range=N, plot=False)
    # the transformation chain ran: unwrap -> dipole moment (stored as float32)
    unw = {s: ot.run_unwrap(data[s]["Positions"], box, batch_size=900) for s in data}
    M_ref = ot.dipole_moment_transform_batch({
        "Na": {"Unwrapped_Positions": unw["Na"], "Charge": np.array([[[1.0]]])},
        "Cl": {"Unwrapped_Positions": unw["Cl"], "Charge": np.array([[[-1.0]]])}})
    M = exp.store.host("Observables/Translational_Dipole_Moment")
    assert M.shape == (1, 900, 3)
    np.testing.assert_allclose(M[0], M_ref, rtol=2e-7, atol=1e-4)

    class _S:
        shape = (1, 900, 3)

    plan = plan_trajectory_calculator(ArrayDatabase({"x": _S()}), ["x"], N, 1,
                                      {"linear": {"scale_factor": 5}}, MEM)
    tau, _, _, times = od.handle_tau_values(np.s_[:], N, 0.002, 10)
    pref = od.eh_ionic_prefactor(1e-10, 1e-12, 1400.0, float(np.prod(box)))
    msd = od.eh_ionic_msd(M.astype(np.float64), plan, N, 1, tau, pref)
    np.testing.assert_allclose(res["System"]["msd"], msd, rtol=RTOL, atol=1e-9 * np.abs(msd).max())
    popt, pcov, _, _ = od.fit_einstein_curve(times, msd, N - 1)
    assert res["System"]["ionic_conductivity"] == pytest.approx(popt[0] / 6, rel=1e-4)


def _flux_experiment(tmp_path, n_atoms=216, n_frames=700, seed=21):
    """Two species with per-atom Stress (6), Velocities, Kinetic / Potential energy and wrapped
    positions: the inputs of the MomentumFlux / ThermalFlux / IntegratedHeatCurrent chain."""
    from lammps_analysis_b200.config import config
    from lammps_analysis_b200.file_io import ScriptInput
    from lammps_analysis_b200.project import Project
    from lammps_analysis_b200.synthetic import nacl_trajectory

    config.planner_memory_bytes = MEM
    data, box = nacl_trajectory(n_atoms, n_frames, 18.0, seed=seed, sigma_step=0.2)
    rng = np.random.default_rng(seed)
    for sp in data:
        A = data[sp]["Positions"].shape[0]
        # slowly varying per-atom fields so that the flux autocorrelations are not pure noise
        walk = np.cumsum(rng.normal(0, 0.05, size=(A, n_frames, 6)), axis=1)
        data[sp]["Stress"] = (rng.normal(0, 1.0, size=(A, 1, 6)) + walk).astype(np.float32)
        data[sp]["Kinetic_Energy"] = rng.uniform(0.5, 1.5, size=(A, n_frames, 1)).astype(np.float32)
        data[sp]["Potential_Energy"] = rng.normal(-3.0, 0.3, size=(A, n_frames, 1)).astype(np.float32)
    project = Project("flux", storage_path=str(tmp_path))
    exp = project.add_experiment("NaCl", timestep=0.002, temperature=1400.0, units="metal")
    exp.add_data(ScriptInput(data, box, sample_rate=10, atom_major=True))
    return exp, data, box


def _system_plan(n_frames, N):
    from oracle.planner import ArrayDatabase, plan_trajectory_calculator

    class _S:
        shape = (1, n_frames, 3)

    return plan_trajectory_calculator(ArrayDatabase({"x": _S()}), ["x"], N, 1,
                                      {"linear": {"scale_factor": 5}}, MEM)


def test_flux_transformations_match_oracle(tmp_path, cuda):
    """MomentumFlux, ThermalFlux, IntegratedHeatCurrent against the restated
    transform_batch bodies; the observables are stored as float32 (1, T, 3)."""
    from oracle import transformations as ot

    exp, data, box = _flux_experiment(tmp_path)
    exp.run.MomentumFlux()
    exp.run.ThermalFlux()
    exp.run.IntegratedHeatCurrent()      # runs the unwrap first
    ref_m = ot.momentum_flux_transform_batch({s: {"Stress": data[s]["Stress"]} for s in data})
    ref_t = ot.thermal_flux_transform_batch({s: {k: data[s][k] for k in
                                                 ("Stress", "Velocities", "Kinetic_Energy",
                                                  "Potential_Energy")} for s in data})
    unw = {s: ot.run_unwrap(data[s]["Positions"], box, batch_size=data[s]["Positions"].shape[1])
           for s in data}
    ref_h = ot.integrated_heat_current_transform_batch(
        {s: {"Unwrapped_Positions": unw[s], "Kinetic_Energy": data[s]["Kinetic_Energy"],
             "Potential_Energy": data[s]["Potential_Energy"]} for s in data})
    for name, ref in (("Momentum_Flux", ref_m), ("Thermal_Flux", ref_t),
                      ("Integrated_Heat_Current", ref_h)):
        got = exp.store.host(f"Observables/{name}")
        assert got.shape == (1, ref.shape[0], 3) and got.dtype == np.float32
        np.testing.assert_allclose(got[0], ref, rtol=2e-7, atol=1e-6 * np.abs(ref).max())
    # a second call finds the datasets and skips (transformations.py:572-579)
    exp.run.MomentumFlux()


def test_thermal_flux_from_an_ingested_lammps_dump(tmp_path, cuda):
    """A dump with c_KE, c_PE and c_Stress[1..6] columns is ingested under the reference's
    property names (lammps_trajectory_files.py:56-57 -> Kinetic_Energy / Potential_Energy), so
    the thermal-flux transformation and the calculators behind it find their inputs."""
    from lammps_analysis_b200.file_io import write_lammps_dump
    from lammps_analysis_b200.project import Project
    from oracle import transformations as ot

    rng = np.random.default_rng(77)
    box = [9.0, 9.0, 9.0]
    data = {}
    for sp, n in (("Na", 7), ("Cl", 5)):
        data[sp] = {
            "Positions": (rng.random((n, 12, 3)) * 9).astype(np.float32),
            "Velocities": rng.normal(size=(n, 12, 3)).astype(np.float32),
            "Kinetic_Energy": rng.uniform(0.5, 1.5, size=(n, 12, 1)).astype(np.float32),
            "Potential_Energy": rng.normal(-3, 0.3, size=(n, 12, 1)).astype(np.float32),
            "Stress": rng.normal(size=(n, 12, 6)).astype(np.float32),
        }
    dump = str(tmp_path / "flux.lammpstraj")
    write_lammps_dump(dump, data, box)
    assert "c_KE c_PE c_Stress[1]" in open(dump).read(400)
    project = Project("dumpflux", storage_path=str(tmp_path))
    exp = project.add_experiment("NaCl", timestep=0.002, temperature=300.0, units="metal",
                                 simulation_data=dump)
    for prop in ("Kinetic_Energy", "Potential_Energy", "Stress"):
        assert np.array_equal(exp.store.host(f"Na/{prop}"), data["Na"][prop])
    exp.run.ThermalFlux()
    ref = ot.thermal_flux_transform_batch({s: {k: data[s][k] for k in
                                               ("Stress", "Velocities", "Kinetic_Energy",
                                                "Potential_Energy")} for s in data})
    got = exp.store.host("Observables/Thermal_Flux")[0]
    np.testing.assert_allclose(got, ref, rtol=2e-7, atol=1e-6 * np.abs(ref).max())


@pytest.mark.parametrize("which", ["thermal", "viscosity"])
def test_green_kubo_flux_calculators(tmp_path, cuda, which):
    """GreenKuboThermalConductivity / GreenKuboViscosity: dependency resolution runs the flux
    transformation, the windowed ACF kernels reproduce the restated reference (value = first
    window's integral, "uncertainty" = second window's, acf = sum over windows)."""
    from lammps_analysis_b200.units import METAL
    from oracle import dynamics as od

    exp, data, box = _flux_experiment(tmp_path, seed=22)
    N, ir = 100, 80
    if which == "thermal":
        res = exp.run.GreenKuboThermalConductivity(data_range=N, integration_range=ir)
        prop, key = "Thermal_Flux", "computation_results"
        pref = od.gk_thermal_prefactor(METAL, 1400.0, float(np.prod(box)), N)
    else:
        res = exp.run.GreenKuboViscosity(data_range=N, integration_range=ir)
        prop, key = "Momentum_Flux", "viscosity"
        pref = od.gk_viscosity_prefactor(METAL, 1400.0, float(np.prod(box)), N)
    J = exp.store.host(f"Observables/{prop}").astype(np.float64)
    T = J.shape[1]
    _, _, _, times = od.handle_tau_values(np.s_[:], N, 0.002, 10)
    ref = od.gk_flux(J, _system_plan(T, N), N, 1, times, ir, pref, key)
    got = res["System"]
    assert set(got) == {key, "uncertainty", "time", "acf"}
    scale = np.abs(ref["acf"]).max()
    np.testing.assert_allclose(got["acf"], ref["acf"], rtol=RTOL, atol=1e-7 * scale)
    np.testing.assert_allclose(got["time"], ref["time"], rtol=1e-12)
    assert got[key] == pytest.approx(ref[key], rel=1e-4)
    assert got["uncertainty"] == pytest.approx(ref["uncertainty"], rel=1e-4)


def test_green_kubo_viscosity_flux_from_a_flux_file(tmp_path, cuda):
    """GreenKuboViscosityFlux: Stress_Visc read from a LAMMPS flux table by
    LAMMPSFluxFile, windowed ACF on the device, the reference's reporting (volume in the
    numerator of the prefactor, first two windows' integrals, normalised constant acf)."""
    from lammps_analysis_b200.config import config
    from lammps_analysis_b200.file_io import LAMMPSFluxFile
    from lammps_analysis_b200.project import Project
    from lammps_analysis_b200.units import REAL
    from oracle import dynamics as od

    config.planner_memory_bytes = MEM
    rng = np.random.default_rng(61)
    T, N, ir = 900, 120, 100
    p = np.empty((T, 3))
    p[0] = rng.normal(size=3)
    for t in range(1, T):
        p[t] = 0.95 * p[t - 1] + rng.normal(0, 0.3, size=3)
    path = str(tmp_path / "visc.lmp")
    with open(path, "w") as fh:
        fh.write("# Fix print output\ntime temp pxy pxz pyz\n")
        for t in range(T):
            fh.write("%d 300.0 %r %r %r\n" % (t, float(p[t, 0]), float(p[t, 1]), float(p[t, 2])))
    box = [20.0, 20.0, 20.0]
    project = Project("visc", storage_path=str(tmp_path))
    exp = project.add_experiment("LJ", timestep=0.004, temperature=85.0, units="real",
                                 simulation_data=LAMMPSFluxFile(path, sample_rate=2, box_l=box))
    res = exp.run.GreenKuboViscosityFlux(data_range=N, integration_range=ir)
    J = exp.store.host("Observables/Stress_Visc").astype(np.float64)
    _, _, _, times = od.handle_tau_values(np.s_[:], N, 0.004, 2)
    pref = od.gk_viscosity_flux_prefactor(REAL, 85.0, float(np.prod(box)), N)
    ref = od.gk_viscosity_flux(J, _system_plan(T, N), N, 1, times, ir, pref)
    got = res["System"]
    assert set(got) == {"viscosity", "uncertainty", "time", "acf"}
    assert np.isfinite(ref["viscosity"]) and np.isfinite(ref["uncertainty"])
    assert got["viscosity"] == pytest.approx(ref["viscosity"], rel=1e-4)
    assert got["uncertainty"] == pytest.approx(ref["uncertainty"], rel=1e-4)
    np.testing.assert_allclose(got["acf"], ref["acf"], rtol=1e-6)
    assert np.allclose(got["acf"], 1.0)
    # no transformation produces Stress_Visc (trajectory_calculator.py:171-174)
    exp2 = project.add_experiment("empty", timestep=0.004, temperature=85.0, units="real")
    exp2.box_array, exp2.number_of_configurations = box, T
    with pytest.raises(KeyError):
        exp2.run.GreenKuboViscosityFlux(data_range=N)


def test_einstein_helfand_thermal_conductivity(tmp_path, cuda):
    from lammps_analysis_b200.units import METAL
    from oracle import dynamics as od

    exp, data, box = _flux_experiment(tmp_path, seed=23)
    N = 90
    res = exp.run.EinsteinHelfandThermalConductivity(data_range=N, plot=False)
    Q = exp.store.host("Observables/Integrated_Heat_Current").astype(np.float64)
    tau, _, _, times = od.handle_tau_values(np.s_[:], N, 0.002, 10)
    pref = od.eh_thermal_prefactor(METAL, 1400.0, float(np.prod(box)))
    msd = od.eh_ionic_msd(Q, _system_plan(Q.shape[1], N), N, 1, tau, pref)
    np.testing.assert_allclose(res["System"]["msd"], msd, rtol=RTOL, atol=1e-9 * np.abs(msd).max())
    popt, _, _, _ = od.fit_einstein_curve(times, msd, N - 1)
    assert res["System"]["thermal_conductivity"] == pytest.approx(popt[0] / 6, rel=1e-4)


def test_block_input_and_pinned_pool(tmp_path, cuda):
    """BlockInput writes atom blocks straight into the page-locked datasets (exact-size
    cudaHostRegister blocks the kernels can read in place); a removed dataset hands its block to
    the pool and the next dataset of that size reuses it."""
    from lammps_analysis_b200.file_io import BlockInput
    from lammps_analysis_b200.project import Project
    from lammps_analysis_b200.store import pinned_pool

    rng = np.random.default_rng(8)
    full = {sp: rng.random((n, 20, 3)).astype(np.float32) for sp, n in (("A", 50), ("B", 31))}

    def blocks():
        for sp, arr in full.items():
            for lo in range(0, arr.shape[0], 17):
                hi = min(arr.shape[0], lo + 17)
                yield sp, "Positions", (lo, hi), arr[lo:hi]

    project = Project("blk", storage_path=str(tmp_path), persist=False)
    exp = project.add_experiment("e", timestep=1.0, temperature=1.0, units="real")
    exp.add_data(BlockInput(20, {"A": (50, {"Positions": 3}), "B": (31, {"Positions": 3})},
                            [9.0, 9.0, 9.0], blocks))
    assert exp.species["A"].n_particles == 50 and exp.number_of_configurations == 20
    for sp in full:
        assert np.array_equal(exp.store.host(f"{sp}/Positions"), full[sp])
    pin = exp.store.pinned_tensor("A/Positions")
    assert pin is not None and pin.is_pinned()
    assert np.array_equal(exp.store.device("A/Positions").cpu().numpy(), full["A"])
    rdf = exp.run.RadialDistributionFunction(number_of_configurations=5, plot=False)  # zero-copy pack
    assert np.nansum(rdf["A_B"]["y"][1:]) > 0
    pinned_pool.clear()
    ptr = pin.data_ptr()
    del pin
    exp.store.remove("A/Positions")
    assert pinned_pool.cached_bytes == 50 * 20 * 3 * 4
    again = exp.store.add_dataset("A/Positions", (50, 20, 3))
    assert exp.store.pinned_tensor("A/Positions").data_ptr() == ptr and pinned_pool.cached_bytes == 0
    assert again.shape == (50, 20, 3)


def test_in_memory_pinned_store_matches_persistent_store(tmp_path, cuda):
    """persist=False keeps datasets in page-locked host memory: the RDF pack kernels gather the
    sampled frames from it in place (zero copy), uploads are single DMA transfers.  Same numbers
    as the file-backed store."""
    from lammps_analysis_b200.config import config
    from lammps_analysis_b200.file_io import ScriptInput
    from lammps_analysis_b200.project import Project
    from lammps_analysis_b200.synthetic import nacl_trajectory

    config.planner_memory_bytes = MEM
    data, box = nacl_trajectory(512, 300, 26.0, seed=12, sigma_step=0.3)
    results = []
    for persist in (True, False):
        project = Project(f"p{int(persist)}", storage_path=str(tmp_path), persist=persist)
        exp = project.add_experiment("NaCl", timestep=0.002, temperature=1400.0, units="metal")
        exp.add_data(ScriptInput(data, box, sample_rate=10, atom_major=True))
        assert exp.store.pinned == (not persist)
        rdf = exp.run.RadialDistributionFunction(number_of_configurations=23, start=5, plot=False)
        exp.run.CoordinateUnwrapper()
        ein = exp.run.EinsteinDiffusionCoefficients
        type(ein).__call__.__wrapped__(ein, data_range=60, plot=False)
        ein._handle_tau_values()
        msd = {sp: ein.compute_msd(sp)[0] for sp in ("Na", "Cl")}
        gk = exp.run.GreenKuboDiffusionCoefficients(data_range=60, plot=False)
        results.append((rdf.data_dict, msd, gk.data_dict))
    (r0, e0, g0), (r1, e1, g1) = results
    for key in r0:
        assert np.array_equal(np.array(r0[key]["y"])[1:], np.array(r1[key]["y"])[1:])
    for sp in e0:
        np.testing.assert_allclose(e0[sp], e1[sp], rtol=1e-12)   # fp64 atomics: order may differ
        np.testing.assert_allclose(g0[sp]["acf"], g1[sp]["acf"], rtol=1e-12)
