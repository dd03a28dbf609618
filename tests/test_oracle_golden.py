"""Pins the oracle (oracle/) against the known answers the reference's own unit tests hold
for the hot path: unwrap with carry-over, unwrap via indices, ionic current,
memory-manager planning arithmetic, fit_einstein_curve.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import dynamics as od
from oracle import planner as op
from oracle import transformations as ot

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


def test_unwrap_with_carry_known_answer():
    g = _load("unwrap_carry.json")
    out, carry = ot.unwrap_transform_batch(
        np.asarray(g["pos"]), np.asarray(g["box"]),
        {"last_pos": np.asarray(g["last_pos"]), "last_image_box": np.asarray(g["last_image_box"])},
    )
    np.testing.assert_allclose(out, np.asarray(g["expected"]), atol=1e-12)
    np.testing.assert_allclose(carry["last_pos"], g["expected_last_pos"])
    np.testing.assert_allclose(carry["last_image_box"], g["expected_image_box"])


def test_unwrap_via_indices_known_answer():
    # CI/unit_tests/transformations/test_transformations.py:192-213
    rng = np.random.default_rng(0)
    pos = rng.random((5, 7, 3))
    box_im = rng.integers(-10, 10, size=(5, 7, 3)).astype(float)
    box_l = np.array([1.1, 2.2, 3.3])
    np.testing.assert_allclose(
        ot.unwrap_via_indices_transform_batch(pos, box_im, box_l), pos + box_im * box_l
    )


def test_ionic_current_known_answer():
    # CI/unit_tests/transformations/test_transformations.py:59-79
    rng = np.random.default_rng(1)
    batch, should = {}, np.zeros((7, 3))
    for sp in ["Na", "Cl"]:
        vel = rng.random((5, 7, 3))
        charge = np.array([[[rng.random()]]])
        batch[sp] = {"Velocities": vel, "Charge": charge}
        should += np.sum(vel * charge, axis=0)
    np.testing.assert_allclose(ot.ionic_current_transform_batch(batch), should)


class _FakeDB:
    # CI/unit_tests/memory_manager/test_memory_manager.py:28-46
    def __init__(self, data_size=500, rows=10, columns=10):
        self.data_size, self.rows, self.columns = data_size, rows, columns

    def get_data_size(self, item):
        return self.rows, self.columns, self.data_size


def test_planner_known_answers():
    g = _load("planner_cases.json")
    for case in g["get_batch_size"]:
        mm = op.MemoryManager(
            data_path=["Test/Path"],
            database=_FakeDB(case["data_size"], case["rows"], case["columns"]),
            memory_fraction=case["fraction"], memory=case["memory"],
        )
        assert list(mm.get_batch_size()) == case["expect"]
    mm = op.MemoryManager()
    with pytest.raises(ValueError):
        mm.get_batch_size()
    c = g["atomwise_minibatch"]
    mm = op.MemoryManager(data_path=["Test/Path"],
                          database=_FakeDB(c["data_size"], c["rows"], c["columns"]),
                          memory_fraction=c["fraction"], memory=c["memory"])
    mm._compute_atomwise_minibatch(c["data_range"])
    assert mm.batch_size == c["expect"]["batch_size"]
    assert mm.n_batches == c["expect"]["n_batches"]
    assert mm.n_atom_batches == c["expect"]["n_atom_batches"]
    assert mm.atom_remainder == c["expect"]["atom_remainder"]
    for case in g["get_ensemble_loop"]:
        mm = op.MemoryManager(data_path=["Test/Path"], database=_FakeDB(), memory=60e9)
        mm.batch_size = case["batch_size"]
        loops, minibatch = mm.get_ensemble_loop(case["data_range"], case["correlation_time"])
        assert [loops, minibatch] == case["expect"]
    for case in g["scale_functions"]:
        fn, par = op.MemoryManager._select_scale_function(case["fn"])
        assert fn(case["x"], **par) == case["expect"]
    fn, par = op.MemoryManager._select_scale_function({"linear": {"scale_factor": 2}})
    assert fn(10, **par) == 20
    fn, par = op.MemoryManager._select_scale_function({"log-linear": {"scale_factor": 2}})
    assert fn(10, **par) == 20 * np.log(10)


def test_fit_einstein_curve_known_answers():
    # CI/unit_tests/utils/test_calculator_helper_methods.py:42-69 (coarser grid: 200 points
    # keeps the ~200 curve_fit calls inside the CPU-suite time budget)
    x = np.linspace(0, 1000, 200)
    popt, _, _, _ = od.fit_einstein_curve(x, 5 * x + 3, fit_max_index=199)
    assert popt[0] == pytest.approx(5.0, 0.01)
    y = np.exp(-0.05 * x) * x**2 + 5 * x + 3
    popt, _, _, _ = od.fit_einstein_curve(x, y, fit_max_index=199)
    assert popt[0] == pytest.approx(5.0, 0.01)


def test_tfp_autocorrelation_is_unbiased_direct_sum():
    rng = np.random.default_rng(2)
    x = rng.normal(size=(3, 37, 3))
    acf = od.tfp_auto_correlation(x)
    N = x.shape[1]
    direct = np.stack(
        [(x[:, : N - m] * x[:, m:]).sum(axis=1) / (N - m) for m in range(N)], axis=1
    )
    np.testing.assert_allclose(acf, direct, rtol=1e-10, atol=1e-12)


def test_random_walk_diffusion_coefficient():
    # CI/integration_tests/calculators/test_einstein_diffusion_coefficients.py:52-99 scaled
    # down (100 atoms x 2000 steps): D from the MSD slope within the reference's rtol=0.2
    rng = np.random.default_rng(3)
    D, dt = 1.2345, 0.1
    A, T, N = 100, 2000, 200
    x = np.cumsum(rng.normal(0, np.sqrt(2 * D * dt), size=(A, T, 3)), axis=1).astype(np.float32)
    plan = dict(batch_size=T, n_batches=1, remainder=0, minibatch=False)
    tau = np.arange(N)
    msd_sum, count = od.einstein_msd(x, plan, N, 1, tau)
    assert count == (T - N) * (A + 1)
    res = od.einstein_finish(msd_sum, count, tau * dt, 1.0, 1.0, N - 1)
    assert res["diffusion_coefficient"] == pytest.approx(D, rel=0.2)
    np.testing.assert_allclose(res["msd"], 6 * D * tau * dt, rtol=0.9, atol=1e-9)


def test_rdf_counts_independent_of_batch_plan():
    from oracle import rdf as orc

    rng = np.random.default_rng(4)
    box = np.array([12.0, 12.0, 12.0])
    pos = {"Na": (rng.random((30, 3, 3)) * 12).astype(np.float32),
           "Cl": (rng.random((26, 3, 3)) * 12).astype(np.float32)}
    cutoff = orc.default_cutoff(box)
    nbins = orc.default_number_of_bins(cutoff)
    a = orc.rdf_counts(pos, ["Na", "Cl"], box, np.arange(3), cutoff, nbins, 7, 3)
    b = orc.rdf_counts(pos, ["Na", "Cl"], box, np.arange(3), cutoff, nbins, 56, 1)
    allpos = np.concatenate([pos["Na"], pos["Cl"]], axis=0)
    c = orc.rdf_counts_direct(allpos, [0, 30], [30, 26], box, cutoff, nbins)
    for p, key in enumerate(["Na_Na", "Na_Cl", "Cl_Cl"]):
        assert np.array_equal(a[key], b[key])
    assert np.array_equal(a["Na_Na"], c[(0, 0)])
    assert np.array_equal(a["Na_Cl"], c[(0, 1)])
    assert np.array_equal(a["Cl_Cl"], c[(1, 1)])
    # Q1: the first atom of every species is dropped
    assert a["Na_Na"].sum() <= 3 * 29 * 28 // 2
    assert a["Na_Cl"].sum() <= 3 * 29 * 25


# ---------------------------------------------------------------------------------------------
# vectors produced by executing the reference's own Python source (tests/golden/
# make_reference_goldens.py: ast-extracted upstream functions run under a NumPy TF shim)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ref_run():
    return _load("reference_run.json")


def test_reference_run_rdf_counts(ref_run):
    from oracle import rdf as orc

    g = ref_run["rdf"]
    pos = {s: np.asarray(g["positions"][s], dtype=np.float32) for s in g["species"]}
    n_frames = pos[g["species"][0]].shape[1]
    got = orc.rdf_counts(pos, g["species"], np.asarray(g["box"]), np.arange(n_frames),
                         g["cutoff"], g["nbins"], 7, n_frames)
    for key, want in g["counts"].items():
        assert np.array_equal(got[key], np.asarray(want)), key
    assert sum(int(np.sum(v)) for v in g["counts"].values()) > 500


def test_reference_run_transformations(ref_run):
    g = ref_run["unwrap"]
    pos = np.asarray(g["positions"], dtype=np.float32).astype(np.float64)
    carry, pieces = None, []
    for lo, hi in g["batches"]:
        res, carry = ot.unwrap_transform_batch(pos[:, lo:hi], np.asarray(g["box"]), carry)
        pieces.append(res)
    assert np.array_equal(np.concatenate(pieces, axis=1), np.asarray(g["unwrapped"]))
    gi = ref_run["unwrap_indices"]
    assert np.array_equal(
        ot.unwrap_via_indices_transform_batch(pos, np.asarray(gi["images"]), np.asarray(g["box"])),
        np.asarray(gi["unwrapped"]))
    gc = ref_run["ionic_current"]
    batch = {s: {"Velocities": np.asarray(gc["velocities"][s], dtype=np.float32),
                 "Charge": np.array([[[gc["charge"][s]]]])} for s in gc["velocities"]}
    assert np.array_equal(ot.ionic_current_transform_batch(batch), np.asarray(gc["current"]))
    batch = {s: {"Unwrapped_Positions": np.asarray(gc["velocities"][s], dtype=np.float32),
                 "Charge": np.array([[[gc["charge"][s]]]])} for s in gc["velocities"]}
    assert np.array_equal(ot.dipole_moment_transform_batch(batch),
                          np.asarray(ref_run["dipole_moment"]["moment"]))


def test_reference_run_flux_transformations():
    """Flux transformations and the Einstein-Helfand thermal window operation against vectors
    from the reference's own source (tests/golden/make_reference_flux_goldens.py)."""
    from oracle import dynamics as od

    g = _load("reference_flux.json")
    batch = {sp: {k: np.asarray(v, dtype=np.float32) for k, v in d.items()}
             for sp, d in g["inputs"].items()}
    assert np.array_equal(ot.momentum_flux_transform_batch(batch), np.asarray(g["momentum_flux"]))
    assert np.array_equal(ot.thermal_flux_transform_batch(batch), np.asarray(g["thermal_flux"]))
    assert np.array_equal(ot.integrated_heat_current_transform_batch(batch),
                          np.asarray(g["integrated_heat_current"]))
    e = g["eh_thermal_ensemble_operation"]
    window = np.asarray(e["window"])
    N = window.shape[0]
    plan = {"batch_size": N, "n_batches": 1, "remainder": 0, "minibatch": False}
    msd = od.eh_ionic_msd(window[None], plan, N, 1, np.arange(N), e["prefactor"])
    np.testing.assert_allclose(msd, np.asarray(e["msd"]), rtol=1e-14)


def test_reference_run_planner(ref_run):
    from lammps_analysis_b200.planner import plan_batches

    n_minibatch = 0
    for c in ref_run["planner"]:
        db = _FakeDB(c["nbytes"], c["rows"], c["cols"])
        mm = op.MemoryManager(data_path=["x"], database=db, memory_fraction=0.5,
                              scale_function=c["scale_function"], memory=c["memory"])
        assert list(mm.get_batch_size()) == c["get_batch_size"]
        loops, minibatch = mm.get_ensemble_loop(c["data_range"], c["correlation_time"])
        assert (loops, minibatch) == (c["ensemble_loop"], c["minibatch"])
        assert (mm.batch_size, mm.n_batches, mm.remainder) == (
            c["batch_size"], c["n_batches"], c["remainder"])
        assert (mm.atom_batch_size, mm.n_atom_batches, mm.atom_remainder) == (
            c["atom_batch_size"], c["n_atom_batches"], c["atom_remainder"])
        # the product planner as well
        p = plan_batches([(c["rows"], c["cols"], c["nbytes"])], c["data_range"],
                         c["correlation_time"], c["scale_function"], memory=c["memory"],
                         memory_fraction=0.5)
        assert (p.batch_size, p.n_batches, p.remainder, p.ensemble_loop, p.minibatch) == (
            c["batch_size"], c["n_batches"], c["remainder"], c["ensemble_loop"], c["minibatch"])
        if c["minibatch"]:
            assert (p.atom_batch_size, p.n_atom_batches, p.atom_remainder) == (
                c["atom_batch_size"], c["n_atom_batches"], c["atom_remainder"])
        n_minibatch += c["minibatch"]
    assert 10 < n_minibatch < len(ref_run["planner"]) - 10


def test_reference_run_windows_and_batch_slices(ref_run):
    for w in ref_run["ensemble_windows"]:
        got = [[s, min(e, w["data_size"])] for s, e in
               op.iter_ensembles(w["data_size"], w["data_range"], w["correlation_time"])]
        assert got == w["windows"]
    plain, mini = ref_run["batch_slices"]
    plan = dict(plain["plan"], minibatch=False)
    got = [[[None, None], [start, stop]] for _, start, stop, _ in op.iter_batches(plan)]
    assert got == plain["slices"]
    got = [[[sel.start, sel.stop], [start, stop]] for sel, start, stop, _ in
           op.iter_batches(mini["plan"])]
    assert got == mini["slices"]


def test_reference_run_fits(ref_run):
    from lammps_analysis_b200.calculators.coordination_number_calculation import \
        golden_section_search
    from lammps_analysis_b200.calculators.einstein_diffusion_coefficients import \
        fit_einstein_curve
    from oracle import coordination as oc

    g = ref_run["fit_einstein_curve"]
    x, y = np.asarray(g["x"]), np.asarray(g["y"])
    popt, pcov, grads, _ = od.fit_einstein_curve(x, y, g["fit_max_index"])
    np.testing.assert_allclose(popt, g["popt"], rtol=1e-10)
    np.testing.assert_allclose(grads, g["gradients"], rtol=1e-10)
    popt2, pcov2, grads2, _ = fit_einstein_curve(x, y, g["fit_max_index"])   # closed form
    np.testing.assert_allclose(popt2, g["popt"], rtol=1e-6)
    np.testing.assert_allclose(np.diag(pcov2), np.diag(np.asarray(g["pcov"])), rtol=1e-4)
    np.testing.assert_allclose(grads2, g["gradients"], rtol=1e-6)
    s = ref_run["golden_section_search"]
    data = [np.asarray(s["r"]), np.asarray(s["g"])]
    assert list(oc.golden_section_search(data, s["a"], s["b"])) == s["result"]
    assert list(golden_section_search(data, s["a"], s["b"])) == s["result"]
    e = ref_run["einstein_ensemble_operation"]
    ens = np.asarray(e["ensemble"])
    plan = dict(batch_size=ens.shape[1], n_batches=1, remainder=0, minibatch=False)
    msd, count = od.einstein_msd(ens, plan, ens.shape[1], 1, np.arange(ens.shape[1]))
    np.testing.assert_allclose(msd, e["msd"], rtol=1e-13)
    assert count == e["count"] + 1          # + 1 per window is added by run_calculator (:244)


def test_reference_run_rdf_normalisation_and_cn(ref_run, tmp_path):
    """ideal_correction / _calculate_prefactor / _ang_to_nm and the coordination-number
    post-processing executed from upstream source vs the oracle AND the product calculators."""
    from collections import OrderedDict
    from types import SimpleNamespace

    from lammps_analysis_b200.calculators.coordination_number_calculation import (
        CoordinationNumbers)
    from lammps_analysis_b200.calculators.radial_distribution_function import (
        RadialDistributionFunction)
    from lammps_analysis_b200.project import Computation, Species
    from lammps_analysis_b200.units import REAL
    from oracle import coordination as oc
    from oracle import rdf as orc

    def decode(v):
        v = np.asarray(v, dtype=float)
        out = v.copy()
        out[v == -1.0] = np.nan
        out[v == -2.0] = np.inf
        return out

    for c in ref_run["rdf_normalisation"]:
        corr = orc.ideal_correction(c["cutoff"], c["nbins"], c["box"][0])
        want = decode(c["ideal_correction"])
        np.testing.assert_allclose(np.nan_to_num(corr, nan=-1.0), np.nan_to_num(want, nan=-1.0),
                                   rtol=1e-13)
        # product: prefactor and x axis
        exp = SimpleNamespace(box_array=c["box"], volume=float(np.prod(c["box"])), units=REAL,
                              species=OrderedDict((k, Species(k, n))
                                                  for k, n in c["n_particles"].items()))
        calc = RadialDistributionFunction.__new__(RadialDistributionFunction)
        calc.experiment = exp
        calc.args = SimpleNamespace(cutoff=c["cutoff"], number_of_bins=c["nbins"],
                                    atom_selection=np.s_[:], number_of_configurations=c["n_configs"])
        counts = {k: np.ones(c["nbins"], dtype=np.int64) for k in c["prefactor"]}
        ref = orc.rdf_normalise(counts, c["n_particles"], c["box"], c["cutoff"], c["nbins"],
                                c["n_configs"], 1e-10)
        for key, pref in c["prefactor"].items():
            want = decode(pref)
            with np.errstate(all="ignore"):
                got = calc._calculate_prefactor(key)
            fin = np.isfinite(want)
            np.testing.assert_allclose(got[fin], want[fin], rtol=1e-13)
            assert np.array_equal(np.isfinite(got), fin)
            np.testing.assert_allclose(np.asarray(ref[key]["y"])[fin], want[fin], rtol=1e-13)
            np.testing.assert_allclose(ref[key]["x"], c["x"], rtol=1e-15)

    g = ref_run["coordination_numbers"]
    rdf_dict = {"Na_Cl": {"x": g["x"], "y": g["y"]}}
    # oracle (density passed through n_particles / volume)
    ref = oc.coordination_numbers(rdf_dict, {"Na": g["density"]}, 1.0, number_of_shells=2)
    np.testing.assert_allclose(ref["Na_Cl"]["cn"], g["cn"], rtol=1e-13)
    for k, v in g["values"].items():
        assert ref["Na_Cl"][k] == pytest.approx(v, rel=1e-12, abs=1e-15)
    # product calculator on the same RDF
    calc = CoordinationNumbers.__new__(CoordinationNumbers)
    calc._queued_data = []
    calc.experiment = SimpleNamespace(volume=1.0, units=SimpleNamespace(length=1e-9),
                                      species={"Na": Species("Na", g["density"])})
    calc.args = SimpleNamespace(savgol_order=2, savgol_window_length=17, number_of_shells=2)
    calc.rdf_data = Computation("Radial_Distribution_Function", "x", {},
                                OrderedDict(rdf_dict))
    calc.run_calculator()
    key, data = calc._queued_data[0]
    assert key == "Na_Cl"
    np.testing.assert_allclose(data["cn"], g["cn"], rtol=1e-13)
    for k, v in g["values"].items():
        assert data[k] == pytest.approx(v, rel=1e-12, abs=1e-15)
