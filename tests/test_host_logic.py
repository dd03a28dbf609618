"""CPU-only tests of the host side: planner, store, ingest, fits, cache protocol, C ABI exports."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_libmdk_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports everything include/mdk.h declares."""
    from lammps_analysis_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "mdk.h")).read()
    declared = set(re.findall(r"\b(mdk_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("mdk_stream_t")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in mdk.h but not exported"
    assert set(_lib.PROTOTYPES) | {"mdk_last_error"} == declared
    assert _lib.load().mdk_version() == 100
    # constants mirrored on the Python side
    from lammps_analysis_b200 import kernels as K

    assert K.RDF_SUBTILE == int(re.search(r"#define MDK_RDF_SUBTILE (\d+)", hdr).group(1))
    assert _lib.MDK_RDF_WRAPPED == int(re.search(r"#define MDK_RDF_WRAPPED (\d+)", hdr).group(1))
    assert _lib.MDK_RDF_EXACT_DIV == int(re.search(r"#define MDK_RDF_EXACT_DIV (\d+)", hdr).group(1))


def test_c_abi_rejects_bad_arguments_without_a_gpu():
    """Argument checks run before any CUDA call: error code < 0 and a message, no launch."""
    from lammps_analysis_b200 import _lib

    lib = _lib.load()
    one = ctypes.c_double(1.0)
    buf = (ctypes.c_float * 16)()
    ptr = ctypes.cast(buf, ctypes.c_void_p)
    cases = [
        lib.mdk_flux_sum(None, 1, 1, 3, 0, None, None, None, None),            # null pointers
        lib.mdk_flux_sum(ptr, 1, 1, 2, 0, None, None, ptr, None),              # ncomp < 3
        lib.mdk_flux_sum(ptr, 1, 1, 6, 4, None, None, ptr, None),              # comp0 + 3 > ncomp
        lib.mdk_flux_sum(ptr, 1, 1, 3, 0, None, ptr, ptr, None),               # w2 without w1
        lib.mdk_thermal_flux(ptr, ptr, ptr, None, 1, 1, ptr, None),            # null pe
        lib.mdk_ionic_current(ptr, 1, 1, ctypes.cast(ctypes.byref(one), ctypes.c_void_p), 7, ptr,
                              None),                                           # bad q_mode
        lib.mdk_msd_dense(ptr, 1, 4, 0, 1, 0, 3, 4, ptr, None),                # windows exceed T
        lib.mdk_rdf_thresholds(ctypes.c_float(-1.0), 10, ptr, ptr),            # cutoff <= 0
    ]
    assert all(rc < 0 for rc in cases), cases
    assert lib.mdk_last_error().decode() != ""
    # empty work is not an error and launches nothing
    assert lib.mdk_flux_sum(ptr, 0, 1, 3, 0, None, None, ptr, None) == 0
    assert lib.mdk_thermal_flux(ptr, ptr, ptr, ptr, 0, 1, ptr, None) == 0


def test_rdf_thresholds_reproduce_double_step_binning():
    """mdk_rdf_thresholds (host helper, no GPU): bin(d2) via the table equals
    int(min(double(d)/step, nbins-1)) for every sampled fp32 distance."""
    from lammps_analysis_b200 import kernels as K

    cutoff, nbins = np.float32(15.9), 1590
    thr, cut2 = K.rdf_thresholds(float(cutoff), nbins)
    assert thr[0] == 0 and np.isinf(thr[nbins]) and np.all(np.diff(thr[:nbins]) > 0)
    rng = np.random.default_rng(0)
    d2 = (rng.random(2_000_000).astype(np.float32) * np.float32(cutoff) ** 2).astype(np.float32)
    # adversarial samples: the thresholds themselves and their neighbours
    d2 = np.concatenate([d2, thr[1:nbins], np.nextafter(thr[1:nbins], np.float32(0)),
                         np.nextafter(thr[1:nbins], np.float32(np.inf))]).astype(np.float32)
    d = np.sqrt(d2)                       # correctly rounded fp32 sqrt
    inside = d < cutoff
    assert np.array_equal(inside, d2 < np.float32(cut2))
    step = float(cutoff) / nbins
    want = np.minimum(d[inside].astype(np.float64) / step, nbins - 1).astype(np.int64)
    got = np.searchsorted(thr[1:nbins], d2[inside], side="right")
    assert np.array_equal(got, want)


def test_product_planner_matches_oracle_planner():
    from lammps_analysis_b200.planner import plan_batches
    from oracle.planner import MemoryManager

    class DB:
        def __init__(self, size):
            self.size = size

        def get_data_size(self, item):
            return self.size

    rng = np.random.default_rng(0)
    specs = [{"linear": {"scale_factor": 150}}, {"linear": {"scale_factor": 5}},
             {"quadratic": {"inner_scale_factor": 5, "outer_scale_factor": 10}}, None]
    for _ in range(1500):
        rows, cols = int(rng.integers(1, 5000)), int(rng.integers(2, 20000))
        size = (rows, cols, rows * cols * 12)
        mem = float(10 ** rng.uniform(2, 11))
        N, ct = int(rng.integers(1, 600)), int(rng.integers(1, 6))
        sf = specs[int(rng.integers(0, 4))]
        mm = MemoryManager(data_path=[0], database=DB(size), memory_fraction=0.5,
                           scale_function=sf, memory=mem)
        mm.get_batch_size()
        loops, minibatch = mm.get_ensemble_loop(N, ct)
        p = plan_batches([size], N, ct, sf, memory=mem, memory_fraction=0.5)
        assert (p.batch_size, p.n_batches, p.remainder, p.ensemble_loop, p.minibatch) == (
            mm.batch_size, mm.n_batches, mm.remainder, loops, minibatch)
        if minibatch:
            assert (p.atom_batch_size, p.n_atom_batches, p.atom_remainder) == (
                mm.atom_batch_size, mm.n_atom_batches, mm.atom_remainder)


def test_survey_worked_plans():
    """SURVEY.md A.5 worked plans at 60 GB."""
    from lammps_analysis_b200.planner import plan_batches

    ein = {"linear": {"scale_factor": 150}}
    p = plan_batches([(500, 5000, 500 * 5000 * 12)], 500, 1, ein, memory=60e9, memory_fraction=0.5)
    assert (p.batch_size, p.n_batches, p.ensemble_loop, p.minibatch) == (5000, 1, 4500, False)
    p = plan_batches([(500_000, 2000, 500_000 * 2000 * 12)], 500, 1, ein, memory=60e9,
                     memory_fraction=0.5)
    assert p.minibatch and p.batch_size == 2000 and p.atom_batch_size == 250_000
    assert p.n_atom_batches == 2 and p.ensemble_loop == 1500
    p = plan_batches([(1, 10000, 10000 * 12)], 500, 1, {"linear": {"scale_factor": 5}},
                     memory=60e9, memory_fraction=0.5)
    assert (p.batch_size, p.n_batches, p.ensemble_loop) == (10000, 1, 9500)


def test_plan_windows_matches_oracle_iteration():
    from lammps_analysis_b200.engine import plan_windows
    from lammps_analysis_b200.planner import plan_batches
    from oracle.planner import iter_batches, iter_ensembles

    for (A, T, N, ct, mem) in [(64, 400, 50, 1, 60e9), (60, 1000, 100, 2, 2.0e5 * 60 / 64),
                               (200, 400, 120, 1, 150 * 200 * 12 * 100 / 0.5)]:
        p = plan_batches([(A, T, A * T * 12)], N, ct, {"linear": {"scale_factor": 150}},
                         memory=mem, memory_fraction=0.5).as_dict()
        want = []
        for atom_sel, start, stop, size in iter_batches(p):
            wins = [(s, e) for s, e in iter_ensembles(size, N, ct) if e <= size]
            if wins:
                lo = 0 if atom_sel == slice(None) else atom_sel.start
                hi = A if atom_sel == slice(None) else atom_sel.stop
                want.append((lo, hi, start, size, len(wins)))
        assert plan_windows(p, N, ct, A) == want


def test_store_roundtrip_and_float32_rounding(tmp_path):
    from lammps_analysis_b200.store import TrajectoryStore

    st = TrajectoryStore(str(tmp_path / "db"))
    x = np.random.default_rng(0).normal(size=(5, 7, 3))
    st.add_dataset("Na/Positions", (5, 7, 3))
    st.add_data("Na/Positions", x[:, :4], start=0)
    st.add_data("Na/Positions", x[:, 4:], start=4)
    assert st.check_existence("Na/Positions") and not st.check_existence("Na/Velocities")
    assert st.get_data_size("Na/Positions") == (5, 7, 5 * 7 * 3 * 4)
    got = st.load_data("Na/Positions", np.s_[1:3, 2:5])
    assert got.dtype == np.float64
    assert np.array_equal(got, x.astype(np.float32).astype(np.float64)[1:3, 2:5])
    st2 = TrajectoryStore(str(tmp_path / "db"))          # re-open from disk
    assert np.array_equal(st2.host("Na/Positions"), x.astype(np.float32))
    st2.resize_dataset("Na/Positions", 10)
    assert st2.shape("Na/Positions") == (5, 10, 3)
    assert np.array_equal(st2.host("Na/Positions")[:, :7], x.astype(np.float32))


def test_lammps_dump_roundtrip_with_unsorted_ids(tmp_path):
    from lammps_analysis_b200.file_io import LAMMPSTrajectoryFile, write_lammps_dump
    from lammps_analysis_b200.synthetic import nacl_trajectory

    data, box = nacl_trajectory(64, 4, 12.0, 1)
    path = str(tmp_path / "t.lammpstraj")
    write_lammps_dump(path, data, box, step_stride=5)
    # shuffle atom rows inside every frame: the reader must sort by id
    lines = open(path).read().splitlines()
    rng = np.random.default_rng(0)
    out = []
    for f in range(4):
        blk = lines[f * 73:(f + 1) * 73]
        rows = blk[9:]
        rng.shuffle(rows)
        out += blk[:9] + list(rows)
    open(path, "w").write("\n".join(out) + "\n")
    results = {}
    for native in (True, False):     # C++ tokenizer of libmdk and the pure-Python parser
        reader = LAMMPSTrajectoryFile(path, native=native)
        meta = reader.metadata
        assert meta.n_configurations == 4 and meta.sample_rate == 5 and meta.box_l == [12.0] * 3
        assert [(s.name, s.n_particles) for s in meta.species_list] == [("Na", 32), ("Cl", 32)]
        chunks = list(reader.get_configurations_generator(3))
        for sp in ("Na", "Cl"):
            for prop in ("Positions", "Velocities"):
                got = np.concatenate([c.data[sp][prop] for c in chunks], axis=1)
                assert np.array_equal(got.astype(np.float32), data[sp][prop])
                results[(native, sp, prop)] = got
    for (native, sp, prop), got in results.items():
        assert np.array_equal(got, results[(not native, sp, prop)])   # bit-identical float64


def test_native_lammps_reader_number_formats(tmp_path):
    """The native tokenizer (exact fast path + strtod fallback, frame-parallel) returns the same
    float64 as Python's float() for every spelling: exponents, long mantissas, subnormals,
    signed zero, nan / inf, bare dots; CRLF line ends; no trailing newline; several threads."""
    import os

    from lammps_analysis_b200.file_io import LAMMPSTrajectoryFile

    toks = ["1e-5", "-0.0", "123456789012345678901", "1.7976931348623157e308", "4.9e-324",
            "0.1", "+3.", ".5", "1E+3", "2.5e-23", "9007199254740993", "0.30000000000000004",
            "1e22", "1e23", "123456.789e-30", "-7.25", "nan", "inf", "-inf", "1e-400",
            "3.14159265358979323846264338327950288", "00012.50", "6.02214076e23", "1e", "-"]
    n_atoms = 5
    rng = np.random.default_rng(4)
    frames = []
    for f in range(7):
        rows = []
        for a in rng.permutation(n_atoms):
            vals = [toks[int(k)] for k in rng.integers(0, len(toks), 6)]
            rows.append(f"{a + 1} {1 + a % 2} {'Na' if a % 2 == 0 else 'Cl'} " + " ".join(vals))
        hdr = ["ITEM: TIMESTEP", str(10 * f), "ITEM: NUMBER OF ATOMS", str(n_atoms),
               "ITEM: BOX BOUNDS pp pp pp", "0 9.5", "0 9.5", "0 9.5",
               "ITEM: ATOMS id type element x y z vx vy vz"]
        frames.append("\r\n".join(hdr + rows) if f % 2 else "\n".join(hdr + rows))
    path = str(tmp_path / "odd.lammpstraj")
    with open(path, "w", newline="") as fh:
        fh.write("\n".join(frames))          # no trailing newline
    results = {}
    for native, threads in ((True, "1"), (True, "3"), (False, "1")):
        os.environ["MDK_INGEST_THREADS"] = threads
        try:
            reader = LAMMPSTrajectoryFile(path, native=native)
            assert reader.metadata.n_configurations == 7
            chunks = list(reader.get_configurations_generator(3))
        finally:
            os.environ.pop("MDK_INGEST_THREADS", None)
        results[(native, threads)] = {
            (sp, prop): np.concatenate([c.data[sp][prop] for c in chunks], axis=1)
            for sp in ("Na", "Cl") for prop in ("Positions", "Velocities")}
    ref = results[(False, "1")]
    for key, got in results.items():
        for k in ref:
            a, b = got[k], ref[k]
            assert a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)), (key, k)
            assert np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)]), (key, k)
            assert np.array_equal(np.signbit(a), np.signbit(b)), (key, k)


def test_native_lammps_reader_rejects_truncated_file(tmp_path):
    from lammps_analysis_b200._lib import MdkError
    from lammps_analysis_b200.file_io import LAMMPSTrajectoryFile, write_lammps_dump
    from lammps_analysis_b200.synthetic import nacl_trajectory

    data, box = nacl_trajectory(8, 3, 6.0, 1)
    path = str(tmp_path / "t.lammpstraj")
    write_lammps_dump(path, data, box)
    lines = open(path).read().splitlines()
    open(path, "w").write("\n".join(lines[:-3]) + "\n")
    with pytest.raises(MdkError):
        LAMMPSTrajectoryFile(path).metadata


def test_closed_form_einstein_fit_matches_curve_fit():
    from lammps_analysis_b200.calculators.einstein_diffusion_coefficients import fit_einstein_curve
    from oracle import dynamics as od

    rng = np.random.default_rng(0)
    x = np.arange(150) * 0.02e-12
    y = 6 * 1.3e-9 * x + 1e-21 * np.sqrt(np.arange(150)) + rng.normal(0, 1e-23, 150)
    a = od.fit_einstein_curve(x, y, 149)
    b = fit_einstein_curve(x, y, 149)
    np.testing.assert_allclose(b[0], a[0], rtol=1e-6)
    np.testing.assert_allclose(np.sqrt(np.diag(b[1])), np.sqrt(np.diag(a[1])), rtol=1e-4)
    np.testing.assert_allclose(b[2], a[2], rtol=1e-6)
    # reference known answers (CI/unit_tests/utils/test_calculator_helper_methods.py:42-69)
    xs = np.linspace(0, 1000, 1000)
    assert fit_einstein_curve(xs, 5 * xs + 3, 999)[0][0] == pytest.approx(5.0, 0.01)
    ys = np.exp(-0.05 * xs) * xs**2 + 5 * xs + 3
    assert fit_einstein_curve(xs, ys, 999)[0][0] == pytest.approx(5.0, 0.01)


def test_golden_section_search_matches_recursive_oracle():
    from lammps_analysis_b200.calculators.coordination_number_calculation import \
        golden_section_search
    from oracle import coordination as oc

    r = np.linspace(0.05, 1.5, 600)
    g = 1 + np.exp(-3 * r) * np.cos(14 * r) * 2
    assert golden_section_search([r, g], r[300], r[80]) == oc.golden_section_search(
        [r, g], r[300], r[80])


def test_cache_protocol_with_fake_calculator(tmp_path):
    """calculator.py:52-148: identical args + version -> stored object, no recompute."""
    from dataclasses import dataclass

    from lammps_analysis_b200.calculators.calculator import Calculator, call
    from lammps_analysis_b200.file_io import ScriptInput
    from lammps_analysis_b200.project import Project

    runs = []

    @dataclass
    class Args:
        k: int
        sel: object

    class Fake(Calculator):
        analysis_name = "Fake"

        @call
        def __call__(self, k: int = -1, sel=np.s_[:]):
            self.args = Args(k=k, sel=sel)

        def run_calculator(self):
            if self.args.k == -1:
                self.args.k = 7          # default resolved during the run
            runs.append(self.args.k)
            self.queue_data({"value": self.args.k * 2}, subjects=["Na", "Na"])
            self.queue_data({"value": 1}, subjects=["System"])

    project = Project("p", storage_path=str(tmp_path))
    exp = project.add_experiment("e", timestep=1.0, temperature=1.0, units="si")
    exp.add_data(ScriptInput({"Na": {"Positions": np.zeros((3, 2, 3))}}, [1, 1, 1]))
    a = Fake(experiment=exp)()
    b = Fake(experiment=exp)()
    assert runs == [7] and a.id == b.id
    assert a["Na_Na"] == {"value": 14} and a.keys() == ["Na_Na", "System"]
    assert a.computation_parameter == {"k": -1, "sel": "slice(None, None, None)", "version": 1}
    with pytest.raises(KeyError):
        a["Cl"]
    Fake(experiment=exp)(k=3)
    assert runs == [7, 3]
    exp.add_data(ScriptInput({"Na": {"Positions": np.zeros((3, 2, 3))}}, [1, 1, 1], name="more"))
    assert exp.version == 2 and exp.number_of_configurations == 6
    Fake(experiment=exp)()               # new data -> new version -> recompute
    assert runs == [7, 3, 7]
    both = Fake(experiments=[exp])()     # project-style call returns a dict
    assert list(both) == ["e"]
    # a re-opened project finds the stored computation
    again = Project("p", storage_path=str(tmp_path))
    assert Fake(experiment=again.experiments["e"])().id == both["e"].id


def test_reopened_project_keeps_units_and_results(tmp_path):
    """experiment.py:188-191: the unit system is a stored property -- a re-opened experiment
    must not fall back to REAL; results come back from a JSON + float64 record (never
    unpickled), with identical structure and values."""
    import sqlite3

    from lammps_analysis_b200.file_io import ScriptInput
    from lammps_analysis_b200.project import Project, decode_results, encode_results
    from lammps_analysis_b200.units import METAL, REAL, Units

    project = Project("u", storage_path=str(tmp_path))
    exp = project.add_experiment("e", timestep=0.5, temperature=3.0, units="metal")
    exp.add_data(ScriptInput({"Na": {"Positions": np.zeros((3, 2, 3))}}, [1, 1, 1]))
    custom = Units(time=2.0, length=3.0, energy=4.0, NkTV2p=5.0, boltzmann=6.0, temperature=7.0,
                   pressure=8.0)
    project.add_experiment("c", timestep=0.5, temperature=3.0, units=custom)
    series = np.linspace(0.0, 1.0, 500)
    results = {"Na": {"x": series.tolist(), "y": [float("nan"), float("inf")] + [0.25] * 30,
                      "value": 1.5, "list": [1.0, 2.0], "nested": {"k": [3, 4]}, "none": None},
               "System": {"msd": (series**2).tolist(), "n": 3}}
    params = {"a": 1, "version": exp.version}
    project.store_computation("Calc", "e", params, results, metadata={"tie_report": {"ties": 0}})

    again = Project("u", storage_path=str(tmp_path))
    assert again.experiments["e"].units == METAL and again.experiments["e"].units != REAL
    assert again.experiments["c"].units == custom
    assert again.experiments["e"].time_step == 0.5
    # units given explicitly on re-creation still win over nothing stored (new experiment)
    assert again.add_experiment("new", timestep=1.0, temperature=1.0).units == REAL
    got = again.find_computation("Calc", "e", params)
    assert got.keys() == ["Na", "System"] and got.metadata == {"tie_report": {"ties": 0}}
    assert got["Na"]["x"] == series.tolist() and got["System"]["msd"] == (series**2).tolist()
    assert np.isnan(got["Na"]["y"][0]) and np.isinf(got["Na"]["y"][1])
    assert got["Na"]["value"] == 1.5 and got["Na"]["list"] == [1.0, 2.0]
    assert got["Na"]["nested"] == {"k": [3, 4]} and got["Na"]["none"] is None
    assert got["System"]["n"] == 3
    # the stored record is the documented format, not a pickle
    blob = sqlite3.connect(os.path.join(str(tmp_path), "MDSuite_Project_u", "project.db")) \
        .execute("SELECT results FROM computations").fetchone()[0]
    assert bytes(blob[:8]) == b"MDKR0001" and b"pickle" not in bytes(blob[:64])
    assert decode_results(encode_results(results))["Na"]["x"] == series.tolist()
    with pytest.raises(ValueError):
        decode_results(b"\x80\x05" + bytes(30))     # a pickle stream is refused, not loaded


def test_lammps_column_names_follow_the_reference():
    """lammps_trajectory_files.py:39-66 + mdsuite_properties.py:80-81."""
    from lammps_analysis_b200.file_io import var_names

    assert var_names["Kinetic_Energy"] == ["c_KE"] and var_names["Potential_Energy"] == ["c_PE"]
    assert "KE" not in var_names and "PE" not in var_names
    assert len(var_names["Stress"]) == 6


def _write_flux_log(path, table, columns, n_header_lines=2, tail=True):
    with open(path, "w") as fh:
        for k in range(n_header_lines - 1):
            fh.write("# LAMMPS flux output, header line %d\n" % k)
        fh.write(" ".join(columns) + "\n")
        for row in table:
            fh.write(" ".join(repr(float(v)) for v in row) + "\n")
        if tail:
            fh.write("Loop time of 12.3 on 4 procs for 1000 steps\n1.0 2.0\n")


def test_lammps_flux_file_reader(tmp_path):
    """lammps_flux_files.py:41-156: header lines, column-name mapping (incl. a custom map),
    reading stops where the column count changes, observables stored as (1, n_steps, n_dims)."""
    from lammps_analysis_b200.file_io import LAMMPSFluxFile
    from lammps_analysis_b200.project import Project

    rng = np.random.default_rng(3)
    cols = ["time", "temp", "c_flux_thermal[1]", "c_flux_thermal[2]", "c_flux_thermal[3]",
            "pxy", "pxz", "pyz", "v_extra"]
    table = rng.normal(size=(37, len(cols)))
    path = str(tmp_path / "flux.lmp")
    _write_flux_log(path, table, cols, n_header_lines=3)
    reader = LAMMPSFluxFile(path, sample_rate=5, box_l=[10.0, 11.0, 12.0], n_header_lines=3,
                            custom_data_map={"Extra": ["v_extra"]})
    meta = reader.metadata
    assert meta.n_configurations == 37 and meta.sample_rate == 5
    assert [s.name for s in meta.species_list] == ["Observables"]
    assert {p.name: p.n_dims for p in meta.species_list[0].properties} == {
        "Temperature": 1, "Time": 1, "Thermal_Flux": 3, "Stress_Visc": 3, "Extra": 1}
    project = Project("flux", storage_path=str(tmp_path))
    exp = project.add_experiment("e", timestep=0.001, temperature=300.0, units="real",
                                 simulation_data=reader)
    assert exp.number_of_configurations == 37 and exp.box_array == [10.0, 11.0, 12.0]
    assert "Observables" not in exp.species and exp.sample_rate == 5
    visc = exp.store.host("Observables/Stress_Visc")
    assert visc.shape == (1, 37, 3) and visc.dtype == np.float32
    assert np.array_equal(visc[0], table[:, 5:8].astype(np.float32))
    assert np.array_equal(exp.store.host("Observables/Extra")[0, :, 0],
                          table[:, 8].astype(np.float32))
    with pytest.raises(ValueError):
        empty = str(tmp_path / "empty.lmp")
        open(empty, "w").write("# a\ntime temp\n")
        LAMMPSFluxFile(empty, 1, [1, 1, 1]).metadata


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lammps_analysis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_kernels_refuse_cpu_tensors():
    import torch

    from lammps_analysis_b200 import kernels as K
    from lammps_analysis_b200._lib import MdkError

    with pytest.raises(MdkError):
        K.msd_windowed(torch.zeros(2, 4, 3), 0, 2, 0, 1, 1, torch.zeros(2, dtype=torch.int32), 2,
                       torch.zeros(2, dtype=torch.float64))


def test_row_block_helpers():
    """engine._row_blocks / _merge_blocks: how a launch range is cut along upload blocks."""
    from lammps_analysis_b200.engine import _merge_blocks, _row_blocks

    blocks = [(0, 10, "a"), (10, 20, "b"), (20, 25, "c")]
    assert list(_row_blocks(3, 22, blocks)) == [(3, 10, "a"), (10, 20, "b"), (20, 22, "c")]
    assert list(_row_blocks(12, 15, blocks)) == [(12, 15, "b")]
    assert list(_row_blocks(4, 9, None)) == [(4, 9, None)]
    many = [(10 * i, 10 * i + 10, i) for i in range(23)]
    merged = _merge_blocks(many, 4)
    assert len(merged) == 4 and merged[0] == (0, 60, 5) and merged[-1] == (180, 230, 22)
    assert [b[0] for b in merged[1:]] == [b[1] for b in merged[:-1]]      # contiguous
    assert _merge_blocks(blocks, 4) == blocks and _merge_blocks(None, 4) is None
