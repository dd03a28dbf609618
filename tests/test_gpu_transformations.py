"""The transformation driver (a13): input resolution chain, dataset placement, extension after
new data.  The cases of the reference's CI/unit_tests/transformations/test_transformator_parent.py
(:66-237) restated against this package, plus value checks of the extension the reference only
smoke-tests."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _exp(tmp_path, name="TestExp", timestep=0.123, box=None, persist=True):
    from lammps_analysis_b200.project import Project

    project = Project(name, storage_path=str(tmp_path), persist=persist)
    exp = project.add_experiment(name=name, timestep=timestep, temperature=300.0)
    if box is not None:
        exp.box_array = list(box)
    return project, exp


def _load_pos(exp, seed, sp_name="test_species", unwrapped=True, n_config=100, n_part=10,
              box=(1.1, 2.2, 3.3), name="pos"):
    """test_transformator_parent.py:42-65 (load_pos_into_exp)."""
    from lammps_analysis_b200.file_io import ScriptInput

    rng = np.random.default_rng(seed)
    pos = rng.random((n_config, n_part, 3))
    prop = "Unwrapped_Positions" if unwrapped else "Positions"
    if not unwrapped:
        pos /= np.max(pos)
    exp.add_data(ScriptInput({sp_name: {prop: pos}}, box, name=name))
    return sp_name, pos.astype(np.float32)


def _test_trafos():
    """mdsuite/transformations/test_trafos.py:38-96: 'use' every input, return 516s, carry 17."""
    import torch
    from lammps_analysis_b200.transformations import MultiSpeciesTrafo, SingleSpeciesTrafo

    class TestSingleSpecies(SingleSpeciesTrafo):
        def transform_batch(self, batch, carryover=None):
            for prop in self.input_properties:
                assert batch[prop.name] is not None
            assert carryover is None or carryover == 17
            shape = batch[self.input_properties[0].name].shape
            return torch.full((1, shape[1], self.output_property.n_dims), 516.0,
                              dtype=torch.float32, device="cuda"), 17

    class TestMultispecies(MultiSpeciesTrafo):
        def transform_batch(self, batch, carryover=None):
            for props in batch.values():
                for prop in self.input_properties:
                    assert props[prop.name] is not None
            sp = next(iter(batch))
            shape = batch[sp][self.input_properties[0].name].shape
            return torch.full((1, shape[1], self.output_property.n_dims), 516.0,
                              dtype=torch.float32, device="cuda"), 17

    return TestSingleSpecies, TestMultispecies


def test_automatic_coordinate_unwrapping(tmp_path, cuda):
    """:68-91 -- VelocityFromPositions needs Unwrapped_Positions; only wrapped Positions are
    stored, so the property -> transformation table is consulted: UnwrapViaIndices cannot find
    Box_Images, the fallback CoordinateUnwrapper runs.  Values are checked as well."""
    from oracle import transformations as ot

    box = [1.1, 2.2, 3.3]
    project, exp = _exp(tmp_path, box=box)
    sp, pos = _load_pos(exp, 1, unwrapped=False)
    exp.run.VelocityFromPositions()
    unw = exp.load_matrix(property_name="Unwrapped_Positions", species=[sp])
    assert unw.shape == (10, 100, 3)
    ref_unw = ot.run_unwrap(np.swapaxes(pos, 0, 1).copy(), np.array(box), batch_size=100)
    assert np.array_equal(unw.astype(np.float32), ref_unw)
    vel = exp.store.host(f"{sp}/Velocities_From_Positions")
    dt = np.float32(0.123) * np.float32(1)
    want = (ref_unw[:, 1:] - ref_unw[:, :-1]) / dt
    assert np.array_equal(vel[:, :-1], want) and np.array_equal(vel[:, -1], want[:, -1])


def test_full_transformation_with_values(tmp_path, cuda):
    """:94-120."""
    project, exp = _exp(tmp_path)
    sp, pos = _load_pos(exp, 2)
    exp.run.VelocityFromPositions()
    vels = exp.load_matrix(property_name="Velocities_From_Positions", species=[sp])
    pos = np.swapaxes(pos, 0, 1).astype(np.float64)
    ref = (pos[:, 1:] - pos[:, :-1]) / 0.123
    ref = np.concatenate((ref, ref[:, -1:]), axis=1)
    np.testing.assert_almost_equal(vels, ref, decimal=4)


def test_not_found_errors(tmp_path, cuda):
    """:123-143 -- a made-up input property raises CannotFindPropertyError and leaves no empty
    output dataset behind."""
    from lammps_analysis_b200.transformations import CannotFindPropertyError, PropertyInfo

    single, multi = _test_trafos()
    project, exp = _exp(tmp_path, name="TestExp1234", timestep=12345)
    _load_pos(exp, 3)
    for cls, name in ((single, "test1"), (multi, "test2")):
        trafo = cls(input_properties=[PropertyInfo("MadeUpProperty", 42)],
                    output_property=PropertyInfo(name, 2))
        with pytest.raises(CannotFindPropertyError):
            exp.cls_transformation_run(trafo)
    assert not any(p.endswith(("test1", "test2")) for p in exp.store.paths())


def test_save_to_correct_name(tmp_path, cuda):
    """:146-174 -- species datasets under the species, system observables under Observables."""
    from lammps_analysis_b200.transformations import PropertyInfo, mdsuite_properties

    single, multi = _test_trafos()
    project, exp = _exp(tmp_path, timestep=12345)
    sp, _ = _load_pos(exp, 4)
    exp.cls_transformation_run(single(input_properties=[mdsuite_properties.unwrapped_positions],
                                      output_property=PropertyInfo("test_single", 2)))
    got = exp.load_matrix(species=[sp], property_name="test_single")
    assert got.shape == (10, 100, 2) and np.all(got == 516)
    exp.cls_transformation_run(multi(input_properties=[mdsuite_properties.unwrapped_positions],
                                     output_property=PropertyInfo("test_multi", 2)))
    got = exp.load_matrix(species=["Observables"], property_name="test_multi")
    assert got.shape == (1, 100, 2) and np.all(got == 516)


def test_data_from_species_and_experiment(tmp_path, cuda):
    """:177-210 -- positions per configuration, charge from the species, box from the
    experiment, all handed to transform_batch."""
    import torch
    from lammps_analysis_b200.transformations import (PropertyInfo, SingleSpeciesTrafo,
                                                      mdsuite_properties)

    single, multi = _test_trafos()
    project, exp = _exp(tmp_path)
    sp, _ = _load_pos(exp, 5)
    exp.box_array = [1.1, 2.2, 3.3]
    exp.species[sp].charge = 1.23435
    props = [mdsuite_properties.unwrapped_positions, mdsuite_properties.charge,
             mdsuite_properties.box_length]
    for cls, name in ((single, "test1"), (multi, "test2")):
        exp.cls_transformation_run(cls(input_properties=props,
                                       output_property=PropertyInfo(name, 2)))
    seen = {}

    class Probe(SingleSpeciesTrafo):
        def transform_batch(self, batch, carryover=None):
            seen.update(batch)
            return torch.zeros(1, batch["Unwrapped_Positions"].shape[1], 1, device="cuda")

    exp.cls_transformation_run(Probe(input_properties=props,
                                     output_property=PropertyInfo("probe", 1)))
    assert seen["Unwrapped_Positions"].shape == (10, 100, 3)
    assert np.allclose(seen["Charge"], [1.23435]) and np.allclose(seen["Box_Array"], [1.1, 2.2, 3.3])


def test_transformation_on_new_data_(tmp_path, cuda):
    """:213-237 -- after new data the transformation still works ... and the output covers the
    appended frames (the carry-over check of the test trafos holds across the extension)."""
    from lammps_analysis_b200.transformations import PropertyInfo, mdsuite_properties

    single, multi = _test_trafos()
    for cls, name, path in ((single, "test1", "test_species/test_prop"),
                            (multi, "test2", "Observables/test_prop")):
        project, exp = _exp(tmp_path, name=name, timestep=12345)
        trafo = cls(input_properties=[mdsuite_properties.unwrapped_positions],
                    output_property=PropertyInfo("test_prop", 2))
        _load_pos(exp, 6, name="first")
        exp.cls_transformation_run(trafo)
        assert exp.store.shape(path)[1] == 100
        _load_pos(exp, 7, name="second")
        exp.cls_transformation_run(trafo)
        assert exp.store.shape(path)[1] == 200
        assert np.all(exp.store.host(path) == 516)
        exp.cls_transformation_run(trafo)           # complete: skipped
        assert exp.store.shape(path)[1] == 200


def test_extending_the_hot_path_transformations_matches_one_run(tmp_path, cuda):
    """transformations.py:300-311 done as documented: after Experiment.add_data appends frames,
    CoordinateUnwrapper and IonicCurrent compute only the new frames (offset) and the result
    equals a single run over the whole trajectory -- the unwrap carry (last position, image
    count) is rebuilt from the last stored frame.  A calculator that depends on the extended
    property sees the grown experiment."""
    from lammps_analysis_b200.config import config
    from lammps_analysis_b200.file_io import ScriptInput
    from lammps_analysis_b200.synthetic import nacl_trajectory
    from oracle import transformations as ot

    config.planner_memory_bytes = 60e9
    data, box = nacl_trajectory(216, 400, 7.0, seed=44, sigma_step=0.5)   # many box crossings
    first = {s: {p: a[:, :250] for p, a in d.items()} for s, d in data.items()}
    second = {s: {p: a[:, 250:] for p, a in d.items()} for s, d in data.items()}
    for persist in (True, False):
        project, exp = _exp(tmp_path, name=f"ext{int(persist)}", timestep=0.002, persist=persist)
        exp.add_data(ScriptInput(first, box, atom_major=True, name="first"))
        exp.species["Na"].charge, exp.species["Cl"].charge = 1.0, -1.0
        exp.run.CoordinateUnwrapper()
        exp.run.IonicCurrent()
        short = exp.run.EinsteinDiffusionCoefficients(data_range=50, plot=False)
        exp.add_data(ScriptInput(second, box, atom_major=True, name="second"))
        assert exp.number_of_configurations == 400
        exp.run.CoordinateUnwrapper()
        exp.run.IonicCurrent()
        for sp in ("Na", "Cl"):
            want = ot.run_unwrap(data[sp]["Positions"], box, batch_size=400)
            assert np.array_equal(exp.store.host(f"{sp}/Unwrapped_Positions"), want)
        J = ot.run_ionic_current({s: data[s]["Velocities"] for s in data}, {"Na": 1.0, "Cl": -1.0})
        got = exp.store.host("Observables/Ionic_Current")
        assert got.shape == (1, 400, 3)
        np.testing.assert_allclose(got, J, rtol=2e-7, atol=1e-6)
        # the dependency check of a calculator extends a stale dataset on its own
        project2, exp2 = _exp(tmp_path, name=f"dep{int(persist)}", timestep=0.002,
                              persist=persist)
        exp2.add_data(ScriptInput(first, box, atom_major=True, name="first"))
        exp2.run.EinsteinDiffusionCoefficients(data_range=50, plot=False)
        exp2.add_data(ScriptInput(second, box, atom_major=True, name="second"))
        long = exp2.run.EinsteinDiffusionCoefficients(data_range=50, plot=False)
        assert exp2.store.shape("Na/Unwrapped_Positions")[1] == 400
        full = exp.run.EinsteinDiffusionCoefficients(data_range=50, plot=False)
        np.testing.assert_allclose(long["Na"]["msd"], full["Na"]["msd"], rtol=1e-12)
        assert long["Na"]["msd"] != short["Na"]["msd"]


def test_row_block_pipeline_matches_single_block(tmp_path, cuda, monkeypatch):
    """In-memory (page-locked) stores run a transformation in blocks of rows pipelined over
    upload / compute / write-back streams; many small blocks, a ragged last block and the
    resident result must equal the one-block run and the oracle."""
    from lammps_analysis_b200.file_io import ScriptInput
    from lammps_analysis_b200.synthetic import nacl_trajectory
    from lammps_analysis_b200.transformations import CoordinateUnwrapper
    from oracle import transformations as ot

    data, box = nacl_trajectory(343, 150, 9.0, seed=45, sigma_step=0.6)
    project, exp = _exp(tmp_path, name="blocks", timestep=0.002, persist=False)
    exp.add_data(ScriptInput(data, box, atom_major=True))
    # 150 frames x 24 B per row: 7 rows per block -> 25 blocks for the 172 / 171 atom species
    monkeypatch.setattr(CoordinateUnwrapper, "block_bytes", 7 * 150 * 24)
    exp.run.CoordinateUnwrapper()
    for sp in ("Na", "Cl"):
        want = ot.run_unwrap(data[sp]["Positions"], box, batch_size=150)
        assert np.array_equal(exp.store.host(f"{sp}/Unwrapped_Positions"), want)
        assert exp.store.is_resident(f"{sp}/Unwrapped_Positions")
        assert np.array_equal(exp.store.device(f"{sp}/Unwrapped_Positions").cpu().numpy(), want)
    res = exp.run.EinsteinDiffusionCoefficients(data_range=40, plot=False)
    assert np.isfinite(res["Na"]["diffusion_coefficient"])


def test_streamed_calculators_match_one_block(tmp_path, cuda, monkeypatch):
    """Einstein / Green-Kubo consume their datasets in row blocks as the blocks arrive (unwrap
    pipeline events, block-wise velocity upload, merged lag-product launches, read-back through
    mdk_store_mapped): many small ragged blocks must give what one block gives, and the oracle."""
    from lammps_analysis_b200.config import config
    from lammps_analysis_b200.file_io import ScriptInput
    from lammps_analysis_b200.synthetic import nacl_trajectory
    from lammps_analysis_b200.transformations import CoordinateUnwrapper
    from oracle import dynamics as od
    from oracle import transformations as ot

    data, box = nacl_trajectory(216, 160, 9.0, seed=46, sigma_step=0.5)
    results = []
    for tag, rows in (("one", 10_000), ("many", 9)):
        project, exp = _exp(tmp_path / tag, name="stream", timestep=0.002, persist=False)
        exp.add_data(ScriptInput(data, box, atom_major=True))
        monkeypatch.setattr(CoordinateUnwrapper, "block_bytes", rows * 160 * 24)
        monkeypatch.setattr(config, "upload_block_bytes", rows * 160 * 12)
        ein = exp.run.EinsteinDiffusionCoefficients(data_range=50, plot=False)
        exp.store.invalidate()          # velocities (and positions) start on the host again
        gk = exp.run.GreenKuboDiffusionCoefficients(data_range=50, plot=False)
        results.append((ein.data_dict, gk.data_dict))
        _, blocks = exp.store.device_blocks("Na/Velocities")
        assert len(blocks) == (12 if tag == "many" else 1)      # 108 rows in blocks of 9
    (e1, g1), (e2, g2) = results
    for sp in ("Na", "Cl"):
        np.testing.assert_allclose(e2[sp]["msd"], e1[sp]["msd"], rtol=1e-6, atol=1e-300)
        # the lag products run in fp32 over up to 256 atoms before they are folded into fp64:
        # another atom grouping moves them at the 1e-7 level
        np.testing.assert_allclose(g2[sp]["acf"], g1[sp]["acf"], rtol=1e-6, atol=1e-18)
        np.testing.assert_allclose(g2[sp]["integral_uncertainty"], g1[sp]["integral_uncertainty"],
                                   rtol=1e-4, atol=1e-18)
        unw = ot.run_unwrap(data[sp]["Positions"], box, batch_size=160)
        plan = dict(batch_size=160, n_batches=1, remainder=0, minibatch=False)
        ref, count = od.einstein_msd(unw, plan, 50, 1, np.arange(50))
        got = np.array(e2[sp]["msd"]) / (exp.units.length ** 2)
        np.testing.assert_allclose(got, ref / count, rtol=1e-5, atol=1e-12)


def test_read_back_equals_cpu_copy(cuda):
    import torch
    from lammps_analysis_b200 import kernels as K

    gen = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(1500, 37, device="cuda", dtype=torch.float64, generator=gen)
    b = torch.randn(5, device="cuda", dtype=torch.float64, generator=gen)      # 40 bytes: tail path
    c = torch.arange(1001, device="cuda", dtype=torch.int64)
    ga, gb, gc = K.read_back(a, b, c)
    assert np.array_equal(ga, a.cpu().numpy()) and np.array_equal(gb, b.cpu().numpy())
    assert np.array_equal(gc, c.cpu().numpy())
