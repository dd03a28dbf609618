"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN PYTHON SOURCE for the hot path.

`import mdsuite` is impossible here (tensorflow, tensorflow_probability, h5py, sqlalchemy, ... are
not installed and there is no network), so this script

  * parses the reference files under /root/reference with `ast`, pulls out the functions /
    methods / classes on the hot path (nothing is copied into this repository), and
  * executes them against a small NumPy stand-in for the TensorFlow ops they call (`TFShim`
    below: gather, boolean_mask, where, band_part, rint/round, norm, histogram_fixed_width,
    cumsum, diff, ...).

What these vectors pin: the reference's *Python-level* logic -- index construction, the strict
species masks (Q1), minibatch bookkeeping, the unwrap carry-over, the planner and the window
generator, the Einstein fit -- executed line by line from upstream source.  What they cannot
pin: TensorFlow's own numerics (the shim restates them; see oracle/__init__.py), and
tfp.stats.auto_correlation (not executed here).

Run (in the authoring container, where /root/reference exists):
    python tests/golden/make_reference_goldens.py
Writes tests/golden/reference_run.json.  The GPU box never needs /root/reference.
"""
import ast
import itertools
import json
import logging
import os
import types
from timeit import default_timer as timer

import numpy as np
from scipy.interpolate import UnivariateSpline
from scipy.optimize import curve_fit
from scipy.signal import savgol_filter

REF = "/root/reference/mdsuite"
HERE = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------------------------------
# NumPy stand-in for the TensorFlow ops used by the extracted code
# ------------------------------------------------------------------------------------------
class _Math:
    @staticmethod
    def rint(x):
        return np.rint(x)

    round = rint  # tf.math.round is round-half-to-even

    @staticmethod
    def cumsum(x, axis=0):
        return np.cumsum(x, axis=axis)

    @staticmethod
    def squared_difference(a, b):
        d = np.asarray(a) - np.asarray(b)
        return d * d


class _Linalg:
    @staticmethod
    def band_part(mat, num_lower, num_upper):
        m, n = mat.shape
        rows, cols = np.arange(m)[:, None], np.arange(n)[None, :]
        keep = np.ones_like(mat, dtype=bool)
        if num_lower >= 0:
            keep &= (rows - cols) <= num_lower
        if num_upper >= 0:
            keep &= (cols - rows) <= num_upper
        return mat & keep

    @staticmethod
    def norm(x, axis=-1):
        sq = x * x  # products rounded in the array dtype
        return np.sqrt((sq[..., 0] + sq[..., 1]) + sq[..., 2])


class _ExpNumpy:
    @staticmethod
    def diff(x, axis=-1):
        return np.diff(x, axis=axis)


class TFShim:
    float32, float64, int32, int16, bool = np.float32, np.float64, np.int32, np.int16, np.bool_
    Tensor = np.ndarray
    math, linalg = _Math, _Linalg
    experimental = types.SimpleNamespace(numpy=_ExpNumpy)

    @staticmethod
    def function(*a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return lambda f: f

    @staticmethod
    def shape(x):
        return np.shape(x)

    @staticmethod
    def cast(x, dtype):
        return np.asarray(x).astype(dtype)

    @staticmethod
    def ones(shape, dtype=np.float32):
        return np.ones(shape, dtype=dtype)

    @staticmethod
    def zeros(shape, dtype=np.float32):
        return np.zeros(shape, dtype=dtype)

    zeros_like = staticmethod(np.zeros_like)

    @staticmethod
    def constant(x, dtype=None):
        return np.asarray(x, dtype=dtype)

    @staticmethod
    def stack(xs, axis=0):
        return np.stack([np.asarray(x) for x in xs], axis=axis)

    @staticmethod
    def concat(xs, axis=0):
        return np.concatenate(xs, axis=axis)

    @staticmethod
    def expand_dims(x, axis):
        return np.expand_dims(x, axis)

    @staticmethod
    def transpose(x):
        return np.transpose(x)

    @staticmethod
    def where(cond):
        return np.argwhere(cond)

    @staticmethod
    def gather(x, idx, axis=0):
        return np.take(x, idx, axis=axis)

    @staticmethod
    def boolean_mask(x, mask, axis=0):
        mask = np.asarray(mask)
        if mask.ndim == 1:
            return np.compress(mask, x, axis=axis)
        return np.asarray(x)[mask]  # mask of the tensor's rank: selected elements, flattened

    @staticmethod
    def less(a, b):
        return np.less(a, b)

    @staticmethod
    def reduce_sum(x, axis=None):
        return np.sum(x, axis=axis)

    @staticmethod
    def add_n(xs):
        out = xs[0]
        for x in xs[1:]:
            out = out + x
        return out

    @staticmethod
    def histogram_fixed_width(values, value_range, nbins):
        # tensorflow/core/kernels/histogram_op.cc, CPU functor
        values = np.asarray(values, dtype=np.float32).ravel()
        lo, hi = np.float32(value_range[0]), np.float32(value_range[1])
        step = float(np.float32(hi - lo)) / float(nbins)
        shifted = (np.maximum(values, lo) - lo).astype(np.float64)
        idx = np.minimum(shifted / step, float(int(nbins) - 1)).astype(np.int32)
        return np.bincount(idx, minlength=int(nbins)).astype(np.int32)

    @staticmethod
    def device(_):
        return None


tf = TFShim


# ------------------------------------------------------------------------------------------
# source extraction
# ------------------------------------------------------------------------------------------
def extract(path, names, ns, strip_decorators=True):
    """Compile the top-level functions / classes / methods called `names` of a reference file
    into namespace `ns` (methods become plain functions taking `self`)."""
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    found = set()

    def take(node):
        if strip_decorators and isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            node.decorator_list = []
        mod = ast.Module(body=[node], type_ignores=[])
        exec(compile(mod, os.path.join(REF, path), "exec"), ns)
        found.add(node.name)

    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            take(node)
        elif isinstance(node, ast.ClassDef):
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name in names:
                    take(sub)
    missing = set(names) - found
    assert not missing, f"{path}: {missing} not found"


def jsonable(x):
    if isinstance(x, dict):
        return {str(k): jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.floating,)):
        return float(x)
    return x


def main():
    log = logging.getLogger("golden")
    out = {"generator": "tests/golden/make_reference_goldens.py",
           "note": "reference Python source executed under a NumPy TensorFlow shim"}
    rng = np.random.default_rng(20261018)

    # ---- RDF: utils/linalg.py + radial_distribution_function.py ------------------------------
    ns = {"tf": tf, "np": np, "itertools": itertools, "log": log, "timer": timer}
    extract("utils/linalg.py", ["apply_minimum_image", "get_partial_triu_indices",
                                "apply_system_cutoff"], ns)
    extract("calculators/radial_distribution_function.py",
            ["bin_minibatch", "get_dij", "compute_species_values", "run_minibatch_loop",
             "_get_species_names", "combine_dictionaries"], ns)
    species = ["Na", "Cl"]
    n_part = {"Na": 23, "Cl": 17}
    box = [6.0, 6.5, 7.0]
    n_frames, cutoff, nbins, rdf_minibatch = 5, 2.9, 290, 7
    pos = {s: (rng.random((n_part[s], n_frames, 3)) * np.array(box)).astype(np.float32)
           for s in species}
    self_ = types.SimpleNamespace()
    self_.args = types.SimpleNamespace(number_of_bins=nbins, cutoff=cutoff, species=species)
    self_.dtype = tf.float32
    self_.index_list = [0, 1]
    self_.particles_list = [n_part[s] for s in species]
    self_.bin_range = [0, cutoff]
    self_.experiment = types.SimpleNamespace(box_array=box)
    self_._get_species_names = lambda t: ns["_get_species_names"](self_, t)
    self_.key_list = [self_._get_species_names(x)
                      for x in itertools.combinations_with_replacement([0, 1], r=2)]
    self_.bin_minibatch = ns["bin_minibatch"]
    self_.get_dij = ns["get_dij"]
    self_.compute_species_values = lambda i, s, d: ns["compute_species_values"](self_, i, s, d)
    total = {k: np.zeros(nbins, dtype=np.int64) for k in self_.key_list}
    # batch loop of run_calculator (:846-885): one frame per batch here
    for f in range(n_frames):
        positions_tensor = np.concatenate([pos[s][:, f:f + 1] for s in species], axis=0)
        n_atoms = positions_tensor.shape[0]
        minibatch_start, stop = 0, 0
        rdf = {k: np.zeros(nbins, dtype=np.int32) for k in self_.key_list}
        for lo in range(0, n_atoms, rdf_minibatch):
            atoms = positions_tensor[lo:lo + rdf_minibatch]
            mb, minibatch_start, stop = ns["run_minibatch_loop"](
                self_, atoms, stop, n_atoms, minibatch_start, positions_tensor)
            for k in rdf:
                rdf[k] = rdf[k] + mb[k]
        for k in total:
            total[k] += rdf[k]
    out["rdf"] = {"species": species, "box": box, "cutoff": cutoff, "nbins": nbins,
                  "positions": jsonable(pos), "counts": jsonable(total)}

    # ---- transformations ---------------------------------------------------------------------
    props = types.SimpleNamespace(
        positions=types.SimpleNamespace(name="Positions"),
        box_length=types.SimpleNamespace(name="Box_Length"),
        box_images=types.SimpleNamespace(name="Box_Images"),
        velocities=types.SimpleNamespace(name="Velocities"),
        charge=types.SimpleNamespace(name="Charge"),
        unwrapped_positions=types.SimpleNamespace(name="Unwrapped_Positions"))
    import typing
    ns_t = {"tf": tf, "np": np, "mdsuite_properties": props, "typing": typing}
    src_names = {"unwrap": "transformations/unwrap_coordinates.py",
                 "indices": "transformations/unwrap_via_indices.py",
                 "ionic": "transformations/ionic_current.py",
                 "dipole": "transformations/translational_dipole_moment.py"}
    fns = {}
    for key, path in src_names.items():
        local = dict(ns_t)
        extract(path, ["transform_batch"], local)
        fns[key] = local["transform_batch"]
    A, T = 6, 40
    box_l = np.array([3.1, 2.9, 3.3])
    walk = np.cumsum(rng.normal(0, 0.9, size=(A, T, 3)), axis=1) + rng.random((A, 1, 3)) * box_l
    wrapped = np.mod(walk, box_l).astype(np.float32)
    carry, pieces = None, []
    for lo, hi in ((0, 13), (13, 30), (30, 40)):      # three batches with carry-over
        res, carry = fns["unwrap"](None, {"Positions": wrapped[:, lo:hi].astype(np.float64),
                                          "Box_Length": box_l.reshape(1, 1, 3)}, carry)
        pieces.append(np.asarray(res))
    out["unwrap"] = {"box": box_l.tolist(), "positions": wrapped.tolist(),
                     "batches": [[0, 13], [13, 30], [30, 40]],
                     "unwrapped": np.concatenate(pieces, axis=1).tolist()}
    img = rng.integers(-4, 5, size=(A, T, 3)).astype(np.float64)
    res = fns["indices"](None, {"Positions": wrapped.astype(np.float64), "Box_Images": img,
                                "Box_Length": box_l.reshape(1, 1, 3)})
    out["unwrap_indices"] = {"images": img.tolist(), "unwrapped": np.asarray(res).tolist()}
    vel = {s: rng.normal(size=(n, T, 3)).astype(np.float32) for s, n in (("Na", 5), ("Cl", 4))}
    q = {"Na": 1.0, "Cl": -1.0}
    res = fns["ionic"](None, {s: {"Velocities": vel[s].astype(np.float64),
                                  "Charge": np.array([[[q[s]]]])} for s in vel})
    out["ionic_current"] = {"velocities": jsonable(vel), "charge": q,
                            "current": np.asarray(res).tolist()}
    res = fns["dipole"](None, {s: {"Unwrapped_Positions": vel[s].astype(np.float64),
                                   "Charge": np.array([[[q[s]]]])} for s in vel})
    out["dipole_moment"] = {"moment": np.asarray(res).tolist()}

    # ---- planner: memory_manager.py + scale_functions.py ------------------------------------
    ns_p = {"np": np, "tf": tf, "log": log, "Tuple": typing.Tuple,
            "Database": object, "gpu_available": lambda: False,
            "config": types.SimpleNamespace(memory_fraction=0.5),
            "get_machine_properties": lambda: {"memory": 0.0, "gpu": {}}}
    extract("utils/scale_functions.py", ["linear_scale_function", "linearithmic_scale_function",
                                         "quadratic_scale_function",
                                         "polynomial_scale_function"], ns_p)
    extract("memory_management/memory_manager.py", ["MemoryManager"], ns_p)

    class FakeDB:
        def __init__(self, rows, cols, nbytes):
            self.t = (rows, cols, nbytes)

        def get_data_size(self, item):
            return self.t

    specs = [{"linear": {"scale_factor": 150}}, {"linear": {"scale_factor": 5}},
             {"quadratic": {"inner_scale_factor": 5, "outer_scale_factor": 10}},
             {"linear": {"scale_factor": 2}}]
    cases = []
    for _ in range(120):
        rows, cols = int(rng.integers(1, 4000)), int(rng.integers(2, 12000))
        nbytes = rows * cols * 12
        mem = float(10 ** rng.uniform(2, 11))
        N, ct = int(rng.integers(1, 600)), int(rng.integers(1, 6))
        sf = specs[int(rng.integers(0, len(specs)))]
        mm = ns_p["MemoryManager"](data_path=["x"], database=FakeDB(rows, cols, nbytes),
                                   scale_function=sf, gpu=False)
        mm.machine_properties["memory"] = mem
        bs, nb, rem = mm.get_batch_size()
        loops, minibatch = mm.get_ensemble_loop(N, ct)
        cases.append({"rows": rows, "cols": cols, "nbytes": nbytes, "memory": mem, "data_range": N,
                      "correlation_time": ct, "scale_function": sf,
                      "get_batch_size": [bs, nb, rem], "ensemble_loop": loops,
                      "minibatch": bool(minibatch), "batch_size": mm.batch_size,
                      "n_batches": mm.n_batches, "remainder": mm.remainder,
                      "atom_batch_size": mm.atom_batch_size, "n_atom_batches": mm.n_atom_batches,
                      "atom_remainder": mm.atom_remainder})
    out["planner"] = cases

    # ---- data_manager.py: window generator and batch slices ----------------------------------
    loads = []

    class RecordingDB:
        def __init__(self, path):
            self.path = path

        def load_data(self, path_list=None, select_slice=None, dictionary=False, scaling=None,
                      d_size=None):
            loads.append(select_slice)
            return {}

    ns_d = {"np": np, "tf": tf, "log": log, "tqdm": lambda x, **k: x, "Database": RecordingDB}
    extract("database/data_manager.py", ["DataManager"], ns_d)
    DM = ns_d["DataManager"]
    windows = []
    for data_size, N, ct in ((50, 10, 5), (50, 50, 1), (37, 10, 3), (8, 10, 1), (500, 100, 7)):
        dm = DM(database=types.SimpleNamespace(path="p"), data_path=["x"], data_range=N,
                correlation_time=ct, ensemble_loop=1)
        gen, args = dm.ensemble_generator(glob_data={b"data_size": data_size,
                                                     b"x": np.arange(data_size)[None, :, None]})
        got = [[int(o[b"x"][0, 0, 0]), int(o[b"x"][0, -1, 0]) + 1] for o in gen(*args)]
        windows.append({"data_size": data_size, "data_range": N, "correlation_time": ct,
                        "windows": got})
    out["ensemble_windows"] = windows
    slices = []
    for kw in (dict(batch_size=30, n_batches=3, remainder=7),
               dict(batch_size=40, n_batches=2, remainder=0, minibatch=True, atom_batch_size=5,
                    n_atom_batches=2, atom_remainder=0)):
        loads.clear()
        dm = DM(database=types.SimpleNamespace(path="p"), data_path=["x"], data_range=1, **kw)
        gen, args = dm.batch_generator(remainder=kw.get("remainder", 0) > 0)
        for _ in gen(*args):
            pass
        rec = []
        for s in loads:
            rec.append([[x.start, x.stop] if isinstance(x, slice) else x for x in
                        (s if isinstance(s, tuple) else (s,))])
        slices.append({"plan": kw, "slices": jsonable(rec)})
    out["batch_slices"] = slices

    # ---- fits and searches ---------------------------------------------------------------------
    ns_f = {"np": np, "UnivariateSpline": UnivariateSpline, "curve_fit": curve_fit,
            "Tuple": typing.Tuple, "Union": typing.Union, "Iterable": typing.Iterable,
            "Any": typing.Any, "ndarray": np.ndarray}
    extract("utils/calculator_helper_methods.py", ["fit_einstein_curve"], ns_f)
    x = np.arange(120) * 0.02
    y = 6 * 1.3 * x + 0.4 * np.sqrt(x) + rng.normal(0, 0.01, 120)
    popt, pcov, grads, gerrs = ns_f["fit_einstein_curve"](x_data=x, y_data=y, fit_max_index=119)
    out["fit_einstein_curve"] = {"x": x.tolist(), "y": y.tolist(), "fit_max_index": 119,
                                 "popt": np.asarray(popt).tolist(),
                                 "pcov": np.asarray(pcov).tolist(),
                                 "gradients": np.asarray(grads).tolist()}
    ns_g = {"np": np, "savgol_filter": savgol_filter, "golden_ratio": 1.618033988749895,
            "typing": typing, "Callable": typing.Callable}
    extract("utils/meta_functions.py", ["closest_point", "golden_section_search",
                                        "apply_savgol_filter"], ns_g)
    r = np.linspace(0.05, 1.5, 400)
    g = 1 + 2 * np.exp(-3 * r) * np.cos(14 * r)
    lo_hi = ns_g["golden_section_search"]([r, g], r[200], r[60])
    out["golden_section_search"] = {"r": r.tolist(), "g": g.tolist(), "a": float(r[200]),
                                    "b": float(r[60]), "result": [float(lo_hi[0]), float(lo_hi[1])]}

    # ---- Einstein ensemble_operation (:168-190) -----------------------------------------------
    ns_e = {"tf": tf, "np": np}
    extract("calculators/einstein_diffusion_coefficients.py", ["ensemble_operation"], ns_e)
    ens = np.cumsum(rng.normal(size=(5, 12, 3)), axis=1)
    fake = types.SimpleNamespace(args=types.SimpleNamespace(tau_values=np.arange(12)), count=0)
    msd = ns_e["ensemble_operation"](fake, ens)
    out["einstein_ensemble_operation"] = {"ensemble": ens.tolist(), "msd": np.asarray(msd).tolist(),
                                          "count": int(fake.count)}

    # ---- RDF normalisation (:299-393, 719-826) and CN / PMF / KBI post-processing ------------
    from scipy.integrate import cumulative_trapezoid
    from scipy.signal import find_peaks

    ns_n = {"np": np, "Union": typing.Union, "log": log}
    extract("utils/meta_functions.py", ["split_array"], ns_n)
    extract("calculators/radial_distribution_function.py",
            ["ideal_correction", "_calculate_prefactor", "_ang_to_nm"], ns_n)
    norm_cases = []
    for box0, cut, nb in ((20.0, 9.9, 99), (10.0, 6.5, 65), (8.0, 6.9, 69)):  # beyond L/2 too
        fake = types.SimpleNamespace(
            args=types.SimpleNamespace(cutoff=cut, number_of_bins=nb, molecules=False,
                                       atom_selection=np.s_[:], number_of_configurations=7),
            experiment=types.SimpleNamespace(
                box_array=[box0, box0 * 1.1, box0 * 0.9], volume=box0 * box0 * 1.1 * box0 * 0.9,
                species={"Na": types.SimpleNamespace(n_particles=23),
                         "Cl": types.SimpleNamespace(n_particles=17)},
                units=types.SimpleNamespace(length=1e-10)))
        with np.errstate(all="ignore"):
            fake.ideal_correction = ns_n["ideal_correction"](fake)
            pref = {k: ns_n["_calculate_prefactor"](fake, k) for k in ("Na_Na", "Na_Cl", "Cl_Cl")}
            x = ns_n["_ang_to_nm"](fake, np.linspace(0.0, cut, nb))
        norm_cases.append({"box": fake.experiment.box_array, "cutoff": cut, "nbins": nb,
                           "n_configs": 7, "n_particles": {"Na": 23, "Cl": 17},
                           "ideal_correction": np.nan_to_num(fake.ideal_correction, nan=-1.0,
                                                             posinf=-2.0).tolist(),
                           "prefactor": {k: np.nan_to_num(v, nan=-1.0, posinf=-2.0).tolist()
                                         for k, v in pref.items()},
                           "x": x.tolist()})
    out["rdf_normalisation"] = norm_cases

    ns_c = {"np": np, "cumulative_trapezoid": cumulative_trapezoid, "find_peaks": find_peaks,
            "log": log, "golden_section_search": ns_g["golden_section_search"],
            "apply_savgol_filter": ns_g["apply_savgol_filter"],
            "CannotPerformThisAnalysis": ValueError}
    extract("calculators/coordination_number_calculation.py",
            ["_integrate_rdf", "_get_rdf_peaks", "_find_minima", "_get_coordination_numbers"],
            ns_c)
    rr = np.linspace(0.0, 1.2, 500)
    gg = np.where(rr < 0.18, 0.0, 1 + 2.2 * np.exp(-4 * (rr - 0.18)) * np.cos(18 * (rr - 0.25)))
    radii, rdf = rr[1:], gg[1:]
    fake = types.SimpleNamespace(args=types.SimpleNamespace(savgol_order=2,
                                                            savgol_window_length=17,
                                                            number_of_shells=2))
    fake._get_rdf_peaks = lambda r: ns_c["_get_rdf_peaks"](fake, r)
    fake._find_minima = lambda a, b: ns_c["_find_minima"](fake, a, b)
    density = 31.7
    integral = ns_c["_integrate_rdf"](radii, rdf, density)
    cn = ns_c["_get_coordination_numbers"](fake, integral, radii, rdf)
    out["coordination_numbers"] = {"x": rr.tolist(), "y": gg.tolist(), "density": density,
                                   "cn": integral.tolist(), "values": jsonable(cn)}

    with open(os.path.join(HERE, "reference_run.json"), "w") as fh:
        json.dump(jsonable(out), fh)
    print("written", os.path.join(HERE, "reference_run.json"))


if __name__ == "__main__":
    main()
