"""Golden vectors for the flux transformations and the flux Green-Kubo / Einstein-Helfand
calculators, produced by EXECUTING THE REFERENCE'S OWN PYTHON SOURCE: the
`transform_batch` bodies of momentum_flux.py, thermal_flux.py, integrated_heat_current.py and
the `ensemble_operation` of einstein_helfand_thermal_conductivity.py are pulled out of the
upstream files with `ast` (nothing is copied into this repository) and run against the NumPy
TensorFlow stand-in of make_reference_goldens.py.  tfp.stats.auto_correlation is not executed
(not installed); the Green-Kubo flux calculators stay pinned to its restated semantics.

Run (in the authoring container, where /root/reference exists):
    python tests/golden/make_reference_flux_goldens.py
Writes tests/golden/reference_flux.json.  The GPU box never needs /root/reference.
"""
import json
import os
import sys
import types
import typing

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_reference_goldens import extract, jsonable, tf  # noqa: E402


def main():
    rng = np.random.default_rng(2026101808)
    props = types.SimpleNamespace(
        stress=types.SimpleNamespace(name="Stress"),
        velocities=types.SimpleNamespace(name="Velocities"),
        kinetic_energy=types.SimpleNamespace(name="Kinetic_Energy"),
        potential_energy=types.SimpleNamespace(name="Potential_Energy"),
        unwrapped_positions=types.SimpleNamespace(name="Unwrapped_Positions"),
        momentum_flux=types.SimpleNamespace(name="Momentum_Flux"),
        thermal_flux=types.SimpleNamespace(name="Thermal_Flux"),
        integrated_heat_current=types.SimpleNamespace(name="Integrated_Heat_Current"))
    ns_t = {"tf": tf, "np": np, "mdsuite_properties": props, "typing": typing}
    fns = {}
    for key, path in (("momentum", "transformations/momentum_flux.py"),
                      ("thermal", "transformations/thermal_flux.py"),
                      ("heat", "transformations/integrated_heat_current.py")):
        local = dict(ns_t)
        extract(path, ["transform_batch"], local)
        fns[key] = local["transform_batch"]
    T = 9
    data = {}
    for sp, n in (("Na", 5), ("Cl", 3)):
        data[sp] = {
            "Stress": rng.normal(size=(n, T, 6)).astype(np.float32),
            "Velocities": rng.normal(size=(n, T, 3)).astype(np.float32),
            "Kinetic_Energy": rng.uniform(0.5, 1.5, size=(n, T, 1)).astype(np.float32),
            "Potential_Energy": rng.normal(-3, 0.4, size=(n, T, 1)).astype(np.float32),
            "Unwrapped_Positions": (rng.normal(size=(n, T, 3)) * 7).astype(np.float32)}
    batch = {sp: {k: v.astype(np.float64) for k, v in d.items()} for sp, d in data.items()}
    out = {"generator": "tests/golden/make_reference_flux_goldens.py",
           "inputs": jsonable(data),
           "momentum_flux": np.asarray(fns["momentum"](None, batch)).tolist(),
           "thermal_flux": np.asarray(fns["thermal"](None, batch)).tolist(),
           "integrated_heat_current": np.asarray(fns["heat"](None, batch)).tolist()}

    # EinsteinHelfandThermalConductivity.ensemble_operation (:187-203) on a (N, 3) window
    ns_e = {"tf": tf, "np": np}
    extract("calculators/einstein_helfand_thermal_conductivity.py", ["ensemble_operation"], ns_e)
    window = np.cumsum(rng.normal(size=(14, 3)), axis=0)
    fake = types.SimpleNamespace(prefactor=0.37, msd_array=np.zeros(14))
    ns_e["ensemble_operation"](fake, window)
    out["eh_thermal_ensemble_operation"] = {"window": window.tolist(), "prefactor": 0.37,
                                            "msd": np.asarray(fake.msd_array).tolist()}
    with open(os.path.join(HERE, "reference_flux.json"), "w") as fh:
        json.dump(jsonable(out), fh)
    print("written", os.path.join(HERE, "reference_flux.json"))


if __name__ == "__main__":
    main()
