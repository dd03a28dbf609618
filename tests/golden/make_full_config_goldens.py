#!/usr/bin/env python
"""Oracle results at BASELINE config C2's FULL size (1,000-atom NaCl, 5,000 frames,
data_range 500, correlation_time 1): the per-window oracle needs ~10 minutes of CPU per
species and calculator, too slow for the test run, so its output is committed as a fixture
(tests/golden/c2_full.json) and the GPU path is compared against it in
tests/test_gpu_fullsize.py.  Inputs are regenerated in the test from the same seeded generator
(lammps_analysis_b200.synthetic.nacl_trajectory(1000, 5000, 32.0, seed=2, sigma_step=0.3)).

    python tests/golden/make_full_config_goldens.py        # writes c2_full.json (4 processes)
"""
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

N_ATOMS, N_FRAMES, BOX, SEED, SIGMA = 1000, 5000, 32.0, 2, 0.3
DATA_RANGE, MEM = 500, 60e9
TIME_STEP, SAMPLE_RATE = 0.002, 10
U_LENGTH, U_TIME = 1e-10, 1e-12      # METAL


def _plan(A, T):
    from oracle.planner import ArrayDatabase, plan_trajectory_calculator

    class _S:
        shape = (A, T, 3)

    return plan_trajectory_calculator(ArrayDatabase({"x": _S()}), ["x"], DATA_RANGE, 1,
                                      {"linear": {"scale_factor": 150}}, MEM)


def _job(args):
    kind, sp = args
    from lammps_analysis_b200.synthetic import nacl_trajectory
    from oracle import dynamics as od
    from oracle import transformations as ot

    data, box = nacl_trajectory(N_ATOMS, N_FRAMES, BOX, seed=SEED, sigma_step=SIGMA)
    A = data[sp]["Positions"].shape[0]
    plan = _plan(A, N_FRAMES)
    tau, _, _, times = od.handle_tau_values(np.s_[:], DATA_RANGE, TIME_STEP, SAMPLE_RATE)
    if kind == "einstein":
        unw = ot.run_unwrap(data[sp]["Positions"], box, batch_size=N_FRAMES)
        msd_sum, count = od.einstein_msd(unw, plan, DATA_RANGE, 1, tau)
        ref = od.einstein_finish(msd_sum, count, times, U_LENGTH, U_TIME, DATA_RANGE - 1)
        out = {"count": int(count), "msd": list(map(float, ref["msd"])),
               "time": list(map(float, ref["time"])),
               "diffusion_coefficient": float(ref["diffusion_coefficient"]),
               "uncertainty": float(ref["uncertainty"]),
               "unwrapped_checksum": float(np.asarray(unw, dtype=np.float64).sum())}
    else:
        time = times * U_TIME
        acf_sum, count, sigmas = od.gk_diffusion_acf(data[sp]["Velocities"], plan, DATA_RANGE, 1,
                                                     time, U_LENGTH, U_TIME)
        ref = od.gk_diffusion_finish(acf_sum, count, sigmas, time, DATA_RANGE - 1)
        out = {"count": int(count), "n_windows": len(sigmas),
               "acf": list(map(float, ref["acf"])), "integral": list(map(float, ref["integral"])),
               "integral_uncertainty": list(map(float, ref["integral_uncertainty"])),
               "diffusion_coefficient": float(ref["diffusion_coefficient"][0]),
               "uncertainty": float(ref["uncertainty"][0])}
    return kind, sp, out


def main():
    jobs = [(k, sp) for k in ("green_kubo", "einstein") for sp in ("Na", "Cl")]
    out = {"config": {"n_atoms": N_ATOMS, "n_frames": N_FRAMES, "box": BOX, "seed": SEED,
                      "sigma_step": SIGMA, "data_range": DATA_RANGE, "time_step": TIME_STEP,
                      "sample_rate": SAMPLE_RATE, "units": "metal", "planner_memory": MEM},
           "einstein": {}, "green_kubo": {}}
    with ProcessPoolExecutor(max_workers=4) as ex:
        for kind, sp, res in ex.map(_job, jobs):
            out[kind][sp] = res
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c2_full.json")
    with open(path, "w") as fh:
        json.dump(out, fh)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
