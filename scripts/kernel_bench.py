"""Kernel-level timings (CUDA events) for tuning; not the driver's bench (see bench.py)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lammps_analysis_b200 import kernels as K  # noqa: E402
from lammps_analysis_b200.engine import RdfEngine  # noqa: E402
from lammps_analysis_b200.synthetic import device_fluid  # noqa: E402


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="peak,rdf,msd,acf,unwrap,ionic")
    ap.add_argument("--rdf-atoms", type=int, default=100000)
    ap.add_argument("--rdf-frames", type=int, default=2)
    ap.add_argument("--rdf-box", type=float, default=170.0)
    ap.add_argument("--tunings", default="0")
    ap.add_argument("--sort", type=int, default=-1, help="-1 auto, 0 off, 1 on")
    ap.add_argument("--cutoff-frac", type=float, default=0.0)
    ap.add_argument("--adf-atoms", type=int, default=100000)
    ap.add_argument("--adf-frames", type=int, default=4)
    ap.add_argument("--dyn-atoms", type=int, default=20000)
    ap.add_argument("--dyn-frames", type=int, default=2000)
    ap.add_argument("--data-range", type=int, default=500)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    what = args.what.split(",")
    out = {}
    if "peak" in what:
        out["fp32_ffma_tflops"] = K.peak_fp32(False)
        out["fp32_ffma2_tflops"] = K.peak_fp32(True)
        print(json.dumps(out), flush=True)
    if "rdf" in what:
        n, F, L = args.rdf_atoms, args.rdf_frames, args.rdf_box
        traj = device_fluid(n, F, L, 4, dev)
        cutoff = L / 2 - 0.1 if args.cutoff_frac <= 0 else args.cutoff_frac * L
        nbins = int(cutoff / 0.01)
        sort = None if args.sort < 0 else bool(args.sort)
        for tun in [int(t, 0) for t in args.tunings.split(",")]:
            eng = RdfEngine([n], [L, L, L], cutoff, nbins, drop_first=False, device=dev,
                            spatial_sort=sort)
            t = timed(lambda: eng.add_frames([traj], np.arange(F), check_extent=args.sort >= -1, tuning=tun),
                      reps=2)
            pairs = F * n * (n - 1) / 2
            inside = eng.counts().sum() / eng.frames_done
            r = {"rdf_tuning": hex(tun), "sorted": eng.spatial_sort, "n": n, "F": F, "nbins": nbins, "s": t,
                 "pairs_per_s": pairs / t, "tflops20": 20 * pairs / t * 1e-12,
                 "inside_frac": float(inside / (n * (n - 1) / 2))}
            print(json.dumps(r), flush=True)
    if "adf" in what:
        # triplet-angle histograms (csrc/adf.cu): two species at rho = 0.05 / A^3, cutoff 6 A
        # (~45 neighbours, ~2,000 ordered neighbour pairs per centre atom)
        from lammps_analysis_b200.engine import AdfEngine

        n, F = args.adf_atoms, args.adf_frames
        L = (n / 0.05) ** (1.0 / 3.0)
        gen = torch.Generator(device=dev)
        gen.manual_seed(11)
        pos = torch.rand(F, n, 3, device=dev, generator=gen) * L
        eng = AdfEngine([n // 2, n - n // 2], [L] * 3, 6.0, 500, 4, device=dev)
        counts = {}

        def run():
            _, c = eng.add_batch(pos)
            counts["triples"] = c
        t = timed(run, reps=2)
        triples = int(counts["triples"].sum().item())
        print(json.dumps({"adf_atoms": n, "adf_frames": F, "cutoff": 6.0, "nbins": 500, "s": t,
                          "triples_histogrammed": triples, "triples_per_s": triples / t,
                          "centre_atoms_per_s": n * F / t}), flush=True)
    A, T, N = args.dyn_atoms, args.dyn_frames, args.data_range
    if set(what) & {"msd", "acf", "unwrap", "ionic"}:
        traj = device_fluid(A, T, 60.0, 5, dev)
    if "unwrap" in what:
        o = torch.empty_like(traj)
        ci = torch.zeros(A, 3, dtype=torch.float64, device=dev)
        cp = torch.zeros(A, 3, dtype=torch.float32, device=dev)
        for L in (60.0, 60.1):
            t = timed(lambda: K.unwrap(traj, [L] * 3, cp, ci, False, o))
            print(json.dumps({"unwrap_s": t, "box": L, "atom_frames_per_s": A * T / t,
                              "GBps": 24 * A * T / t * 1e-9}), flush=True)
    if "ionic" in what:
        J = torch.zeros(T, 3, dtype=torch.float64, device=dev)
        t = timed(lambda: K.ionic_current(traj, 1.0, J))
        print(json.dumps({"ionic_s": t, "atom_frames_per_s": A * T / t,
                          "GBps": 12 * A * T / t * 1e-9}), flush=True)
    if "msd" in what:
        W = T - N
        tau = torch.arange(N, dtype=torch.int32, device=dev)
        o = torch.zeros(N, dtype=torch.float64, device=dev)
        t = timed(lambda: K.msd_windowed(traj, 0, A, 0, W, 1, tau, N, o))
        upd = W * A * N
        print(json.dumps({"msd_legacy_s": t, "updates_per_s": upd / t}), flush=True)
        t = timed(lambda: K.msd_windowed(traj, 0, A, 0, W, 1, tau, N, o, dense=True))
        print(json.dumps({"msd_s": t, "updates_per_s": upd / t, "tflops9": 9 * upd / t * 1e-12,
                          "GBps": 12 * A * T / t * 1e-9}), flush=True)
    if "acf" in what:
        W = T - N
        o = torch.zeros(N, dtype=torch.float64, device=dev)
        win = torch.zeros(W, N, dtype=torch.float64, device=dev)
        scratch = torch.empty(T * N, dtype=torch.float64, device=dev)
        t = timed(lambda: K.acf_windowed(traj, 0, A, 0, T, N, W, 1, o, win, scratch))
        upd = W * A * N
        print(json.dumps({"acf_s": t, "updates_per_s": upd / t, "tflops6": 6 * upd / t * 1e-12,
                          "GBps": 12 * A * T / t * 1e-9}), flush=True)


if __name__ == "__main__":
    main()
