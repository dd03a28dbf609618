#!/usr/bin/env python
"""Correlation kernels over the whole lag range: atom-lag updates/s and the fraction of BOTH
rooflines (HBM: 12 B per atom-frame read once; FP32: 9 FLOP per MSD update, 6 per ACF update)
for data_range = 2 .. 500 on a 125,000-atom x 2,000-frame shard.  Environment switches select
tuning variants of libmdk (read at launch time).

    python scripts/lag_sweep.py [--atoms 125000] [--frames 2000] [--variants default,g2,s32]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lammps_analysis_b200 import kernels as K  # noqa: E402
from lammps_analysis_b200.engine import acf_series, msd_series, plan_windows  # noqa: E402

VARIANTS = {
    "default": {},
    "norw": {"MDK_MSD_RW_MAX": "0", "MDK_ACF_RW_MAX": "0"},   # round-1 dispatch: stream <= 16, ring beyond
    "rw4": {"MDK_MSD_RW_MIN": "2"},                            # register-window kernel from 2 lags
    "nostream": {"MDK_MSD_NO_STREAM": "1", "MDK_ACF_NO_STREAM": "1"},
}


def timed(fn, flush, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--atoms", type=int, default=125_000)
    ap.add_argument("--frames", type=int, default=2000)
    ap.add_argument("--ranges", default="2,4,8,12,16,24,32,48,64,100,200,300,500")
    ap.add_argument("--variants", default="default")
    ap.add_argument("--what", default="msd,acf")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    A, T = args.atoms, args.frames
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    x = torch.cumsum(torch.randn(A, T, 3, device=dev, generator=gen) * 0.1, dim=1).contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    hbm = 6531.9
    peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                         "MEASURED_PEAKS.json")
    if os.path.exists(peaks):
        hbm = float(json.load(open(peaks))["hbm_gbs"])
    fp32 = K.peak_fp32(True)
    plan = dict(batch_size=T, n_batches=1, remainder=0, minibatch=False)
    for var in args.variants.split(","):
        for v in VARIANTS.values():
            for k in v:
                os.environ.pop(k, None)
        os.environ.update(VARIANTS[var])
        for N in [int(v) for v in args.ranges.split(",")]:
            launches = plan_windows(plan, N, 1, A)
            upd = launches[0][4] * A * N
            row = {"variant": var, "data_range": N}
            if "msd" in args.what:
                t = timed(lambda: msd_series(x, launches, N, 1, np.arange(N)), flush)
                row.update(msd_updates_per_s=upd / t, msd_frac_fp32=9 * upd / t * 1e-12 / fp32,
                           msd_frac_hbm=12.0 * A * T / t * 1e-9 / hbm)
            if "acf" in args.what:
                t = timed(lambda: acf_series(x, launches, N, 1, per_window=False), flush)
                row.update(acf_updates_per_s=upd / t, acf_frac_fp32=6 * upd / t * 1e-12 / fp32,
                           acf_frac_hbm=12.0 * A * T / t * 1e-9 / hbm)
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
