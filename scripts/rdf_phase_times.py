"""Where an RdfEngine batch spends its time (debug aid): CUDA-event and host-clock timings of the
sorted pack, the boxes, the extent check and the pair kernel for batches of a C4-like system."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from lammps_analysis_b200 import kernels as K  # noqa: E402
from lammps_analysis_b200.engine import RdfEngine  # noqa: E402
from lammps_analysis_b200.synthetic import device_fluid  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    F = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    L = (n / 0.02036) ** (1 / 3) if n == 100_000 else (n / 0.05) ** (1 / 3)
    dev = torch.device("cuda:0")
    traj = device_fluid(n, F, L, 4, dev)
    cutoff = L / 2 - 0.1
    eng = RdfEngine([n], [L] * 3, cutoff, int(cutoff / 0.01), drop_first=True, device=dev)
    frames = np.arange(F)
    eng.add_frames([traj], frames[:16])          # warm-up
    torch.cuda.synchronize()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    rows = []
    for k0 in range(0, F, 16):
        sel = frames[k0:k0 + 16]
        buf = eng._buffer(len(sel))
        e = [ev() for _ in range(5)]
        h = [time.perf_counter()]
        e[0].record()
        eng.pack_sorted([traj], sel, buf)
        h.append(time.perf_counter())
        e[1].record()
        bbox = eng.boxes(buf, len(sel))
        h.append(time.perf_counter())
        e[2].record()
        mm = K.coord_extent(buf, len(sel), eng.layout.n_pad)
        h.append(time.perf_counter())
        e[3].record()
        K.rdf_hist(buf, len(sel), eng.layout, eng.box, eng.cutoff, eng.nbins, eng.thr, eng.cut2,
                   eng.hist, eng.counter, bbox=bbox, wrapped=True)
        h.append(time.perf_counter())
        e[4].record()
        torch.cuda.synchronize()
        rows.append({"gpu_ms": [round(e[i].elapsed_time(e[i + 1]), 3) for i in range(4)],
                     "host_ms": [round(1e3 * (h[i + 1] - h[i]), 3) for i in range(4)]})
    print(json.dumps({"n": n, "phases": ["pack_sorted x16", "bbox", "coord_extent (sync)", "rdf_hist"],
                      "batches": rows}))
    t0 = time.perf_counter()
    eng.add_frames([traj], frames)
    torch.cuda.synchronize()
    print(json.dumps({"add_frames_s": time.perf_counter() - t0, "frames": F}))


main()
