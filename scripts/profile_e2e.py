"""Host-side cProfile of the public calls the bench times end to end (debug aid, not a bench):
one RadialDistributionFunction call on one frame of the C5 system from a host-resident store."""
import cProfile
import io
import pstats
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from lammps_analysis_b200.file_io import ScriptInput  # noqa: E402
from lammps_analysis_b200.project import Project  # noqa: E402
from lammps_analysis_b200.synthetic import device_fluid  # noqa: E402


def main():
    n_sp = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
    n_frames = 6
    box_l = (2 * n_sp / 0.05) ** (1 / 3)
    dev = torch.device("cuda:0")
    sp = [device_fluid(n_sp, n_frames, box_l, 500 + s, dev) for s in range(2)]
    project = Project("prof", storage_path=tempfile.mkdtemp(prefix="mdk_prof_"), persist=False)
    exp = project.add_experiment("rdf", timestep=0.002, temperature=300.0, units="real")
    exp.add_data(ScriptInput({"A": {"Positions": sp[0].cpu().numpy()},
                              "B": {"Positions": sp[1].cpu().numpy()}}, [box_l] * 3,
                             atom_major=True))
    del sp

    def one(i):
        exp.store.invalidate()
        exp.version += 1
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        exp.run.RadialDistributionFunction(start=i, stop=i, number_of_configurations=1, plot=False)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    print("warm", one(0), one(1))
    prof = cProfile.Profile()
    prof.enable()
    ts = [one(i) for i in (2, 3, 4)]
    prof.disable()
    print("timed", ts)
    s = io.StringIO()
    pstats.Stats(prof, stream=s).sort_stats("cumulative").print_stats(45)
    print(s.getvalue())


main()
