"""Multi-GPU check, run under torchrun (one rank per GPU, NCCL; MDK_MG_BACKEND=gloo: two ranks on
one GPU, collectives through the host):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29611 tests/multigpu_check.py

Every rank opens the same synthetic experiment; the calculators shard frames (RDF) / atoms
(Einstein, Green-Kubo, ionic current) across ranks and all-reduce.  The sharded results must
equal the single-rank results: bit-exact for the integer histograms, 1e-12 for the fp64 series
(the reduction order differs)."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from lammps_analysis_b200 import distributed as D  # noqa: E402
from lammps_analysis_b200.config import config  # noqa: E402
from lammps_analysis_b200.file_io import ScriptInput  # noqa: E402
from lammps_analysis_b200.project import Project  # noqa: E402
from lammps_analysis_b200.synthetic import nacl_trajectory  # noqa: E402


def build(tag):
    data, box = nacl_trajectory(1000, 600, 32.0, seed=5, sigma_step=0.3)
    project = Project(tag, storage_path=tempfile.mkdtemp(prefix=f"mdk_mg_{tag}_"), persist=False)
    exp = project.add_experiment("NaCl", timestep=0.002, temperature=1400.0, units="metal")
    exp.add_data(ScriptInput(data, box, sample_rate=10, atom_major=True))
    exp.species["Na"].charge = 1.0
    exp.species["Cl"].charge = -1.0
    return exp


def run_all(exp):
    out = {}
    out["rdf"] = exp.run.RadialDistributionFunction(number_of_configurations=37, plot=False).data_dict
    out["ein"] = exp.run.EinsteinDiffusionCoefficients(data_range=120, plot=False).data_dict
    out["gk"] = exp.run.GreenKuboDiffusionCoefficients(data_range=120, plot=False).data_dict
    out["ion"] = exp.run.GreenKuboIonicConductivity(data_range=120, plot=False).data_dict
    return out


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    if os.environ.get("MDK_MG_BACKEND", "nccl") == "gloo":
        # a box with ONE GPU: both ranks compute on cuda:0 and exchange through the host (gloo):
        # the sharding logic, the frame exchange and the reductions are the production code, only
        # the transport differs
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    config.planner_memory_bytes = 60e9
    sharded = run_all(build(f"sharded{rank}"))
    with D.local_only():
        single = run_all(build(f"single{rank}"))
    for key in ("Na_Na", "Na_Cl", "Cl_Cl"):
        a, b = np.array(sharded["rdf"][key]["y"][1:]), np.array(single["rdf"][key]["y"][1:])
        assert np.array_equal(a, b), f"RDF {key} differs between {dist.get_world_size()} ranks and 1"
    for sp in ("Na", "Cl"):
        np.testing.assert_allclose(sharded["ein"][sp]["msd"], single["ein"][sp]["msd"], rtol=1e-9)
        # the ACF kernel sums fp32 products over runs of <= 512 atoms before folding into fp64;
        # a different atom partition re-associates those runs (parity tolerance is 1e-5)
        np.testing.assert_allclose(sharded["gk"][sp]["acf"], single["gk"][sp]["acf"], rtol=5e-6,
                                   atol=1e-7 * np.abs(single["gk"][sp]["acf"]).max())
        np.testing.assert_allclose(sharded["gk"][sp]["integral_uncertainty"],
                                   single["gk"][sp]["integral_uncertainty"], rtol=1e-4)
    np.testing.assert_allclose(sharded["ion"]["System"]["acf"], single["ion"]["System"]["acf"],
                               rtol=1e-6, atol=1e-9 * np.abs(single["ion"]["System"]["acf"]).max())
    dist.barrier()
    if rank == 0:
        print(f"multigpu_check ok on {dist.get_world_size()} ranks ({dist.get_backend()})")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
