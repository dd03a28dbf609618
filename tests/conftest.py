import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _libmdk_built():
    """libmdk.so is a build artefact (git-ignored): build it once when a fresh checkout runs the
    tests before __graft_entry__.build() (nvcc cross-compiles without a GPU)."""
    lib = os.path.join(ROOT, "lammps_analysis_b200", "libmdk.so")
    if not os.path.exists(lib):
        import subprocess

        subprocess.run(["bash", os.path.join(ROOT, "lammps_analysis_b200", "csrc", "build.sh")],
                       check=True)


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
