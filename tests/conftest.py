import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "hierarchical-3d-gaussians_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run under gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _emulate_gpu_tests(tmp_path_factory):
    """H3DGS_EMULATE=1 (development aid, CPU only): run the `-m gpu` test files against the emulation build of the
    kernels (tests/emul/) -- device="cuda" in test code lands on the CPU.  Graph capture and NCCL tests cannot run
    this way; the big frames are slow.  Without the variable this fixture does nothing."""
    if os.environ.get("H3DGS_EMULATE") != "1":
        yield
        return
    sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
    from build_emu import build
    from fake_device import cpu_as_device, cuda_names_mean_cpu
    so = build(str(tmp_path_factory.mktemp("h3dgs_emu_session")))
    with cpu_as_device(so), cuda_names_mean_cpu():
        yield
