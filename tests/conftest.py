import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _emul_path():
    return os.path.join(ROOT, "tests", "emul", "_build", "libiamrx_emul.so")


@pytest.fixture(scope="session")
def emul_lib():
    """Host emulation build of the library sources (tests only): exercises the host-side
    logic (plans, multigrid drivers, time step) without a GPU."""
    import iamr_b200 as ix
    srcs = [os.path.join(ROOT, "iamr_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "iamr_b200", "csrc"))]
    srcs += [os.path.join(ROOT, "include", "iamrx.h"), os.path.join(ROOT, "tests", "emul", "cuda_emul.h")]
    so = _emul_path()
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-s", "-j8", "-C", ROOT, "emul"])
    os.environ.setdefault("OMP_WAIT_POLICY", "passive")
    return ix.load(so)


@pytest.fixture(scope="session")
def cuda_lib():
    import torch
    import iamr_b200 as ix
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    lib = ix.load()
    assert lib.iamrx_device_ok() == 1
    return lib


@pytest.fixture(scope="session")
def oracle():
    import orc
    orc.lib()
    return orc


@pytest.fixture(params=[pytest.param("emul"), pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    """(library, device): the host-emulation build on CPU, or the CUDA product library on a GPU.
    `-m "not gpu"` runs the first, `-m gpu` the second; the test bodies are shared."""
    if request.param == "emul":
        return request.getfixturevalue("emul_lib"), "cpu"
    return request.getfixturevalue("cuda_lib"), "cuda:0"
