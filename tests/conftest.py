"""pytest configuration: registers the `gpu` marker and makes sure the two native artefacts exist —
rbrt_b200/librbrt_gpu.so (the product, nvcc sm_100a) and oracle/build/librbrt_oracle.so (the CPU checker).
GPU tests never skip: on a box without a usable B200 they fail loudly (there is no CPU fallback)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def native_artefacts():
    gpu_so = os.path.join(ROOT, "rbrt_b200", "librbrt_gpu.so")
    ref_so = os.path.join(ROOT, "oracle", "build", "librbrt_oracle.so")
    if not os.path.exists(gpu_so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "rbrt_b200", "csrc")], check=True)
    if not os.path.exists(ref_so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True)
    yield


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_ffi
    oracle_ffi.lib()
    return oracle_ffi


@pytest.fixture(scope="session")
def gpu():
    """The product library on cuda:0.  Fails (does not skip) when the GPU path is unavailable."""
    import rbrt_b200 as R
    R.gpu_init(0)
    return R
