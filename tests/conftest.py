import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_generate_tests(metafunc):
    """Every GPU test runs against BOTH vote kernels (grouped / one-hit-per-pass): the library picks one
    per model by bucket length, the tests force each in turn through the PPF_B200_VOTE hook."""
    if metafunc.definition.get_closest_marker("gpu") is not None:
        metafunc.parametrize("vote_kernel", ["grouped", "classic"], indirect=True)


@pytest.fixture(autouse=True)
def vote_kernel(request, monkeypatch):
    kind = getattr(request, "param", None)
    if kind in ("grouped", "classic"):
        monkeypatch.setenv("PPF_B200_VOTE", kind)
    yield kind


def _make(target_dir, *args):
    subprocess.run(["make", "-C", os.path.join(ROOT, target_dir), *args], check=True, capture_output=True)


@pytest.fixture(scope="session", autouse=True)
def built_artifacts():
    """Make sure the CUDA library and the oracles exist (nvcc cross-compiles without a GPU)."""
    if not os.path.exists(os.path.join(ROOT, "objective_slam_b200", "lib", "libppf_b200.so")):
        _make("objective_slam_b200/csrc", "-j4")
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        _make("oracle", "liboracle.so")
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libppf_ref.so")) and os.path.isdir("/root/reference"):
        _make("oracle", "ref")
    yield


def have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libppf_ref.so"))


def golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
