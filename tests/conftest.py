"""pytest configuration: the `gpu` marker + shared fixtures.

CPU suite  (`-m "not gpu"`): oracle vs. the reference-derived golden vectors, the host-compiled
device routines (tests/hostemu) vs. the oracle, the C-ABI export list, the gloo multi-rank logic.
GPU suite  (`-m gpu`): the CUDA path through the C ABI (libkzgpu.so) vs. the oracle.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "hostemu"), os.path.join(ROOT, "oracle"),
          os.path.join(ROOT, "nano-kazen_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def kzo():
    import kzo_py
    kzo_py.build()
    return kzo_py


@pytest.fixture(scope="session")
def emu():
    import emu_py
    emu_py.build()
    return emu_py


@pytest.fixture(scope="session")
def gpu_lib():
    """Path of the product library; the GPU tests fail (not skip) when it is missing."""
    import pykazen as pk
    assert os.path.exists(pk.LIB_GPU), f"{pk.LIB_GPU} is missing: run __graft_entry__.build()"
    return pk.LIB_GPU
