import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import c_oracle
    c_oracle.build()
    return c_oracle


@pytest.fixture(scope="session")
def emul_so():
    """host-emulation build of the kernel bodies (tests/emul/Makefile) -- test infrastructure only"""
    d = os.path.join(ROOT, "tests", "emul")
    subprocess.check_call(["make", "-C", d, "-s"], stdout=subprocess.DEVNULL)
    return os.path.join(d, "_build", "libbp_b200_emul.so")


@pytest.fixture(scope="session")
def product_so():
    so = os.path.join(ROOT, "bulletproofs_r1cs_gadgets_b200", "libbp_b200.so")
    if not os.path.exists(so):
        import __graft_entry__
        __graft_entry__.build()
    return so
