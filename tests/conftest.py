import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "zigz_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def po():
    """The CPU oracle (test infrastructure)."""
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def zlib():
    """The product library; built on demand here, prebuilt on the GPU box."""
    import zigz_b200
    if not os.path.exists(zigz_b200.LIB_PATH):
        zigz_b200.build()
    return zigz_b200


@pytest.fixture(scope="session")
def ctx(zlib):
    """One device context for the GPU tests. No skip: a missing GPU / library must fail loudly."""
    c = zlib.Context(0)
    yield c
    c.close()


@pytest.fixture(params=["default", "device_rounds"])
def rounds_mode(request, ctx):
    """Runs a GPU test twice: with the default schedule (small tables are handed to the host twin, which finishes the last
    rounds; d = 1 goes through block sums + multi-variable folds) and with EVERY round on the device (`linear_d1 = 0`,
    `prod_host_tail_log2 = 0`: round kernels, scalar and last-fold kernels, the persistent tail) — so the small-size and
    golden-vector cases keep exercising the CUDA kernels themselves, not only the host part of the default schedule."""
    keys = ("linear_d1", "prod_host_tail_log2")
    old = {k: ctx.get_option(k) for k in keys}
    if request.param == "device_rounds":
        ctx.set_option("linear_d1", 0)
        ctx.set_option("prod_host_tail_log2", 0)
    yield request.param
    for k, v in old.items():
        ctx.set_option(k, v)
