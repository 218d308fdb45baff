import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# CPU tests that pin the oracle, the loader and the golden fixtures to the unmodified reference (libcmref.so).  They are
# the ANCHOR of every GPU parity claim, so they run in both invocations the driver uses: `-m "not gpu"` (they carry no
# gpu mark) and `-m gpu` (the hook below adds the mark for that run), and the GPU box's record shows the anchor green
# next to the kernels it vouches for.
ANCHOR_MODULES = {"test_oracle_vs_ref.py", "test_golden.py", "test_loader.py", "test_abi.py", "test_leaf_queries.py", "test_aas.py",
                  "test_sv_world.py"}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "anchor: pins the oracle / loader / fixtures to the reference; runs with and without -m gpu")


@pytest.hookimpl(tryfirst=True)
def pytest_collection_modifyitems(config, items):
    only_gpu = (config.getoption("markexpr", "") or "").strip() == "gpu"
    for it in items:
        if os.path.basename(str(it.fspath)) in ANCHOR_MODULES and it.get_closest_marker("gpu") is None:
            it.add_marker(pytest.mark.anchor)
            if only_gpu:
                it.add_marker(pytest.mark.gpu)
