import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "uw-com-vision_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


# UWCV_TEST_VARIANT=check runs the whole suite against lib/libuwcv_check.so (device-side bounds
# traps, the memory-safety run kept in profiles/); test infrastructure, not a product knob
if os.environ.get("UWCV_TEST_VARIANT"):
    from uwcv import _lib as _uwcv_lib
    _uwcv_lib.use_library_variant(os.environ["UWCV_TEST_VARIANT"])


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
