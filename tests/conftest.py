"""pytest configuration: registers the `gpu` marker and puts the product package, the oracle and tools on sys.path.

  python -m pytest tests -q -m "not gpu"   # CPU: oracle vs goldens, ABI, surface, host logic (gloo)
  python -m pytest tests -q -m gpu         # B200: CUDA path vs oracle / goldens / full-size properties
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "fov-3dgs_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def scene_small():
    from fovgs import synth
    return synth.make_scene_cube(10000, 0), synth.config1_camera()
