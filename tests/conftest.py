import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "fem-libraries_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    # GPU tests are skipped (not failed) when no device is visible, e.g. a plain `pytest tests/`
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if has:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def square():
    """The reference's only shipped mesh (common/data/square.msh) as a fixture."""
    with open(os.path.join(GOLDEN, "square_mesh.json")) as f:
        m = json.load(f)
    x = np.array([[float(a), float(b)] for a, b in m["x"]], dtype=np.float64)
    tri = np.array(m["triangles"], dtype=np.int32)
    tag = np.array(m["triangle_tag"], dtype=np.int32)
    return {"x": x, "tri": tri, "tag": tag, "edges": np.array(m["edges"], dtype=np.int32),
            "edge_tag": np.array(m["edge_tag"], dtype=np.int32)}


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(GOLDEN, "square_kat.json")) as f:
        return json.load(f)
