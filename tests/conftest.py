import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "2d-to-3d-style-transfer_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def cow(golden_dir):
    import numpy as np
    import torch
    d = np.load(os.path.join(golden_dir, "cow_mesh.npz"))
    return dict(verts=torch.from_numpy(d["verts"]), faces=torch.from_numpy(d["faces"]).long(),
                verts_uvs=torch.from_numpy(d["verts_uvs"]), faces_uvs=torch.from_numpy(d["faces_uvs"]).long(),
                texture=torch.from_numpy(d["texture"]).float() / 255.0)
