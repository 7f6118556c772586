import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "adaptive-edge-aware-jpeg_b200")
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """tests/golden/golden.npz + golden.json: outputs of the real reference (mode S)."""

    def __init__(self):
        d = os.path.join(ROOT, "tests", "golden")
        self.npz = np.load(os.path.join(d, "golden.npz"))
        with open(os.path.join(d, "golden.json")) as f:
            self.meta = json.load(f)

    def case(self, name):
        return self.meta["cases"][name]

    def get(self, name, key):
        return self.npz[f"{name}/{key}"]

    def input_u8(self, name):
        for k in (name, name.split("_")[0]):
            if f"{k}/input" in self.npz.files:
                return self.npz[f"{k}/input"]
        raise KeyError(name)

    def input_f32(self, name):
        return (self.input_u8(name).astype(np.float32) / 255.0).astype(np.float32)


@pytest.fixture(scope="session")
def golden():
    return Golden()
