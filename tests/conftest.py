import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "adaptive-edge-aware-jpeg_b200")
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "oracle"), PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """tests/golden/golden.npz + golden.json (synthetic inputs, lena; make_golden.py) and golden_natural.npz + .json (LIVE,
    baboon, peppers; make_golden_natural.py): outputs of the real reference (mode S)."""

    def __init__(self):
        d = os.path.join(ROOT, "tests", "golden")
        self.files = [np.load(os.path.join(d, "golden.npz")), np.load(os.path.join(d, "golden_natural.npz"))]
        with open(os.path.join(d, "golden.json")) as f:
            self.meta = json.load(f)
        with open(os.path.join(d, "golden_natural.json")) as f:
            nat = json.load(f)
        self.natural_cases = sorted(nat["cases"])
        self.meta["cases"].update(nat["cases"])

    def case(self, name):
        return self.meta["cases"][name]

    def _find(self, key):
        for z in self.files:
            if key in z.files:
                return z[key]
        raise KeyError(key)

    def get(self, name, key):
        return self._find(f"{name}/{key}")

    def input_u8(self, name):
        for k in (self.meta["cases"].get(name, {}).get("input"), name, name.split("_")[0]):
            if k is None:
                continue
            try:
                return self._find(f"{k}/input")
            except KeyError:
                pass
        raise KeyError(name)

    def input_f32(self, name):
        return (self.input_u8(name).astype(np.float32) / 255.0).astype(np.float32)


@pytest.fixture(scope="session")
def golden():
    return Golden()
