"""Import the *real* reference (``/root/reference/src``) inside the build container.

TEST INFRASTRUCTURE ONLY.  This module exists so that the oracle restatement in
``oracle/`` can be pinned against outputs of the reference itself and so that
``tests/golden/make_golden.py`` can generate golden vectors.  ``/root/reference``
does not exist on the GPU box, therefore nothing in ``-m gpu`` tests, ``smoke()``
or ``bench.py`` imports this file.

The reference cannot be imported as-is here: ``image/__init__.py`` pulls in
``lpips`` / ``piq`` (evaluation_metrics.py:21,23) and ``image.py:20`` pulls in
``imageio.v3``; none are installed.  They are stubbed *before* the import (only
file I/O and quality metrics are affected, neither is on the hot path).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(REFERENCE_SRC)


def _install_stubs() -> None:
    import numpy as np

    if "imageio" not in sys.modules:
        iio = types.ModuleType("imageio")
        v3 = types.ModuleType("imageio.v3")

        def imread(path, *a, **k):
            from PIL import Image as PILImage
            return np.asarray(PILImage.open(path))

        def imwrite(path, arr, *a, **k):
            from PIL import Image as PILImage
            PILImage.fromarray(arr).save(path)

        v3.imread = imread
        v3.imwrite = imwrite
        iio.v3 = v3
        sys.modules["imageio"] = iio
        sys.modules["imageio.v3"] = v3
    if "lpips" not in sys.modules:
        lp = types.ModuleType("lpips")

        class LPIPS:  # instantiated at class-body time in evaluation_metrics.py:34-36
            def __init__(self, *a, **k):
                pass

            def __call__(self, *a, **k):
                raise RuntimeError("lpips is stubbed")

        lp.LPIPS = LPIPS
        sys.modules["lpips"] = lp
    if "piq" not in sys.modules:
        sys.modules["piq"] = types.ModuleType("piq")


def load(ipp: bool | None = None):
    """Return a namespace with the reference's public objects.

    ipp: None = library default ("mode D"), False = cv2.ipp.setUseIPP(False) ("mode S").
    """
    if not available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/aeaj_numba_cache")
    _install_stubs()
    # our own drop-in packages use the same top-level names; make sure the reference wins here
    for name in ("jpeg", "color", "image"):
        mod = sys.modules.get(name)
        if mod is not None and REFERENCE_SRC not in (getattr(mod, "__file__", "") or ""):
            for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
                del sys.modules[k]
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import cv2
    if ipp is not None:
        cv2.ipp.setUseIPP(bool(ipp))
    from image import Image
    from jpeg import Jpeg, JpegCompressionSettings
    from jpeg.edge_detection import EdgeDetection
    from jpeg.quadtree import QuadTree
    from jpeg.utils import largest_power_of_2
    from color import convert, apply_normalization, get_color_spaces
    ns = types.SimpleNamespace(
        Image=Image, Jpeg=Jpeg, JpegCompressionSettings=JpegCompressionSettings,
        EdgeDetection=EdgeDetection, QuadTree=QuadTree, largest_power_of_2=largest_power_of_2,
        convert=convert, apply_normalization=apply_normalization,
        get_color_spaces=lambda: sorted(get_color_spaces()), cv2=cv2,
    )
    return ns


def versions() -> dict:
    import cv2
    import numba
    import numpy
    return {
        "cv2": cv2.__version__, "ipp": cv2.ipp.getIppVersion(), "use_ipp": bool(cv2.ipp.useIPP()),
        "numpy": numpy.__version__, "numba": numba.__version__,
        "python": sys.version.split()[0],
    }
