"""Golden fixtures on NATURAL images (BASELINE config C3 and the PQ spaces of C4), from the REAL reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_natural.py

Writes tests/golden/golden_natural.npz + golden_natural.json.  Same recipe as make_golden.py (the unmodified
reference imported through oracle/ref_import.py, oracle mode S): for every case the uint8 input, the reference's
.ajpg bytes, its three edge maps (bit-packed), the sha-256 of its three downsampled layers, a 1-in-3 sample of the
decoded pixels and the PSNR.

Cases:
  * the six odd-sized images of the LIVE database (the general INTER_AREA path, CLAHE reflect padding and partial
    leaves on real data) and baboon / peppers, in OKLAB, quality 30-95, blocks 4-128 -- configuration C3
    (test/analysis/metrics_computation.py:297-335 sweeps this database);
  * one LIVE image per PQ space (ICaCb, ICtCp, JzAzBz), same settings -- the spaces of configuration C4.
The LIVE images are redistributed under the database's licence; its notice is copied next to the fixtures
(tests/golden/LIVE_copyright_notice.txt).
"""
from __future__ import annotations

import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

import ref_import  # noqa: E402
from make_golden import run_case  # noqa: E402

IMAGES = "/root/reference/test_images"
LIVE = os.path.join(IMAGES, "LIVE_image_quality_assessment_database")
ODD_LIVE = ["carnivaldolls", "cemetry", "churchandcapitol", "dancers", "manfishing", "studentsculpture"]
Q, B = (30, 95), (4, 128)


def main():
    from PIL import Image as PILImage
    R = ref_import.load(ipp=False)
    meta = {"versions": ref_import.versions(), "mode": "S (cv2.ipp.setUseIPP(False))", "cases": {}}
    arrays = {}

    def load(path):
        return np.ascontiguousarray(np.asarray(PILImage.open(path).convert("RGB")))

    def add(name, key, rgb_u8, space):
        res = run_case(R, rgb_u8, space, Q, B, False)
        arrays[f"{key}/input"] = rgb_u8
        for k, v in res.items():
            arrays[f"{name}/{k}"] = v
        meta["cases"][name] = {"space": space, "quality": list(Q), "blocks": list(B), "shape": list(rgb_u8.shape), "input": key,
                               "ajpg_bytes": int(res["ajpg"].size), "psnr": float(res["psnr"][0])}
        print(name, rgb_u8.shape, space, "->", res["ajpg"].size, "bytes, psnr", float(res["psnr"][0]), flush=True)

    for n in ODD_LIVE:
        add(f"live_{n}_OKLAB", f"live_{n}", load(os.path.join(LIVE, n + ".bmp")), "OKLAB")
    for n in ("baboon", "peppers"):
        add(f"{n}_OKLAB", n, load(os.path.join(IMAGES, n + ".tiff")), "OKLAB")
    for n, sp in (("cemetry", "ICaCb"), ("dancers", "ICtCp"), ("manfishing", "JzAzBz")):
        add(f"live_{n}_{sp}", f"live_{n}", load(os.path.join(LIVE, n + ".bmp")), sp)

    np.savez_compressed(os.path.join(HERE, "golden_natural.npz"), **arrays)
    with open(os.path.join(HERE, "golden_natural.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    shutil.copyfile(os.path.join(LIVE, "copyright_notice"), os.path.join(HERE, "LIVE_copyright_notice.txt"))
    print("wrote", os.path.getsize(os.path.join(HERE, "golden_natural.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
