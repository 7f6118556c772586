"""Generate the golden fixtures under tests/golden/ by running the REAL reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every fixture is an output of the unmodified reference imported through oracle/ref_import.py, in
oracle mode S (cv2.ipp.setUseIPP(False): OpenCV's open-source arithmetic, the bit-exact target)
unless the key says ``_modeD`` (library default, IPP on).  The fixtures pin the CPU oracle
(tests/test_oracle_golden.py, runs without a GPU); the GPU tests then compare the CUDA path with
the pinned oracle and with these same fixtures.  Library versions are recorded in golden.json.

Inputs that are not synthetic (lena) are stored as uint8 so the fixtures are self-contained on the
GPU box, where /root/reference does not exist.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_import  # noqa: E402
from synth import synth  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_case(R, rgb_u8: np.ndarray, space, qrange, brange, store_stages: bool):
    """Run Jpeg.compress/decompress stage by stage exactly as jpeg.py:240-297 does."""
    img = R.Image.from_array((rgb_u8.astype(np.float32) / 255.0).astype(np.float32), None, ".png")
    H, W, _ = img.data.shape
    j = R.Jpeg(R.JpegCompressionSettings(space, qrange, brange))
    ajpg = j.compress(img)
    out = {"ajpg": np.frombuffer(ajpg, dtype=np.uint8)}
    conv = j._convert_color_space(img.get_flattened()).reshape(H, W, 3).transpose(2, 0, 1)
    ds = j._downsample(conv)
    cv = R.cv2
    for i, lay in enumerate(ds):
        edge = R.EdgeDetection.canny(lay)
        out[f"edge{i}"] = np.packbits(edge.astype(np.uint8), axis=None)
        out[f"layer_sha{i}"] = np.frombuffer(bytes.fromhex(sha(lay)), dtype=np.uint8)
        if store_stages:
            u8 = (lay * 255).astype(np.uint8)
            cl = cv.createCLAHE(clipLimit=0.75, tileGridSize=(4, 4)).apply(u8)
            g = cv.GaussianBlur(cl, (3, 3), 0)
            b = cv.bilateralFilter(g, 5, 75, 75)
            out[f"layer{i}"] = lay.astype(np.float32)
            out[f"u8_{i}"], out[f"clahe{i}"], out[f"gauss{i}"], out[f"bil{i}"] = u8, cl, g, b
            out[f"thr{i}"] = np.array([np.percentile(b, 10.0), np.percentile(b, 30.0)], dtype=np.float64)
    dec = R.Jpeg(R.JpegCompressionSettings()).decompress(ajpg)
    # decoded pixels: a 1-in-3 x 1-in-3 sample keeps the fixture small; the comparison is a
    # tolerance test (<= 1 LSB in the truncating 8-bit view, image.py:127), not a hash
    out["decoded_u8_s3"] = np.ascontiguousarray((dec.data * 255).astype(np.uint8)[::3, ::3])
    if store_stages:
        out["decoded_s3"] = np.ascontiguousarray(dec.data.astype(np.float32)[::3, ::3])
    mse = float(np.mean((dec.data.astype(np.float64) - img.data.astype(np.float64)) ** 2))
    out["psnr"] = np.array([10 * np.log10(1.0 / mse) if mse > 0 else 999.0])
    return out


def main():
    R = ref_import.load(ipp=False)
    cv = R.cv2
    meta = {"versions": ref_import.versions(), "mode": "S (cv2.ipp.setUseIPP(False))", "cases": {}}
    arrays = {}

    def add(name, rgb_u8, space, q, b, stages=False, store_input=True):
        res = run_case(R, rgb_u8, space, q, b, stages)
        if store_input:
            arrays[f"{name}/input"] = rgb_u8
        for k, v in res.items():
            arrays[f"{name}/{k}"] = v
        meta["cases"][name] = {"space": space, "quality": list(q), "blocks": list(b), "shape": list(rgb_u8.shape),
                               "stages": stages, "ajpg_bytes": int(res["ajpg"].size)}
        print(name, rgb_u8.shape, space, q, b, "->", res["ajpg"].size, "bytes", flush=True)

    from PIL import Image as PILImage
    lena = np.asarray(PILImage.open("/root/reference/test_images/lena.png").convert("RGB"))
    arrays["lena/input"] = lena
    # C1: lena 512x512 YCbCr q50-90 b4-64 (+ the q_min sweep of SURVEY 8d)
    add("lena_ycbcr_q50_90_b4_64", lena, "YCbCr", (50, 90), (4, 64), store_input=False)
    crop = np.ascontiguousarray(lena[128:384, 160:416])
    arrays["lena256/input"] = crop
    for sp in ["YCoCg", "YCoCg-R", "OKLAB", "ICaCb", "ICtCp", "JzAzBz"]:
        add(f"lena256_{sp}_q30_95_b4_128", crop, sp, (30, 95), (4, 128), store_input=False)
    add("lena256_YCbCr_stages", crop, "YCbCr", (40, 80), (4, 64), stages=True, store_input=False)
    # odd sizes: general INTER_AREA path, CLAHE reflect padding, partial leaves, [1,4] subsampling
    odd = (synth(135, 241, seed=3) * 255 + 0.5).astype(np.uint8)
    arrays["synth135x241/input"] = odd
    for sp in ["YCbCr", "ICtCp", "JzAzBz"]:
        add(f"synth135x241_{sp}", odd, sp, (30, 95), (4, 128), stages=(sp == "YCbCr"), store_input=False)
    # small / degenerate shapes
    rng = np.random.default_rng(7)
    for (h, w) in [(4, 4), (5, 7), (16, 16), (33, 17), (8, 64)]:
        im = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        add(f"rand{h}x{w}_YCoCg", im, "YCoCg", (40, 80), (4, 64), stages=True)
    add("rand8x64_ICaCb", rng.integers(0, 256, (8, 64, 3), dtype=np.uint8), "ICaCb", (40, 80), (4, 16), stages=True)
    # quality sweep (C5 style) and block-range variants
    s64 = (synth(96, 160, seed=5) * 255 + 0.5).astype(np.uint8)
    arrays["synth96x160/input"] = s64
    for qmin in (1, 25, 50, 75, 99):
        add(f"synth96x160_YCoCg_q{qmin}_99", s64, "YCoCg", (qmin, 99), (4, 64), store_input=False)
    s300 = (synth(300, 260, seed=9) * 255 + 0.5).astype(np.uint8)
    arrays["synth300x260/input"] = s300
    for b in [(2, 256), (8, 8), (16, 32), (4, 4), (64, 128)]:
        add(f"synth300x260_YCbCr_b{b[0]}_{b[1]}", s300, "YCbCr", (30, 95), b, store_input=False)

    # colour conversion vectors: forward and inverse for every space (conversion.py:95-124)
    rgb = (rng.integers(0, 256, (2048, 3)).astype(np.float32) / 255.0).astype(np.float32)
    arrays["color/rgb"] = rgb
    for sp in ["YCbCr", "YCoCg", "YCoCg-R", "OKLAB", "ICaCb", "ICtCp", "JzAzBz"]:
        fwd = R.convert("sRGB", sp, rgb)
        arrays[f"color/fwd_{sp}"] = fwd
        pert = (fwd + rng.standard_normal(fwd.shape).astype(np.float32) * 0.002 * np.abs(fwd).max(axis=0)).astype(np.float32)
        arrays[f"color/inv_in_{sp}"] = pert
        arrays[f"color/inv_{sp}"] = R.convert(sp, "sRGB", pert)
        for ch in range(3):
            col = np.zeros((fwd.shape[0], 3), dtype=np.float32)
            col[:, ch] = fwd[:, ch]
            arrays[f"color/norm_{sp}_{ch}"] = R.apply_normalization(sp, col, False)[:, ch].copy()
            arrays[f"color/denorm_{sp}_{ch}"] = R.apply_normalization(sp, col * 100.0, True)[:, ch].copy()

    # quantisation matrices: every size 2..256, every quality 1..99, both tables -> one hash per (table,size)
    qh = {}
    J = R.Jpeg
    S = R.JpegCompressionSettings
    for tname, base in (("luma", S.LUMINANCE_QUANTIZATION_MATRIX), ("chroma", S.CHROMINANCE_QUANTIZATION_MATRIX)):
        for size in (2, 4, 8, 16, 32, 64, 128, 256):
            allq = np.stack([J._get_quantization_matrix(base, size, q) for q in range(1, 100)])
            qh[f"{tname}_{size}"] = sha(allq.astype(np.int32))
    meta["qmatrix_sha"] = qh
    # same with IPP on, to document whether the q tables depend on the IPP switch
    cv.ipp.setUseIPP(True)
    qd = {}
    for tname, base in (("luma", S.LUMINANCE_QUANTIZATION_MATRIX), ("chroma", S.CHROMINANCE_QUANTIZATION_MATRIX)):
        for size in (2, 4, 8, 16, 32, 64, 128, 256):
            allq = np.stack([J._get_quantization_matrix(base, size, q) for q in range(1, 100)])
            qd[f"{tname}_{size}"] = sha(allq.astype(np.int32))
    meta["qmatrix_sha_modeD_equal"] = qd == qh
    # one mode-D end-to-end record (library default) for the parity report
    res = run_case(R, crop, "YCbCr", (40, 80), (4, 64), False)
    for k in ("ajpg", "edge0", "edge1", "edge2", "decoded_u8_s3", "psnr"):
        arrays[f"lena256_YCbCr_modeD/{k}"] = res[k]
    meta["cases"]["lena256_YCbCr_modeD"] = {"space": "YCbCr", "quality": [40, 80], "blocks": [4, 64],
                                            "shape": list(crop.shape), "stages": False, "mode": "D"}
    cv.ipp.setUseIPP(False)

    # quality-factor table (jpeg.py:688-705)
    qf = {}
    for (bmin, bmax) in [(4, 64), (4, 128), (2, 256), (8, 8), (16, 32)]:
        for (qmin, qmax) in [(50, 90), (30, 95), (1, 99), (40, 80), (99, 99)]:
            jj = R.Jpeg(R.JpegCompressionSettings("YCbCr", (qmin, qmax), (bmin, bmax)))
            sizes = sorted(jj.quantization_matrix_cache[0].keys())
            qf[f"{bmin},{bmax},{qmin},{qmax}"] = [int(jj._get_quality_factor(s)) for s in sizes]
    meta["quality_factor"] = qf
    meta["zigzag_sha"] = {str(s): sha(R.Jpeg._zigzag_ordering(s)) for s in (2, 4, 8, 16, 32, 64, 128)}

    np.savez_compressed(os.path.join(HERE, "golden.npz"), **arrays)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", os.path.getsize(os.path.join(HERE, "golden.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
