"""Host-side tables of the codec: colour constants, subsampling, quantisation matrices, zigzag.

Everything here is settings-time work the reference also does on the host
(jpeg.py:36-174, 216-238, 688-766); the values are uploaded to the device once per settings change.
The expressions mirror the reference's (Python ``math.log``, ``int()`` truncation, numpy float32
``np.linalg.inv``) so that the resulting integers / float32 bit patterns are identical.
"""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np

f32 = np.float32

SPACE_ID = {"YCbCr": 0, "YCoCg": 1, "YCoCg-R": 2, "OKLAB": 3, "ICaCb": 4, "ICtCp": 5, "JzAzBz": 6, "XYZ": 7}
# spaces a Jpeg can be configured with (COLOR_SPACE_SETTINGS keys, jpeg.py:62-147)
CODEC_SPACES = ("ICaCb", "ICtCp", "JzAzBz", "OKLAB", "YCbCr", "YCoCg", "YCoCg-R")


def _m(rows):
    return np.array(rows, dtype=f32)


# single 3x3 spaces: forward / inverse (ycbcr.py:25-38, ycocg.py:25-55)
LINEAR = {
    "YCbCr": (_m([[0.299, 0.587, 0.114], [-0.168736, -0.331264, 0.5], [0.5, -0.418688, -0.081312]]),
              _m([[1.0, 0.000037, 1.401988], [1.0, -0.344113, -0.714104], [1.0, 1.771978, 0.000135]])),
    "YCoCg": (_m([[0.25, 0.5, 0.25], [0.5, 0.0, -0.5], [-0.25, 0.5, -0.25]]), _m([[1, 1, -1], [1, 0, 1], [1, -1, -1]])),
    "YCoCg-R": (_m([[0.25, 0.5, 0.25], [1.0, 0.0, -1.0], [-0.5, 1.0, -0.5]]), _m([[1.0, 0.5, -0.5], [1.0, 0.0, 0.5], [1.0, -0.5, -0.5]])),
}
# two-matrix spaces: XYZ->LMS and LMS'->space (oklab.py:27-44, icacb.py:142-156, ictcp.py:142-156, jzazbz.py:189-203)
NONLINEAR = {
    "OKLAB": (_m([[0.8189330101, 0.3618667424, -0.1288597137], [0.0329845436, 0.9293118715, 0.0361456387], [0.0482003018, 0.2643662691, 0.6338517070]]),
              _m([[0.2104542553, 0.7936177850, -0.0040720468], [1.9779984951, -2.4285922050, 0.4505937099], [0.0259040371, 0.7827717662, -0.8086757660]])),
    "ICaCb": (_m([[0.37613, 0.70431, -0.05675], [-0.21649, 1.14744, 0.05356], [0.02567, 0.16713, 0.74235]]),
              _m([[0.4949, 0.5037, 0.0015], [4.2854, -4.5462, 0.2609], [0.3605, 1.1499, -1.5105]])),
    "ICtCp": (_m([[0.3592, 0.6976, -0.0358], [-0.1922, 1.1004, 0.0755], [0.0070, 0.0749, 0.8434]]),
              _m([[0.5, 0.5, 0.0], [1.6137, -3.3234, 1.7097], [4.3781, -4.2455, -0.1325]])),
    "JzAzBz": (_m([[0.41478972, 0.579999, 0.0146480], [-0.2015100, 1.120649, 0.0531008], [-0.0166008, 0.264800, 0.6684799]]),
               _m([[0.5, 0.5, 0.0], [3.524, -4.066708, 0.542708], [0.199076, 1.096799, -1.295875]])),
}
# MIDPOINTS / SCALE_FACTORS (target range [-127,127]) of every space
NORMALISATION = {
    "YCbCr": ([0.5000000037252903, 7.450580596923828e-09, 0.0], [253.99999810755253, 254.000003784895, 254.0]),
    "YCoCg": ([0.5, 0.0, 0.0], [254.0, 254.0, 254.0]),
    "YCoCg-R": ([0.5, 0.0, 0.0], [254.0, 127.0, 127.0]),
    "OKLAB": ([0.4999999, 0.021152213, -0.056563325], [254.00005, 497.9055, 497.94604]),
    "ICaCb": ([0.07498085, 0.02180194, -0.018250957], [1693.7823, 1838.5665, 1330.3855]),
    "ICtCp": ([0.07497266, -0.0008235276, 0.023989676], [1693.9674, 1133.9044, 1694.004]),
    "JzAzBz": ([0.0087900255, 0.00048353244, -0.0020741792], [14448.194, 7590.505, 5552.201]),
    "XYZ": ([0.47523502, 0.50000006, 0.544415], [267.2362, 253.99997, 233.27792]),
}
# chroma (rh, rw); luma is never subsampled
CHROMA_SUBSAMPLING = {"ICaCb": (1, 4), "ICtCp": (1, 4), "JzAzBz": (2, 2), "OKLAB": (2, 2), "YCbCr": (2, 2), "YCoCg": (2, 2), "YCoCg-R": (2, 2)}


def color_tables(space: str):
    """(fwd1, fwd2, inv1, inv2, mid, scale) as float32 arrays for aeaj_set_color_tables."""
    z = np.zeros((3, 3), dtype=f32)
    if space in LINEAR:
        fwd1, inv1 = LINEAR[space]
        fwd2 = inv2 = z
    elif space in NONLINEAR:
        fwd1, fwd2 = NONLINEAR[space]
        inv1 = np.linalg.inv(fwd2)      # space -> LMS'   (float32, like the reference's class attributes)
        inv2 = np.linalg.inv(fwd1)      # LMS  -> XYZ
    else:                               # XYZ
        fwd1 = fwd2 = inv1 = inv2 = z
    mid, scale = NORMALISATION[space]
    return tuple(np.ascontiguousarray(a, dtype=f32) for a in (fwd1, fwd2, inv1, inv2, np.array(mid, dtype=f32), np.array(scale, dtype=f32)))


@lru_cache(maxsize=None)
def srgb_to_linear_lut() -> np.ndarray:
    """_srgb_to_linear_rgb (common.py:34-60) on the 256 values k/255 an 8-bit image can hold,
    evaluated in float64 with the host libm like the numba kernel does, stored as float32."""
    v = (np.arange(256, dtype=f32) / f32(255.0)).astype(np.float64)
    out = np.where(v <= 0.04045, v / 12.92, np.array([math.pow((x + 0.055) / 1.055, 2.4) for x in v]))
    return out.astype(f32)


# ----------------------------------------------------------------------------------------------
# quantisation
# ----------------------------------------------------------------------------------------------
LUMINANCE_Q = _m([[16, 11, 10, 16, 24, 40, 51, 61], [12, 12, 14, 19, 26, 58, 60, 55], [14, 13, 16, 24, 40, 57, 69, 56],
                  [14, 17, 22, 29, 51, 87, 80, 62], [18, 22, 37, 56, 68, 109, 103, 77], [24, 35, 55, 64, 81, 104, 113, 92],
                  [49, 64, 78, 87, 103, 121, 120, 101], [72, 92, 95, 98, 112, 100, 103, 99]])
CHROMINANCE_Q = _m([[17, 18, 24, 47, 99, 99, 99, 99], [18, 21, 26, 66, 99, 99, 99, 99], [24, 26, 56, 99, 99, 99, 99, 99],
                    [47, 66, 99, 99, 99, 99, 99, 99], [99] * 8, [99] * 8, [99] * 8, [99] * 8])


def block_sizes(block_size_range):
    lo, hi = block_size_range
    return [2 ** i for i in range(int(math.log2(lo)), int(math.log2(hi)) + 1)]


def quality_factor(block_size, quality_range, block_size_range) -> int:
    """Quality interpolated in log(block size); the smallest block gets quality_max (jpeg.py:688-705)."""
    bmin, bmax = block_size_range
    qmin, qmax = quality_range
    if bmin == bmax:
        return int((qmin + qmax) / 2)
    return int(qmin + (qmax - qmin) * (1 - math.log(block_size / bmin) / math.log(bmax / bmin)))


def _bilinear_square(m: np.ndarray, size: int) -> np.ndarray:
    """cv.resize(m, (size,size), INTER_LINEAR) for a square float32 matrix: half-pixel centres, edge
    clamp, horizontal then vertical, a*(1-t)+b*t in float32.  For the 8x8 integer tables and
    power-of-two sizes every weight is a dyadic fraction, so the float32 arithmetic is exact."""
    n = m.shape[0]
    if size == n:
        return m.astype(f32).copy()
    scale = 1.0 / (size / n)
    d = np.arange(size)
    fx = ((d + 0.5) * scale - 0.5).astype(f32)
    s = np.floor(fx).astype(np.int64)
    t = (fx - s.astype(f32)).astype(f32)
    lo = s < 0
    t[lo] = 0
    s[lo] = 0
    hi = s >= n - 1
    t[hi] = 0
    s[hi] = n - 1
    s1 = np.minimum(s + 1, n - 1)
    a0 = (f32(1.0) - t).astype(f32)
    m = m.astype(f32)
    h = (m[:, s] * a0[None, :]).astype(f32) + (m[:, s1] * t[None, :]).astype(f32)
    h = h.astype(f32)
    v = (h[s, :] * a0[:, None]).astype(f32) + (h[s1, :] * t[:, None]).astype(f32)
    return v.astype(f32)


def quantization_matrix(base8: np.ndarray, size: int, quality: int) -> np.ndarray:
    """Standard JPEG quality scaling of the 8x8 table, resized to size x size (jpeg.py:707-724)."""
    scale_factor = 5000 / quality if quality < 50 else 200 - 2 * quality
    scaled = np.floor((scale_factor * base8 + 50) / 100)
    resized = _bilinear_square(scaled.astype(f32), size)
    return np.clip(resized, 1, None).astype(np.int32)


def quantization_cache(quality_range, block_size_range):
    """{layer: {size: int32 matrix}} like Jpeg.quantization_matrix_cache (jpeg.py:228-238)."""
    out = {}
    for layer in range(3):
        base = LUMINANCE_Q if layer == 0 else CHROMINANCE_Q
        out[layer] = {s: quantization_matrix(base, s, quality_factor(s, quality_range, block_size_range))
                      for s in block_sizes(block_size_range)}
    return out


@lru_cache(maxsize=None)
def zigzag_ordering(size: int) -> np.ndarray:
    """Indices that flatten a size x size block in zigzag order (jpeg.py:726-766), built by sorting
    anti-diagonals: even diagonals run bottom-left -> top-right, odd ones the other way."""
    if not isinstance(size, int) or size < 0:
        raise ValueError("Block size must be a non-negative integer")
    r, c = np.divmod(np.arange(size * size), size) if size else (np.zeros(0, int), np.zeros(0, int))
    d = r + c
    key = np.where(d % 2 == 0, c, r)      # within a diagonal: increasing col when going up-right
    order = np.lexsort((key, d))
    return order.astype(np.int32)


def layer_shapes(height: int, width: int, space: str):
    rh, rw = CHROMA_SUBSAMPLING[space]
    return [(height, width), (height // rh, width // rw), (height // rh, width // rw)]
