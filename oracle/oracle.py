"""CPU oracle for the adaptive edge-aware JPEG hot path -- Python side.

TEST INFRASTRUCTURE ONLY (see oracle/aeaj_oracle.c header).  Imported by tests/,
``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` -- never by the product
package.  Parity status: PINNED against the reference itself (tests/golden/make_golden.py).

Stage arithmetic lives in ``aeaj_oracle.c``; this file restates the reference's *host* logic:
settings tables (jpeg.py:36-174), layer shapes (jpeg.py:676-686), quality interpolation and
quantisation matrices (jpeg.py:688-724), zigzag (jpeg.py:726-766), the .ajpg container
(jpeg.py:531-674) and the compress / decompress drivers (jpeg.py:240-297).
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import subprocess
import zlib
from io import BytesIO

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libaeaj_oracle.so")

SPACE_ID = {"YCbCr": 0, "YCoCg": 1, "YCoCg-R": 2, "OKLAB": 3, "ICaCb": 4, "ICtCp": 5, "JzAzBz": 6, "XYZ": 7}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "aeaj_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.ao_root_size.restype = C.c_int
        _lib.ao_quadtree.restype = C.c_int
        _lib.ao_states_to_leaves.restype = C.c_int
        _lib.ao_get_max_threads.restype = C.c_int
    return _lib


def set_threads(n: int) -> None:
    lib().ao_set_threads(C.c_int(n))


def max_threads() -> int:
    return lib().ao_get_max_threads()


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ----------------------------------------------------------------------------------------------
# colour tables
# ----------------------------------------------------------------------------------------------
f32 = np.float32

_RGB2XYZ = np.array([[0.4124564, 0.3575761, 0.1804375], [0.2126729, 0.7151522, 0.0721750],
                     [0.0193339, 0.1191920, 0.9503041]], dtype=f32)          # xyz.py:27-32
_XYZ2RGB = np.array([[3.2404542, -1.5371385, -0.4985314], [-0.9692660, 1.8760108, 0.0415560],
                     [0.0556434, -0.2040259, 1.0572252]], dtype=f32)         # xyz.py:35-40

_FWD = {
    # linear spaces: (forward 3x3, inverse 3x3)
    "YCbCr": (np.array([[0.299, 0.587, 0.114], [-0.168736, -0.331264, 0.5], [0.5, -0.418688, -0.081312]], dtype=f32),
              np.array([[1.0, 0.000037, 1.401988], [1.0, -0.344113, -0.714104], [1.0, 1.771978, 0.000135]], dtype=f32)),
    "YCoCg": (np.array([[0.25, 0.5, 0.25], [0.5, 0.0, -0.5], [-0.25, 0.5, -0.25]], dtype=f32),
              np.array([[1, 1, -1], [1, 0, 1], [1, -1, -1]], dtype=f32)),
    "YCoCg-R": (np.array([[0.25, 0.5, 0.25], [1.0, 0.0, -1.0], [-0.5, 1.0, -0.5]], dtype=f32),
                np.array([[1.0, 0.5, -0.5], [1.0, 0.0, 0.5], [1.0, -0.5, -0.5]], dtype=f32)),
}
_NL = {
    # nonlinear spaces: (XYZ->LMS, LMS'->space); inverses are np.linalg.inv of the f32 matrices
    "OKLAB": (np.array([[0.8189330101, 0.3618667424, -0.1288597137], [0.0329845436, 0.9293118715, 0.0361456387],
                        [0.0482003018, 0.2643662691, 0.6338517070]], dtype=f32),
              np.array([[0.2104542553, 0.7936177850, -0.0040720468], [1.9779984951, -2.4285922050, 0.4505937099],
                        [0.0259040371, 0.7827717662, -0.8086757660]], dtype=f32)),
    "ICaCb": (np.array([[0.37613, 0.70431, -0.05675], [-0.21649, 1.14744, 0.05356], [0.02567, 0.16713, 0.74235]], dtype=f32),
              np.array([[0.4949, 0.5037, 0.0015], [4.2854, -4.5462, 0.2609], [0.3605, 1.1499, -1.5105]], dtype=f32)),
    "ICtCp": (np.array([[0.3592, 0.6976, -0.0358], [-0.1922, 1.1004, 0.0755], [0.0070, 0.0749, 0.8434]], dtype=f32),
              np.array([[0.5, 0.5, 0.0], [1.6137, -3.3234, 1.7097], [4.3781, -4.2455, -0.1325]], dtype=f32)),
    "JzAzBz": (np.array([[0.41478972, 0.579999, 0.0146480], [-0.2015100, 1.120649, 0.0531008],
                         [-0.0166008, 0.264800, 0.6684799]], dtype=f32),
               np.array([[0.5, 0.5, 0.0], [3.524, -4.066708, 0.542708], [0.199076, 1.096799, -1.295875]], dtype=f32)),
}
# per-space (midpoints, scale factors): ycbcr.py:41-42, ycocg.py:41-42,61-62, oklab.py:51-52,
# icacb.py:162-163, ictcp.py:162-163, jzazbz.py:209-210, xyz.py:43-44
NORM = {
    "YCbCr": (np.array([0.5000000037252903, 7.450580596923828e-09, 0.0], dtype=f32),
              np.array([253.99999810755253, 254.000003784895, 254.0], dtype=f32)),
    "YCoCg": (np.array([0.5, 0, 0], dtype=f32), np.array([254, 254, 254], dtype=f32)),
    "YCoCg-R": (np.array([0.5, 0, 0], dtype=f32), np.array([254, 127, 127], dtype=f32)),
    "OKLAB": (np.array([0.4999999, 0.021152213, -0.056563325], dtype=f32), np.array([254.00005, 497.9055, 497.94604], dtype=f32)),
    "ICaCb": (np.array([0.07498085, 0.02180194, -0.018250957], dtype=f32), np.array([1693.7823, 1838.5665, 1330.3855], dtype=f32)),
    "ICtCp": (np.array([0.07497266, -0.0008235276, 0.023989676], dtype=f32), np.array([1693.9674, 1133.9044, 1694.004], dtype=f32)),
    "JzAzBz": (np.array([0.0087900255, 0.00048353244, -0.0020741792], dtype=f32), np.array([14448.194, 7590.505, 5552.201], dtype=f32)),
    "XYZ": (np.array([0.47523502, 0.50000006, 0.544415], dtype=f32), np.array([267.2362, 253.99997, 233.27792], dtype=f32)),
}
# chroma subsampling (jpeg.py:62-147): (rh, rw) for layers 1,2
SUBSAMPLING = {"ICaCb": (1, 4), "ICtCp": (1, 4), "JzAzBz": (2, 2), "OKLAB": (2, 2), "YCbCr": (2, 2),
               "YCoCg": (2, 2), "YCoCg-R": (2, 2)}


class _Tables(C.Structure):
    _fields_ = [("m1", C.c_float * 9), ("m2", C.c_float * 9), ("rgb2xyz", C.c_float * 9),
                ("xyz2rgb", C.c_float * 9), ("lin", C.c_float * 9)]


def _tables(space: str, inverse: bool) -> _Tables:
    t = _Tables()

    def put(dst, m):
        for i, v in enumerate(np.asarray(m, dtype=f32).ravel()):
            dst[i] = float(v)

    put(t.rgb2xyz, _RGB2XYZ)
    put(t.xyz2rgb, _XYZ2RGB)
    if space in _FWD:
        put(t.lin, _FWD[space][1 if inverse else 0])
    elif space in _NL:
        a, b = _NL[space]
        if inverse:
            put(t.m1, np.linalg.inv(b))   # space -> LMS'
            put(t.m2, np.linalg.inv(a))   # LMS -> XYZ
        else:
            put(t.m1, a)
            put(t.m2, b)
    return t


def color_forward(space: str, rgb: np.ndarray) -> np.ndarray:
    """conversion.py:95-124 convert('sRGB', space, x) on an (N,3) float32 array."""
    rgb = np.ascontiguousarray(rgb, dtype=f32)
    out = np.empty_like(rgb)
    t = _tables(space, False)
    lib().ao_color_forward(C.c_int(SPACE_ID[space]), C.byref(t), _p(rgb), _p(out), C.c_size_t(rgb.shape[0]))
    return out


def color_inverse(space: str, x: np.ndarray) -> np.ndarray:
    """conversion.py:95-124 convert(space, 'sRGB', x)."""
    x = np.ascontiguousarray(x, dtype=f32)
    out = np.empty_like(x)
    t = _tables(space, True)
    lib().ao_color_inverse(C.c_int(SPACE_ID[space]), C.byref(t), _p(x), _p(out), C.c_size_t(x.shape[0]))
    return out


def normalize(space: str, channel: int, x: np.ndarray, inverse: bool) -> np.ndarray:
    """conversion.py:126-157 on one channel (call sites jpeg.py:387-390, 452-455)."""
    x = np.ascontiguousarray(x, dtype=f32)
    out = np.empty_like(x)
    mid, sc = NORM[space]
    lib().ao_normalize(_p(x), _p(out), C.c_size_t(x.size), C.c_float(float(mid[channel])), C.c_float(float(sc[channel])),
                       C.c_int(1 if inverse else 0))
    return out


# ----------------------------------------------------------------------------------------------
# resampling / Canny stages / quadtree
# ----------------------------------------------------------------------------------------------
def downsample(layer: np.ndarray, h: int, w: int) -> np.ndarray:
    layer = np.ascontiguousarray(layer, dtype=f32)
    out = np.empty((h, w), dtype=f32)
    lib().ao_downsample_area(_p(layer), C.c_int(layer.shape[0]), C.c_int(layer.shape[1]), _p(out), C.c_int(h), C.c_int(w))
    return out


def resize_linear(layer: np.ndarray, H: int, W: int) -> np.ndarray:
    layer = np.ascontiguousarray(layer, dtype=f32)
    out = np.empty((H, W), dtype=f32)
    lib().ao_resize_linear(_p(layer), C.c_int(layer.shape[0]), C.c_int(layer.shape[1]), _p(out), C.c_int(H), C.c_int(W))
    return out


def cast_u8(layer):
    layer = np.ascontiguousarray(layer, dtype=f32)
    out = np.empty(layer.shape, dtype=np.uint8)
    lib().ao_cast_u8(_p(layer), _p(out), C.c_size_t(layer.size))
    return out


def _u8_stage(fn, src, *extra):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    out = np.empty_like(src)
    fn(_p(src), C.c_int(src.shape[0]), C.c_int(src.shape[1]), _p(out), *extra)
    return out


def clahe(src):
    return _u8_stage(lib().ao_clahe, src)


def gauss3(src):
    return _u8_stage(lib().ao_gauss3, src)


def bilateral5(src, use_fma=True):
    return _u8_stage(lib().ao_bilateral5, src, C.c_int(1 if use_fma else 0))


def percentile_thresholds(src):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    lo, hi = C.c_double(), C.c_double()
    lib().ao_percentile_thresholds(_p(src), C.c_size_t(src.size), C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def canny_u8(src, lo, hi):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    out = np.empty_like(src)
    lib().ao_canny(_p(src), C.c_int(src.shape[0]), C.c_int(src.shape[1]), C.c_double(lo), C.c_double(hi), _p(out))
    return out


def canny(layer: np.ndarray, taps: bool = False):
    """EdgeDetection.canny (edge_detection.py:28-86). Returns float32 {0,1} (and stage taps)."""
    layer = np.ascontiguousarray(layer, dtype=f32)
    h, w = layer.shape
    edge = np.empty((h, w), dtype=np.uint8)
    if taps:
        t = {k: np.empty((h, w), dtype=np.uint8) for k in ("u8", "clahe", "gauss", "bilateral")}
        thr = (C.c_double * 2)()
        lib().ao_canny_pipeline(_p(layer), C.c_int(h), C.c_int(w), _p(edge), _p(t["u8"]), _p(t["clahe"]),
                                _p(t["gauss"]), _p(t["bilateral"]), thr)
        t["thresholds"] = (thr[0], thr[1])
        return edge.astype(f32), t
    lib().ao_canny_pipeline(_p(layer), C.c_int(h), C.c_int(w), _p(edge), None, None, None, None, None)
    return edge.astype(f32)


def root_size(h: int, w: int) -> int:
    return lib().ao_root_size(C.c_int(h), C.c_int(w))


def max_leaves(h: int, w: int, min_size: int) -> int:
    return ((h + min_size - 1) // min_size) * ((w + min_size - 1) // min_size)


def quadtree(edge: np.ndarray, max_size: int, min_size: int):
    """QuadTree(edge, max, min).get_leaves_and_states() (quadtree.py:71-165).
    Returns (leaves int32 (n,3) of x,y,size in DFS order; states uint8 (0 leaf,1 split,2 absent); root)."""
    e = np.ascontiguousarray(edge == 1.0).astype(np.uint8) if edge.dtype != np.uint8 else np.ascontiguousarray(edge)
    h, w = e.shape
    nl_cap = max_leaves(h, w, min_size) + 8
    leaves = np.empty((nl_cap, 3), dtype=np.int32)
    states = np.empty(nl_cap * 3 + 4096, dtype=np.uint8)
    nl, ns = C.c_int(), C.c_int()
    root = lib().ao_quadtree(_p(e), C.c_int(h), C.c_int(w), C.c_int(min_size), C.c_int(max_size), _p(leaves),
                             C.byref(nl), _p(states), C.byref(ns))
    return leaves[:nl.value].copy(), states[:ns.value].copy(), root


def states_to_leaves(states: np.ndarray, root: int, h: int, w: int) -> np.ndarray:
    states = np.ascontiguousarray(states, dtype=np.uint8)
    leaves = np.empty((max(len(states), 1), 3), dtype=np.int32)
    n = lib().ao_states_to_leaves(_p(states), C.c_int(len(states)), C.c_int(root), C.c_int(h), C.c_int(w), _p(leaves))
    return leaves[:n].copy()


# ----------------------------------------------------------------------------------------------
# host tables: quality, quantisation matrices, zigzag
# ----------------------------------------------------------------------------------------------
LUMA_Q = np.array([[16, 11, 10, 16, 24, 40, 51, 61], [12, 12, 14, 19, 26, 58, 60, 55], [14, 13, 16, 24, 40, 57, 69, 56],
                   [14, 17, 22, 29, 51, 87, 80, 62], [18, 22, 37, 56, 68, 109, 103, 77], [24, 35, 55, 64, 81, 104, 113, 92],
                   [49, 64, 78, 87, 103, 121, 120, 101], [72, 92, 95, 98, 112, 100, 103, 99]], dtype=f32)   # jpeg.py:40-49
CHROMA_Q = np.array([[17, 18, 24, 47, 99, 99, 99, 99], [18, 21, 26, 66, 99, 99, 99, 99], [24, 26, 56, 99, 99, 99, 99, 99],
                     [47, 66, 99, 99, 99, 99, 99, 99]] + [[99] * 8] * 4, dtype=f32)                           # jpeg.py:50-59


def block_sizes(bmin: int, bmax: int):
    return [2 ** i for i in range(int(math.log2(bmin)), int(math.log2(bmax)) + 1)]     # jpeg.py:219


def quality_factor(size: int, qrange, brange) -> int:
    """jpeg.py:688-705"""
    bmin, bmax = brange
    qmin, qmax = qrange
    if bmin == bmax:
        return int((qmin + qmax) / 2)
    return int(qmin + (qmax - qmin) * (1 - math.log(size / bmin) / math.log(bmax / bmin)))


def _resize_linear_host(m8: np.ndarray, size: int) -> np.ndarray:
    """cv.resize(8x8 f32 -> size x size, INTER_LINEAR) (jpeg.py:722); exact in f32 because every
    weight is a dyadic fraction and the entries are small integers."""
    return resize_linear(m8, size, size)


def quantization_matrix(base8: np.ndarray, size: int, quality: int) -> np.ndarray:
    """jpeg.py:707-724"""
    scale = 5000 / quality if quality < 50 else 200 - 2 * quality
    scaled = np.floor((scale * base8 + 50) / 100)
    resized = _resize_linear_host(scaled.astype(f32), size)
    resized = np.clip(resized, 1, None)
    return resized.astype(np.int32)


def zigzag(size: int) -> np.ndarray:
    """jpeg.py:726-766"""
    res = np.empty(size * size, dtype=np.int32)
    r = c = 0
    for i in range(size * size):
        res[i] = r * size + c
        if (r + c) % 2 == 0:
            if c == size - 1:
                r += 1
            elif r == 0:
                c += 1
            else:
                r -= 1
                c += 1
        else:
            if r == size - 1:
                c += 1
            elif c == 0:
                r += 1
            else:
                r += 1
                c -= 1
    return res


_ZZ = {}


def zigzag_cached(size):
    if size not in _ZZ:
        _ZZ[size] = zigzag(size)
    return _ZZ[size]


def layer_shapes(H: int, W: int, space: str):
    rh, rw = SUBSAMPLING[space]
    return [(H, W), (H // rh, W // rw), (H // rh, W // rw)]       # jpeg.py:676-686


def qtables(space: str, qrange, brange):
    """quantization_matrix_cache[layer][size] (jpeg.py:228-238)."""
    out = []
    for layer in range(3):
        base = LUMA_Q if layer == 0 else CHROMA_Q
        out.append({s: quantization_matrix(base, s, quality_factor(s, qrange, brange)) for s in block_sizes(*brange)})
    return out


def _qtab_ptrs(tabs: dict):
    arr = (C.c_void_p * 16)()
    keep = []
    for s, t in tabs.items():
        t = np.ascontiguousarray(t, dtype=np.int32)
        keep.append(t)
        arr[int(math.log2(s))] = t.ctypes.data
    return arr, keep


def encode_blocks(layer, space, channel, leaves, tabs, want_dct=False):
    layer = np.ascontiguousarray(layer, dtype=f32)
    leaves = np.ascontiguousarray(leaves, dtype=np.int32)
    n = int((leaves[:, 2].astype(np.int64) ** 2).sum())
    coef = np.empty(n, dtype=np.int32)
    dct = np.empty(n, dtype=f32) if want_dct else None
    arr, keep = _qtab_ptrs(tabs)
    mid, sc = NORM[space]
    lib().ao_encode_blocks(_p(layer), C.c_int(layer.shape[0]), C.c_int(layer.shape[1]), C.c_float(float(mid[channel])),
                           C.c_float(float(sc[channel])), _p(leaves), C.c_int(len(leaves)), arr, _p(coef),
                           _p(dct) if want_dct else None)
    return (coef, dct) if want_dct else coef


def decode_blocks(coef, leaves, tabs, h, w, space, channel):
    coef = np.ascontiguousarray(coef, dtype=np.int32)
    leaves = np.ascontiguousarray(leaves, dtype=np.int32)
    out = np.empty((h, w), dtype=f32)
    arr, keep = _qtab_ptrs(tabs)
    mid, sc = NORM[space]
    lib().ao_decode_blocks(_p(coef), _p(leaves), C.c_int(len(leaves)), arr, C.c_int(h), C.c_int(w),
                           C.c_float(float(mid[channel])), C.c_float(float(sc[channel])), _p(out))
    return out


# ----------------------------------------------------------------------------------------------
# drivers (jpeg.py:240-297) -- hot path only, and with the .ajpg container
# ----------------------------------------------------------------------------------------------
def encode_hot(rgb: np.ndarray, space: str, qrange, brange, taps: bool = False):
    """Jpeg.compress minus _entropy_encode.  rgb: (H,W,3) float32.  Returns per-layer dicts with
    leaves (n,3), states, root, coef (int32, leaf order, row-major blocks)."""
    H, W, _ = rgb.shape
    conv = color_forward(space, rgb.reshape(-1, 3))
    planes = [np.ascontiguousarray(conv[:, i].reshape(H, W)) for i in range(3)]
    shapes = layer_shapes(H, W, space)
    tabs = qtables(space, qrange, brange)
    layers = []
    for i in range(3):
        lay = downsample(planes[i], *shapes[i])
        if taps:
            edge, t = canny(lay, taps=True)
        else:
            edge, t = canny(lay), None
        leaves, states, root = quadtree(edge, brange[1], brange[0])
        if taps:
            coef, dct = encode_blocks(lay, space, i, leaves, tabs[i], want_dct=True)
        else:
            coef, dct = encode_blocks(lay, space, i, leaves, tabs[i]), None
        d = dict(layer=lay, edge=edge, leaves=leaves, states=states, root=root, coef=coef)
        if taps:
            d.update(taps=t, dct=dct)
        layers.append(d)
    return layers


def decode_hot(layers, H: int, W: int, space: str, qrange, brange) -> np.ndarray:
    """Jpeg.decompress minus _entropy_decode.  layers: per-layer dict(leaves, coef). Returns (H,W,3) f32."""
    shapes = layer_shapes(H, W, space)
    tabs = qtables(space, qrange, brange)
    ups = []
    for i in range(3):
        lay = decode_blocks(layers[i]["coef"], layers[i]["leaves"], tabs[i], shapes[i][0], shapes[i][1], space, i)
        ups.append(resize_linear(lay, H, W))
    x = np.stack(ups, axis=2).reshape(-1, 3)
    return color_inverse(space, x).reshape(H, W, 3)


def pack_states(states: np.ndarray) -> bytes:
    """2 bits per state, MSB first, zero padded (jpeg.py:563-571)."""
    s = np.asarray(states, dtype=np.uint8)
    pad = (-len(s)) % 4
    s4 = np.concatenate([s, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
    return ((s4[:, 0] << 6) | (s4[:, 1] << 4) | (s4[:, 2] << 2) | s4[:, 3]).astype(np.uint8).tobytes()


def unpack_states(buf: bytes, n_bits: int) -> np.ndarray:
    b = np.frombuffer(buf, dtype=np.uint8)
    s = np.stack([(b >> 6) & 3, (b >> 4) & 3, (b >> 2) & 3, b & 3], axis=1).reshape(-1)
    return s[: n_bits // 2].astype(np.uint8)


def zigzag_stream(coef: np.ndarray, leaves: np.ndarray, inverse: bool = False) -> np.ndarray:
    """per-block zigzag gather (jpeg.py:579-585) / scatter (jpeg.py:664-672) on the concatenated stream."""
    out = np.empty_like(coef)
    sizes = leaves[:, 2].astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(sizes * sizes)])
    for s in np.unique(sizes):
        idx = np.nonzero(sizes == s)[0]
        zz = zigzag_cached(int(s)).astype(np.int64)
        base = offs[idx][:, None]
        if inverse:
            out[(base + zz[None, :]).ravel()] = coef[(base + np.arange(s * s)[None, :]).ravel()]
        else:
            out[(base + np.arange(s * s)[None, :]).ravel()] = coef[(base + zz[None, :]).ravel()]
    return out


def compress(rgb: np.ndarray, space="YCoCg", qrange=(40, 80), brange=(4, 64), extension=None) -> bytes:
    """Jpeg.compress (jpeg.py:240-272) producing the .ajpg byte stream (jpeg.py:531-597)."""
    H, W, _ = rgb.shape
    layers = encode_hot(rgb, space, qrange, brange)
    out = BytesIO()
    meta = {"height": H, "width": W, "num_layers": 3, "color_space": space, "quality_min": qrange[0],
            "quality_max": qrange[1], "block_size_min": brange[0], "block_size_max": brange[1], "extension": extension}
    mb = json.dumps(meta).encode("utf-8")
    out.write(len(mb).to_bytes(4, "big"))
    out.write(mb)
    for L in layers:
        out.write((2 * len(L["states"])).to_bytes(4, "big"))
        out.write(int(L["root"]).to_bytes(4, "big"))
        out.write(pack_states(L["states"]))
        z = zlib.compress(zigzag_stream(L["coef"], L["leaves"]).tobytes(), level=9)
        out.write(len(z).to_bytes(4, "big"))
        out.write(z)
    return out.getvalue()


def decompress(buf: bytes) -> np.ndarray:
    """Jpeg.decompress (jpeg.py:274-297, 599-674). Returns (H,W,3) float32."""
    s = BytesIO(buf)
    ml = int.from_bytes(s.read(4), "big")
    meta = json.loads(s.read(ml).decode("utf-8"))
    H, W, space = meta["height"], meta["width"], meta["color_space"]
    qrange = (meta["quality_min"], meta["quality_max"])
    brange = (meta["block_size_min"], meta["block_size_max"])
    shapes = layer_shapes(H, W, space)
    layers = []
    for i in range(meta["num_layers"]):
        nb = int.from_bytes(s.read(4), "big")
        root = int.from_bytes(s.read(4), "big")
        states = unpack_states(s.read((nb + 7) // 8), nb)
        leaves = states_to_leaves(states, root, shapes[i][0], shapes[i][1])
        zl = int.from_bytes(s.read(4), "big")
        coef = np.frombuffer(zlib.decompress(s.read(zl)), dtype=np.int32)
        layers.append(dict(leaves=leaves, coef=zigzag_stream(coef, leaves, inverse=True)))
    return decode_hot(layers, H, W, space, qrange, brange)
