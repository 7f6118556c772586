"""Drop-in for the reference's ``color`` package (src/color/conversion.py:86-157): same names,
arguments, return types and exceptions; the arithmetic runs in the fused sm_100a colour kernel
(csrc/color.cu) through the C ABI (aeaj_color_forward / aeaj_color_inverse / aeaj_normalize)."""
from __future__ import annotations

import numpy as np

from aeaj import tables
from aeaj.codec import get_stages

_SPACES = ("sRGB",) + tuple(tables.SPACE_ID.keys())


def get_color_spaces() -> list[str]:
    """Available colour spaces (conversion.py:86-93: everything except sRGB and XYZ)."""
    return [s for s in tables.SPACE_ID if s != "XYZ"]


def _check(data):
    if not isinstance(data, np.ndarray):
        raise TypeError("Data input must be a numpy array.")
    if data.ndim != 2 or data.shape[1] != 3:
        raise ValueError("Data input array must be a 2D with 3 channels.")


def convert(from_space: str, to_space: str, data: np.ndarray) -> np.ndarray:
    """conversion.py:95-124.  One of the two spaces must be sRGB."""
    _check(data)
    if from_space not in _SPACES or to_space not in _SPACES:
        raise ValueError("Invalid color space. Please check the available color spaces.")
    if from_space != "sRGB" and to_space != "sRGB":
        raise ValueError("One of the color spaces must be sRGB.")
    if from_space == "sRGB" and to_space == "sRGB":
        return None                                   # the reference indexes a None function table here
    if from_space == "sRGB":
        return get_stages().color(to_space, data, inverse=False)
    return get_stages().color(from_space, data, inverse=True)


def apply_normalization(color_space: str, data: np.ndarray, inverse: bool) -> np.ndarray:
    """conversion.py:126-157: (d - midpoint) * scale, or d / scale + midpoint, per channel."""
    _check(data)
    if color_space not in _SPACES or color_space == "sRGB":
        raise ValueError("Invalid color space. Please check the available color spaces.")
    st = get_stages()
    cols = [st.normalize(color_space, ch, np.ascontiguousarray(data[:, ch]), inverse) for ch in range(3)]
    return np.stack(cols, axis=1)
