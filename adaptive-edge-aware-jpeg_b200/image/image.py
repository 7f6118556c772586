"""Drop-in for the reference's ``image.Image`` (src/image/image.py:26-149): a float32 HWC [0,1]
container.  It is the boundary type of Jpeg.compress / decompress; file I/O stays on the host.

Addition: an image that came from 8-bit pixels (``Image.load`` of an 8-bit file, ``Image.from_uint8``) remembers them, and
``Jpeg.compress`` then uploads the bytes and lets the GPU do ``astype(float32) / 255.0`` (image.py:84) -- the same floats at a
quarter of the PCIe traffic.  ``data`` is materialised lazily; handing it out (or assigning it) drops the 8-bit shortcut, since
the caller may then change the floats."""
from __future__ import annotations

import os
from typing import Optional, Tuple, Type

import numpy as np


def _imread(path):
    try:
        import imageio.v3 as iio
        return iio.imread(path)
    except ImportError:
        from PIL import Image as PILImage
        return np.asarray(PILImage.open(path))


def _imwrite(path, arr):
    try:
        import imageio.v3 as iio
        iio.imwrite(path, arr)
    except ImportError:
        from PIL import Image as PILImage
        PILImage.fromarray(arr).save(path)


class Image:
    def __init__(self, data: Optional[np.ndarray], shape: Tuple[int, ...], extension: Optional[str]) -> None:
        self._data = data
        self._u8: Optional[np.ndarray] = None           # 8-bit source pixels [H,W,3], valid while `data` has not been handed out
        self.original_shape = shape
        self.extension = extension

    @property
    def data(self) -> np.ndarray:
        if self._data is None:
            self._data = self._u8.astype(np.float32) / 255.0
        self._u8 = None
        return self._data

    @data.setter
    def data(self, value: np.ndarray) -> None:
        self._data = value
        self._u8 = None

    def uint8_source(self) -> Optional[np.ndarray]:
        """the untouched 8-bit pixels this image was made from, or None"""
        return self._u8

    @classmethod
    def from_uint8(cls: Type["Image"], pixels: np.ndarray, extension: Optional[str] = None) -> "Image":
        """8-bit RGB pixels [H,W,3]; equivalent to Image(pixels.astype(float32) / 255.0, ...) (what load() produces)."""
        pixels = np.ascontiguousarray(pixels)
        if pixels.dtype != np.uint8 or pixels.ndim != 3 or pixels.shape[2] != 3:
            raise ValueError("from_uint8 expects uint8 pixels of shape [H,W,3]")
        img = cls(None, pixels.shape, extension)
        img._u8 = pixels
        return img

    @classmethod
    def from_array(cls: Type["Image"], data: np.ndarray, shape: Optional[Tuple[int, ...]] = None,
                   extension: Optional[str] = None) -> "Image":
        if shape is None:
            shape = data.shape
        img = cls(data, shape, extension)
        img.reshape(shape)
        return img

    @classmethod
    def load(cls: Type["Image"], path: str) -> "Image":
        extension = os.path.splitext(path)[1]
        raw = _imread(path)
        if raw.dtype == np.uint8 and raw.ndim in (2, 3) and (raw.ndim == 2 or raw.shape[2] in (3, 4)):
            rgb = np.stack((raw,) * 3, axis=-1) if raw.ndim == 2 else raw[:, :, :3]
            return cls.from_uint8(rgb, extension)
        img = raw.astype(np.float32) / 255.0
        if img.ndim == 2:
            img = np.stack((img,) * 3, axis=-1)
        elif img.ndim == 3 and img.shape[2] == 3:
            pass
        elif img.ndim == 3 and img.shape[2] == 4:
            img = img[:, :, :3]
        else:
            raise ValueError(f"Unsupported image format: {img.shape}")
        return cls(img, img.shape, extension)

    def copy(self) -> "Image":
        if self._data is None:
            return Image.from_uint8(self._u8.copy(), self.extension)
        return Image.from_array(self._data.copy(), self.original_shape, self.extension)

    def save(self, path: str) -> None:
        _imwrite(path, self.get_uint8())

    def get_flattened(self) -> np.ndarray:
        return self.data.reshape(-1, self.original_shape[-1])

    @property
    def ndim(self) -> int:
        return len(self._u8.shape) if self._data is None else self._data.ndim

    def get_uint8(self) -> np.ndarray:
        if self._data is None:
            return ((self._u8.astype(np.float32) / 255.0) * 255).astype(np.uint8).reshape(self.original_shape)
        return (self._data * 255).astype(np.uint8)      # truncating view (image.py:127)

    def reshape(self, shape: Tuple[int, ...]) -> "Image":
        if self._data is None:
            self._u8 = self._u8.reshape(shape)
        else:
            self._data = self._data.reshape(shape)
        return self

    def __str__(self) -> str:
        return self.data.__str__()
