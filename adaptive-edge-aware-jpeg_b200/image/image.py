"""Drop-in for the reference's ``image.Image`` (src/image/image.py:26-149): a float32 HWC [0,1]
container.  It is the boundary type of Jpeg.compress / decompress; file I/O stays on the host."""
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
    def __init__(self, data: np.ndarray, shape: Tuple[int, ...], extension: Optional[str]) -> None:
        self.data = data
        self.original_shape = shape
        self.extension = extension

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
        img = _imread(path).astype(np.float32) / 255.0
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
        return Image.from_array(self.data.copy(), self.original_shape, self.extension)

    def save(self, path: str) -> None:
        _imwrite(path, (self.data * 255).astype(np.uint8))

    def get_flattened(self) -> np.ndarray:
        return self.data.reshape(-1, self.original_shape[-1])

    def get_uint8(self) -> np.ndarray:
        return (self.data * 255).astype(np.uint8)       # truncating view (image.py:127)

    def reshape(self, shape: Tuple[int, ...]) -> "Image":
        self.data = self.data.reshape(shape)
        return self

    def __str__(self) -> str:
        return self.data.__str__()
