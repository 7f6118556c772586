"""Quality metrics used to report parity deltas (PSNR, SSIM on luma).  The reference delegates to
piq / lpips (src/image/evaluation_metrics.py:31-139), neither of which is on the hot path nor
available offline; MS-SSIM and LPIPS are therefore not provided and raise NotImplementedError."""
from __future__ import annotations

import numpy as np

from .image import Image


def _gray(x: np.ndarray) -> np.ndarray:
    return (0.299 * x[..., 0] + 0.587 * x[..., 1] + 0.114 * x[..., 2]).astype(np.float64)


class EvaluationMetrics:
    @staticmethod
    def psnr(original: Image, compressed: Image) -> float:
        a, b = original.data.astype(np.float64), compressed.data.astype(np.float64)
        mse = float(np.mean((a - b) ** 2))
        return float("inf") if mse == 0 else 10.0 * np.log10(1.0 / mse)

    @staticmethod
    def ssim(original: Image, compressed: Image) -> float:
        """Gaussian-window SSIM (11x11, sigma 1.5, K1=0.01, K2=0.03) on the luma plane, data range 1."""
        from scipy.ndimage import gaussian_filter
        x, y = _gray(original.data), _gray(compressed.data)
        f = lambda v: gaussian_filter(v, 1.5, truncate=3.5)
        mx, my = f(x), f(y)
        vx, vy, cxy = f(x * x) - mx * mx, f(y * y) - my * my, f(x * y) - mx * my
        c1, c2 = 0.01 ** 2, 0.03 ** 2
        s = ((2 * mx * my + c1) * (2 * cxy + c2)) / ((mx * mx + my * my + c1) * (vx + vy + c2))
        return float(s.mean())

    @staticmethod
    def ms_ssim(original: Image, compressed: Image) -> float:
        raise NotImplementedError("MS-SSIM (piq) is outside the hot path and not available offline")

    @staticmethod
    def lpips(original: Image, compressed: Image) -> float:
        raise NotImplementedError("LPIPS (lpips) is outside the hot path and not available offline")
