"""Drop-in for the reference's ``image.EvaluationMetrics`` (src/image/evaluation_metrics.py:31-139).

Same call surface -- ``EvaluationMetrics(original, compressed).psnr() / .ssim() / .ms_ssim() / .lpips()`` -- with PSNR,
SSIM and MS-SSIM computed by torch ops on the GPU when one is present (quality metrics are how the
reference's users judge the codec).  The reference delegates to ``piq`` (psnr, ssim, multi_scale_ssim) and ``lpips``;
neither wheel is available offline, so the three piq metrics restate piq's published algorithm (11x11 Gaussian window,
sigma 1.5, K1 = 0.01, K2 = 0.03, piq's down-sampling rule, five MS-SSIM scales with the standard weights) and are
**unpinned** against piq itself.  LPIPS needs the AlexNet weights of the ``lpips`` package and raises NotImplementedError.
Not on the hot path.
"""
from __future__ import annotations

from typing import Union

import numpy as np
import torch
import torch.nn.functional as F

from .image import Image

_MS_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def _device() -> torch.device:
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def _gaussian_kernel(size: int, sigma: float, channels: int, device) -> torch.Tensor:
    coords = torch.arange(size, dtype=torch.float32, device=device) - (size - 1) / 2.0
    g = coords ** 2
    g = (-(g.unsqueeze(0) + g.unsqueeze(1)) / (2 * sigma ** 2)).exp()
    g = g / g.sum()
    return g.unsqueeze(0).unsqueeze(0).repeat(channels, 1, 1, 1)


def _ssim_per_channel(x, y, kernel, k1: float, k2: float):
    """SSIM and contrast-sensitivity means per channel (inputs already scaled to data range 1)."""
    c1, c2 = k1 ** 2, k2 ** 2
    ch = x.size(1)
    mu_x, mu_y = F.conv2d(x, kernel, groups=ch), F.conv2d(y, kernel, groups=ch)
    mu_xx, mu_yy, mu_xy = mu_x ** 2, mu_y ** 2, mu_x * mu_y
    s_xx = F.conv2d(x ** 2, kernel, groups=ch) - mu_xx
    s_yy = F.conv2d(y ** 2, kernel, groups=ch) - mu_yy
    s_xy = F.conv2d(x * y, kernel, groups=ch) - mu_xy
    cs = (2.0 * s_xy + c2) / (s_xx + s_yy + c2)
    ss = (2.0 * mu_xy + c1) / (mu_xx + mu_yy + c1) * cs
    return ss.mean(dim=(-1, -2)), cs.mean(dim=(-1, -2))


def rgb_to_gray_u8(rgb_u8: np.ndarray) -> np.ndarray:
    """cv.cvtColor(x, cv.COLOR_RGB2GRAY) on 8-bit pixels: OpenCV's 15-bit fixed point,
    (9798 R + 19235 G + 3735 B + 16384) >> 15 (checked bit-exact against cv2 4.13 in tests/test_host_shim.py)."""
    x = rgb_u8.astype(np.int32)
    return ((9798 * x[..., 0] + 19235 * x[..., 1] + 3735 * x[..., 2] + 16384) >> 15).astype(np.uint8)


class EvaluationMetrics:
    """A collection of image quality assessment metrics (evaluation_metrics.py:31)."""

    def __init__(self, original_image: Image, compressed_image: Image) -> None:
        self.original_image = original_image
        self.compressed_image = compressed_image

    def psnr(self) -> torch.Tensor:
        """piq.psnr(x, y, data_range=1.0): -10 log10(mean squared error + 1e-8), evaluation_metrics.py:50-61."""
        x, y = self._image_to_tensor(self.original_image).to(_device()), self._image_to_tensor(self.compressed_image).to(_device())
        mse = torch.mean((x.float() - y.float()) ** 2, dim=[1, 2, 3])
        return (-10.0 * torch.log10(mse + 1e-8)).mean().cpu()

    def ssim(self) -> torch.Tensor:
        """piq.ssim on the 8-bit luma (cv.COLOR_RGB2GRAY) with data_range=255, evaluation_metrics.py:63-76."""
        gx = rgb_to_gray_u8(self.original_image.get_uint8())
        gy = rgb_to_gray_u8(self.compressed_image.get_uint8())
        x = self._image_to_tensor(gx).to(_device()).float() / 255.0
        y = self._image_to_tensor(gy).to(_device()).float() / 255.0
        f = max(1, round(min(x.shape[-2:]) / 256))
        if f > 1:
            x, y = F.avg_pool2d(x, kernel_size=f), F.avg_pool2d(y, kernel_size=f)
        kernel = _gaussian_kernel(11, 1.5, x.size(1), x.device)
        ss, _ = _ssim_per_channel(x, y, kernel, 0.01, 0.03)
        return ss.mean(1).mean().cpu()

    def ms_ssim(self) -> torch.Tensor:
        """piq.multi_scale_ssim(x, y, data_range=1.0), evaluation_metrics.py:78-89: five dyadic scales."""
        x, y = self._image_to_tensor(self.original_image).to(_device()).float(), self._image_to_tensor(self.compressed_image).to(_device()).float()
        levels = len(_MS_WEIGHTS)
        min_size = (11 - 1) * 2 ** (levels - 1) + 1
        if x.size(-1) < min_size or x.size(-2) < min_size:
            raise ValueError(f"Invalid size of the input images, expected at least {min_size}x{min_size}.")
        kernel = _gaussian_kernel(11, 1.5, x.size(1), x.device)
        weights = torch.tensor(_MS_WEIGHTS, device=x.device)
        mcs = []
        ss = None
        for it in range(levels):
            if it > 0:
                pad = max(x.shape[2] % 2, x.shape[3] % 2)
                x = F.avg_pool2d(F.pad(x, pad=[pad, 0, pad, 0], mode="replicate"), kernel_size=2, padding=0)
                y = F.avg_pool2d(F.pad(y, pad=[pad, 0, pad, 0], mode="replicate"), kernel_size=2, padding=0)
            ss, cs = _ssim_per_channel(x, y, kernel, 0.01, 0.03)
            mcs.append(cs)
        stack = torch.relu(torch.stack(mcs[:-1] + [ss], dim=0))
        val = torch.prod(stack ** weights.view(-1, 1, 1), dim=0).mean(1)
        return val.mean().cpu()

    def lpips(self) -> float:
        raise NotImplementedError("LPIPS needs the AlexNet weights of the `lpips` package, which is not available offline")

    @staticmethod
    def _image_to_tensor(image: Union[Image, np.ndarray]) -> torch.Tensor:
        """(H,W) -> (1,1,H,W); (H,W,C) -> (1,C,H,W)  (evaluation_metrics.py:112-139)."""
        if isinstance(image, Image):
            data = image.data
        elif isinstance(image, np.ndarray):
            data = image
        else:
            raise TypeError(f"Expected Image or numpy.ndarray, got {type(image)}")
        tensor = torch.from_numpy(np.ascontiguousarray(data))
        if tensor.dim() == 2:
            return tensor.unsqueeze(0).unsqueeze(0)
        if tensor.dim() == 3:
            return tensor.permute(2, 0, 1).unsqueeze(0)
        raise ValueError(f"Unexpected shape: {tensor.shape}")
