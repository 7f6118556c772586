"""Seeded synthetic image generator -- the measurement definition for the C2/C4/C5 workloads
(SURVEY.md Appendix D).  uint8-quantised RGB stored as float32 k/255, which is what the reference's
Image.load produces (image.py:80).  Shared by tests/, bench.py and tests/golden/make_golden.py."""
import numpy as np


def synth(H: int, W: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    img = np.stack([0.45 + 0.25 * np.sin(xx / (W / 3.1) + yy / (H / 1.7)),
                    0.5 + 0.25 * np.cos(xx / (W / 2.3) - yy / (H / 2.9)),
                    0.4 + 0.2 * np.sin((xx + yy) / (W / 4.7))], -1).astype(np.float32)
    for _ in range(int(H * W / 20000)):
        cx = int(rng.random() ** 1.5 * W * 0.6)
        cy = int(rng.random() * H)
        r = int(4 + rng.random() * min(H, W) / 18)
        col = rng.random(3).astype(np.float32)
        if rng.random() < 0.5:                                   # filled circle
            y0, y1, x0, x1 = max(cy - r, 0), min(cy + r + 1, H), max(cx - r, 0), min(cx + r + 1, W)
            m = (yy[y0:y1, x0:x1] - cy) ** 2 + (xx[y0:y1, x0:x1] - cx) ** 2 <= r * r
            img[y0:y1, x0:x1][m] = col
        else:                                                    # filled rectangle
            img[max(cy - r // 2, 0):min(cy + r // 2 + 1, H), max(cx - r, 0):min(cx + r + 1, W)] = col
    mask = ((xx > 0.75 * W) & (yy < 0.5 * H))[..., None]         # textured quadrant
    img = img + (rng.random((H, W, 1)).astype(np.float32) * 0.35 - 0.175) * mask
    return (np.round(np.clip(img, 0, 1) * 255) / 255).astype(np.float32)
