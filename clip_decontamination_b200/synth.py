"""Seeded synthetic remote-sensing-like scenes (uint8 BGR HWC), SURVEY.md §8(d).

`randn` images are degenerate for this path (every pixel gets the same class), so parity tests
and the benchmark use Voronoi scenes: 20-60 regions per tile, each with its own mean colour, a
band-limited texture (sum of 3 random sinusoids) and sigma=8 noise.  Pure numpy, deterministic
for a given (seed, H, W).
"""
import numpy as np

MEAN = np.array([122.771, 116.746, 104.094], dtype=np.float32)   # RGB, segmentor.py:64-67
STD = np.array([68.501, 66.632, 70.323], dtype=np.float32)


def voronoi_scene(H: int, W: int, seed: int = 2) -> np.ndarray:
    """uint8 [H, W, 3] in BGR channel order (what mmseg's LoadImageFromFile hands over)."""
    rng = np.random.RandomState(seed)
    n = rng.randint(20, 61)
    cy, cx = rng.uniform(0, H, n), rng.uniform(0, W, n)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    d = (yy[None] - cy[:, None, None].astype(np.float32)) ** 2 + (xx[None] - cx[:, None, None].astype(np.float32)) ** 2
    region = d.argmin(0)
    colors = rng.uniform(20, 235, (n, 3)).astype(np.float32)
    img = colors[region]
    for _ in range(3):
        fy, fx = rng.uniform(0.01, 0.15, 2)
        ph = rng.uniform(0, 2 * np.pi)
        amp = rng.uniform(4, 14, 3).astype(np.float32)
        img += amp[None, None] * np.sin(fy * yy + fx * xx + ph)[..., None]
    img += rng.normal(0, 8, img.shape).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def preprocess(img_bgr_u8: np.ndarray) -> np.ndarray:
    """mmseg SegDataPreProcessor arithmetic (bgr_to_rgb, (x-mean)/std): uint8 HWC BGR ->
    float32 [3, H, W] RGB.  Restated from mmsegmentation 1.2.2 (external to the reference)."""
    rgb = img_bgr_u8[..., ::-1].astype(np.float32)
    out = (rgb - MEAN[None, None]) / STD[None, None]
    return np.ascontiguousarray(out.transpose(2, 0, 1))


def synthetic_labels(H: int, W: int, num_classes: int, seed: int = 3) -> np.ndarray:
    """uint8 [H, W] ground-truth-like map (Voronoi regions -> classes, a few 255 = ignore)."""
    rng = np.random.RandomState(seed)
    n = rng.randint(10, 30)
    cy, cx = rng.uniform(0, H, n), rng.uniform(0, W, n)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    d = (yy[None] - cy[:, None, None]) ** 2 + (xx[None] - cx[:, None, None]) ** 2
    lab = rng.randint(0, num_classes, n)[d.argmin(0)].astype(np.uint8)
    lab[rng.uniform(size=(H, W)) < 0.01] = 255
    return lab
