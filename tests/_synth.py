"""Synthetic Z500-anomaly-like cubes for parity tests (CPU form).

White N(0,1) float32 noise smoothed with a separable Gaussian (nearest, nearest, wrap-in-longitude), scaled to a
global standard deviation of 100 so that threshold 160 is ~1.6 sigma (3.5-5 % coverage).
"""
import numpy as np
from scipy import ndimage


def synth_cube(seed, T, H, W, sigma):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((T, H, W), np.float32)
    a = ndimage.gaussian_filter(a, sigma, mode=('nearest', 'nearest', 'wrap'))
    a *= 100 / a.std()
    return np.ascontiguousarray(a, dtype=np.float32)


def regular_grid(H, W):
    """lat 90..-90 and lon 0..360 on a regular grid whose float32 coordinate differences are all identical
    (required by the reference's set_up, contrack.py:352-370)."""
    if H > 1 and (180 * 64) % (H - 1) == 0:
        lat = (90 - np.arange(H) * (180.0 / (H - 1))).astype(np.float32)
    else:                       # keep a regular float32 grid that does not reach the poles exactly
        step = np.float32(np.floor(180.0 / H * 4) / 4) if H <= 720 else np.float32(0.125)
        lat = (np.float32(step * (H - 1) / 2) - np.arange(H, dtype=np.float32) * step).astype(np.float32)
    dlon = 360.0 / W
    if (360 * 64) % W != 0:
        dlon = np.floor(dlon * 8) / 8
    lon = (np.arange(W) * dlon).astype(np.float32)
    return lat, lon
