"""Host-side metrics with the reference's signatures (utils/utils_eval.py:4-12).

Inside the resident loop PSNR and c[i] come from the fused dual kernel's partial sums
(engine.metrics_from_traces); these functions serve the driver (PSNR/SSIM of the observation,
final SSIM).  eval_ssim restates scikit-image's structural_similarity defaults (uniform 7-tap
window, K1=.01, K2=.03, sample covariance, channel_axis=0) — scikit-image is not a dependency
here, and its parity is unpinned.
"""
import numpy as np


def eval_psnr(im1, im2):
    mse = np.mean((np.asarray(im1).astype(float) - np.asarray(im2).astype(float)) ** 2)
    return 10 * np.log10(1.0 / mse)


def _box_valid(a, win, axes):
    out = a
    for ax in axes:
        cs = np.cumsum(np.insert(out, 0, 0.0, axis=ax), axis=ax)
        n = out.shape[ax]
        out = (np.take(cs, np.arange(win, n + 1), axis=ax) - np.take(cs, np.arange(0, n - win + 1), axis=ax)) / win
    return out


def _ssim_channel(a, b, data_range, win=7):
    axes = tuple(range(a.ndim))
    npix = win ** a.ndim
    cov_norm = npix / (npix - 1.0)
    ux, uy = _box_valid(a, win, axes), _box_valid(b, win, axes)
    vx = cov_norm * (_box_valid(a * a, win, axes) - ux * ux)
    vy = cov_norm * (_box_valid(b * b, win, axes) - uy * uy)
    vxy = cov_norm * (_box_valid(a * b, win, axes) - ux * uy)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
    return float(np.mean(s))


def eval_ssim(im1, im2):
    im1, im2 = np.asarray(im1, dtype=np.float64), np.asarray(im2, dtype=np.float64)
    data_range = float(im2.max() - im2.min())
    return float(np.mean([_ssim_channel(im1[c], im2[c], data_range) for c in range(im1.shape[0])]))
