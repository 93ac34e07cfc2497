"""Observation noise with the reference's seeds (utils/utils_noise.py:3-41).

These draws must be bit-exact, so they stay on the host with numpy's legacy global generator
(np.random.seed(1234) before every draw, as the reference does) — never re-implemented on the GPU.
Set-up cost only; not on the per-iteration path.
"""
import numpy as np


def add_gaussian_noise(img, noise_level, random_sampling_op):
    """utils_noise.py:33-36"""
    np.random.seed(1234)
    return img + random_sampling_op(noise_level * np.random.randn(*img.shape))


def apply_poisson_noise(img, alpha):
    """utils_noise.py:38-41 (returns int64 counts)"""
    np.random.seed(1234)
    return np.random.poisson(img * alpha)


def add_salt_and_pepper_noise(img, noise_level, random_sampling_op):
    """utils_noise.py:3-31.  Same stream of 2*noise_cnt (x, y) draws, both from randint(0, shape[-2]);
    a draw is kept when the pixel is observed and unused (the reference's retry decrement is dead
    code, so fewer than 2*noise_cnt points may be kept); first noise_cnt kept -> 0, the rest -> 1.
    The reference's O(n^2) membership scan is replaced by a set."""
    H = img.shape[-2]
    noise_cnt = int(img.shape[-2] * img.shape[-1] * noise_level / 2)
    target = random_sampling_op(np.ones([img.shape[-2], img.shape[-1]]))
    np.random.seed(1234)
    kept_x, kept_y, used = [], [], set()
    for _ in range(noise_cnt * 2):
        x = np.random.randint(0, H)
        y = np.random.randint(0, H)
        if target[x][y] == 1 and (x * H + y) not in used:
            used.add(x * H + y)
            kept_x.append(x)
            kept_y.append(y)
    kx, ky = np.asarray(kept_x, dtype=np.int64), np.asarray(kept_y, dtype=np.int64)
    out = np.copy(img)
    for p in ([out] if out.ndim == 2 else [out[i] for i in range(3)]):
        p[(kx[:noise_cnt], ky[:noise_cnt])] = 0
        p[(kx[noise_cnt:], ky[noise_cnt:])] = 1
    return out
