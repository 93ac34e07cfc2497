"""Reference-facing operator API (mirror of the reference's operators.py) on the B200 kernels.

  get_observation_operators(operator, path_kernel, r) -> (phi, adj_phi)     operators.py:60-79
  proj_l2_ball / proj_l1_ball / prox_GKL / grad_x_l2 / grad_s_l2 / denoise  operators.py:81-115

phi / adj_phi are callable objects (numpy in -> numpy out, like the reference closures) that also
carry (kind, blur kernel, r); iteration.test_iter recognises them and runs the whole loop resident
on the GPU.  Foreign callables are rejected there with TypeError — there is no CPU path.
"""
from __future__ import annotations

import functools
import os

import numpy as np

from .engine import Engine, l1_ball_radius, l2_ball_radius

_KINDS = ("blur", "random_sampling", "Id")


def load_blur_kernel(path_kernel):
    """h = loadmat(path_kernel)['blur'] (operators.py:77-78); also accepts an ndarray or .npy/.npz."""
    if isinstance(path_kernel, np.ndarray):
        return np.asarray(path_kernel, dtype=np.float64)
    p = str(path_kernel)
    if not os.path.exists(p) and os.path.exists(os.path.splitext(p)[0] + ".npy"):
        p = os.path.splitext(p)[0] + ".npy"          # converted kernel next to a missing .mat
    if p.endswith(".npy"):
        return np.load(p).astype(np.float64)
    if p.endswith(".npz"):
        z = np.load(p)
        key = "blur_1" if "blur_1" in z else ("blur" if "blur" in z else z.files[0])
        return z[key].astype(np.float64)
    import scipy.io
    return np.array(scipy.io.loadmat(p)["blur"], dtype=np.float64)


@functools.lru_cache(maxsize=32)
def sampling_mask(H: int, W: int, r: float) -> np.ndarray:
    """Keep-mask of get_random_sampling_operator (operators.py:40-58).  Integer RNG work stays on the
    host with numpy's legacy MT19937 so that it is bit-exact; the mask is uploaded once."""
    degraded_cnt = round(H * W * (1 - r))
    q = np.random.RandomState(seed=1234).permutation(H * W)[:degraded_cnt]
    m = np.ones(H * W, dtype=np.uint8)
    m[q] = 0
    m = m.reshape(H, W)
    m.setflags(write=False)
    return m


_engines: dict = {}


_MAX_CACHED_ENGINES = 16


def _cache_put(key, engine):
    """Bounded cache (oldest first out, closed on eviction) shared by the operator engines and the flat prox engines: every
    Engine owns device memory proportional to its image size."""
    while len(_engines) >= _MAX_CACHED_ENGINES:
        _engines.pop(next(iter(_engines))).close()
    _engines[key] = engine


def _engine_for(kind: str, C: int, H: int, W: int, h, r):
    key = (kind, C, H, W, None if h is None else h.tobytes(), r if kind == "random_sampling" else None)
    e = _engines.get(key)
    if e is None:
        e = Engine(1, C, H, W, method="A", deg_op=kind, max_iter=1)
        if kind == "blur":
            e.set_blur_kernel(h)
        elif kind == "random_sampling":
            e.set_mask(sampling_mask(H, W, r))
        _cache_put(key, e)
    return e


class ObservationOperator:
    """phi (adjoint=False) or adj_phi (adjoint=True) of get_observation_operators."""

    def __init__(self, kind: str, adjoint: bool, h, r: float):
        self.kind, self.adjoint, self.h, self.r = kind, bool(adjoint), h, float(r)

    def __call__(self, x):
        if self.kind not in _KINDS:
            return None                       # the reference's closures fall through and return None
        if self.kind == "Id":
            return x
        x = np.asarray(x)
        if x.ndim == 2:
            C, (H, W) = 1, x.shape
        elif x.ndim == 3:
            C, H, W = x.shape
            if C not in (1, 3):
                raise ValueError("colour images must have 3 channels (operators.py:53)")
        else:
            raise ValueError("expected (H,W) or (C,H,W)")
        e = _engine_for(self.kind, C, H, W, self.h if self.kind == "blur" else None, self.r)
        xd = e.to_device(x)
        out = e.phi_adj(xd) if self.adjoint else e.phi(xd)
        y = out.cpu().numpy().reshape(x.shape)
        if self.kind == "random_sampling" and x.ndim == 2:
            return y.astype(x.dtype, copy=False)       # gray keeps the input dtype (operators.py:49-51)
        return y.astype(np.float64)

    def __repr__(self):
        return f"ObservationOperator({self.kind!r}, adjoint={self.adjoint}, r={self.r})"


def get_observation_operators(operator, path_kernel, r):
    h = None
    if operator == "blur" or (isinstance(path_kernel, np.ndarray)) or (path_kernel and os.path.exists(str(path_kernel))):
        try:
            h = load_blur_kernel(path_kernel)
        except Exception:
            if operator == "blur":
                raise
    return ObservationOperator(operator, False, h, r), ObservationOperator(operator, True, h, r)


# ---------------------------------------------------------------- prox library (numpy in / numpy out)
def _flat_engine(n: int):
    key = ("flat", n)
    e = _engines.get(key)
    if e is None:
        e = Engine(1, 1, 1, n, method="A", deg_op="Id", max_iter=1)
        _cache_put(key, e)
    return e


def proj_l2_ball(x, alpha_n, gaussian_nl, sp_nl, x_0, r=1):
    """operators.py:102-108"""
    x = np.asarray(x)
    e = _flat_engine(x.size)
    eps = l2_ball_radius(x.size, alpha_n, gaussian_nl, sp_nl, r)
    out = e.proj_l2_ball(e.to_device(x), e.to_device(np.broadcast_to(x_0, x.shape)), eps)
    return out.cpu().numpy().reshape(x.shape).astype(np.float64)


def proj_l1_ball(x, alpha_s, sp_nl, r=1):
    """operators.py:94-100"""
    x = np.asarray(x)
    e = _flat_engine(x.size)
    eta = l1_ball_radius(x.size, alpha_s, sp_nl, r)
    out = e.proj_l1_ball(e.to_device(x), eta)
    return out.cpu().numpy().reshape(x.shape).astype(np.float64)


def prox_GKL(x, gamma, alpha, x_0):
    """operators.py:114-115"""
    x = np.asarray(x)
    e = _flat_engine(x.size)
    out = e.prox_gkl(e.to_device(x), e.to_device(np.broadcast_to(x_0, x.shape)), gamma, alpha)
    return out.cpu().numpy().reshape(x.shape).astype(np.float64)


def grad_x_l2(x, s, phi, adj_phi, x_0):
    """operators.py:88-89"""
    return 2 * adj_phi(phi(x) + s - x_0)


def grad_s_l2(x, s, phi, x_0):
    """operators.py:91-92"""
    return phi(x) + s - x_0


def denoise(x, path_prox, ch):
    """operators.py:81-83 — the reference re-reads the checkpoint on every call; here the converted
    weights and the engine are cached per (path, shape)."""
    from .models.denoiser import Denoiser
    return Denoiser.cached(path_prox, ch).denoise(x)
