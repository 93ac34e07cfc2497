"""The PnP-PDS iteration with the reference's entry point (iteration.py:10-196), resident on one B200.

test_iter(...) keeps the reference's 21-argument signature and 6-tuple result; run_batch(...) is the
batched form (B independent restorations, per-item hyper-parameters) that the drivers and the
multi-GPU sharding use.  The loop itself is pds_run / pds_restore_host in libpnp_pds.so.
"""
from __future__ import annotations

import time
from typing import Sequence

import numpy as np

from .engine import (DENOISER_FREE, RESIDENT_METHODS, Engine, canonical_method, l1_ball_radius, l2_ball_radius, metrics_from_traces,
                     ssim_from_traces)
from .models.weights import DnCNNWeights, load_weights
from .operators import ObservationOperator, sampling_mask
from .utils.utils_eval import eval_ssim

_OUT_OF_SCOPE = {
    "A-PnPPDS-BM3D", "A-PnPFBS-BM3D", "comparisonB-1", "C-PnPPDS-BM3D",      # need the bm3d wheel (CPU-only algorithm)
}
_NO_REFERENCE_BEHAVIOUR = {"comparisonB-4", "comparisonB-5"}   # the reference dies with UnboundLocalError (iteration.py:40,143,148)


def _check_ops(phi, adj_phi):
    if not isinstance(phi, ObservationOperator) or not isinstance(adj_phi, ObservationOperator):
        raise TypeError("phi/adj_phi must come from pnp_pds_b200.operators.get_observation_operators "
                        "(arbitrary Python callables would need a CPU path, which does not exist)")
    if phi.kind != adj_phi.kind or phi.adjoint or not adj_phi.adjoint:
        raise ValueError("phi/adj_phi are not a matching (phi, adj_phi) pair")
    if phi.kind not in ("blur", "random_sampling", "Id"):
        raise ValueError(f"unknown deg_op {phi.kind!r}")
    return phi.kind


def item_params(method_id: str, n: int, gamma1, gamma2, alpha_s, alpha_n, myLambda, gaussian_nl, sp_nl, poisson_alpha, r):
    """Hyper-parameters as the kernels consume them.  Quirk Q1: A-Proposed never passes r
    to proj_l2_ball (iteration.py:52) so epsilon uses r = 1; B-Proposed passes r to both projections
    (iteration.py:56,58)."""
    r_l2 = r if method_id == "B" else 1.0
    # comparisonB-2 and comparisonB-3 call both projections without r (iteration.py:131,135,140, admm.py:43)
    r_l1 = 1.0 if method_id in ("ADMM_B2", "TV_B3") else r
    return dict(gamma1=gamma1, gamma2=gamma2, epsilon=l2_ball_radius(n, alpha_n, gaussian_nl, sp_nl, r_l2),
                eta=l1_ball_radius(n, alpha_s, sp_nl, r_l1), lam=myLambda, alpha=poisson_alpha)


# ---------------------------------------------------------------------------------------------------------------------
# Engine cache.  The reference rebuilds its denoiser (and re-reads the .pth) on every test_iter call (iteration.py:34-41,
# operators.py:81-83); creating an Engine costs a cudaMalloc of the whole workspace plus the weight swizzle and upload, so
# consecutive run_batch / test_iter / grid_search calls of one shape, method, operator and checkpoint share one handle.
# pds_restore_host resets the whole loop state, so a reused handle starts from a clean slate.
# ---------------------------------------------------------------------------------------------------------------------
_ENGINE_CACHE: "dict[tuple, Engine]" = {}
_ENGINE_CACHE_MAX = 4
_ENGINE_CACHE_BYTES = 48 << 30          # keep at most this much device workspace alive in cached engines


def clear_engine_cache():
    while _ENGINE_CACHE:
        _ENGINE_CACHE.popitem()[1].close()


def _weights_key(weights):
    if weights is None:
        return None
    k = getattr(weights, "_cache_key", None)
    if k is None:
        import hashlib
        k = hashlib.blake2b(weights.to_blob(), digest_size=16).hexdigest()
        try:
            weights._cache_key = k
        except Exception:
            pass
    return k


def _operator_key(kind, phi):
    if kind == "blur":
        import hashlib
        return hashlib.blake2b(np.ascontiguousarray(phi.h, dtype=np.float64).tobytes(), digest_size=16).hexdigest()
    if kind == "random_sampling":
        return float(phi.r)
    return None


def _acquire_engine(B, C, H, W, mid, kind, phi, max_iter, conv_engine, device, denoiser_chunk, weights):
    import torch
    dev = torch.cuda.current_device() if (device is None and torch.cuda.is_available()) else device
    key = (B, C, H, W, mid, kind, _operator_key(kind, phi), conv_engine, dev, int(denoiser_chunk), _weights_key(weights))
    eng = _ENGINE_CACHE.pop(key, None)
    if eng is not None and eng.max_iter < max_iter:        # trace capacity too small: rebuild
        eng.close()
        eng = None
    if eng is None:
        # trace capacity in powers of two from 32: a short warm-up call followed by the real one (or a sweep over iteration counts)
        # reuses the handle instead of rebuilding it (workspace, weights, page-locked staging) inside the second call
        cap = 32
        while cap < max_iter:
            cap *= 2
        eng = Engine(B, C, H, W, method=mid, deg_op=kind, max_iter=cap, conv_engine=conv_engine, device=device,
                     denoiser_chunk=denoiser_chunk)
        try:
            if kind == "blur":
                eng.set_blur_kernel(phi.h)
            elif kind == "random_sampling":
                eng.set_mask(sampling_mask(H, W, phi.r))
            if weights is not None:
                eng.load_dncnn(weights)
        except Exception:
            eng.close()
            raise
    return key, eng


def _release_engine(key, eng):
    _ENGINE_CACHE[key] = eng                               # most recently used last
    while len(_ENGINE_CACHE) > _ENGINE_CACHE_MAX or (len(_ENGINE_CACHE) > 1 and
                                                     sum(e.workspace_bytes for e in _ENGINE_CACHE.values()) > _ENGINE_CACHE_BYTES):
        _ENGINE_CACHE.pop(next(iter(_ENGINE_CACHE))).close()


def run_batch(x_0, x_obsrv, x_true, phi, adj_phi, params: Sequence[dict] | dict, path_prox, max_iter: int,
              method: str = "A-Proposed", ch: int = 3, conv_engine: str = "tcgen05", device=None, ssim: str = "final",
              denoiser_chunk: int = 0, m1: int = 15, m2: int = 15, gammaInADMMStep1: float = 0.1):
    """B restorations at once.  x_0, x_obsrv, x_true: (B,H,W) for ch=1 or (B,3,H,W).
    params: dict(s) with gamma1, gamma2, alpha_s, alpha_n, myLambda, gaussian_nl, sp_nl, poisson_alpha, r.
    Returns dict(x, s, c, psnr, ssim, time_per_iter, traces, launches)."""
    name = canonical_method(method)
    if name in _OUT_OF_SCOPE:
        raise ValueError(f"method {method!r} is outside the B200 hot path (BM3D baselines, see DESIGN.md)")
    if name in _NO_REFERENCE_BEHAVIOUR:
        raise NotImplementedError(f"method {method!r} cannot run in the reference either (denoiser_J is never constructed for it)")
    if name not in RESIDENT_METHODS:
        raise ValueError(f"Unknown method: {method}")
    mid = RESIDENT_METHODS[name]
    kind = _check_ops(phi, adj_phi)
    # x_0 / x_obsrv / x_true: one array (B, ...) or a sequence of B per-item arrays (converted straight into the engine's page-locked
    # staging buffers, without stacking them first)
    as_items = isinstance(x_0, (list, tuple))
    if not as_items:
        x_0 = np.asarray(x_0)
    B = len(x_0)
    item_shape = tuple(np.shape(x_0[0]))
    H, W = item_shape[-2:]
    C = 1 if len(item_shape) == 2 else item_shape[0]
    batch_shape = (B,) + item_shape
    if C != ch:
        raise ValueError(f"ch={ch} but the images have {C} channels")
    n = C * H * W
    if mid in DENOISER_FREE:
        if C != 3:
            raise ValueError(f"'{name}' is defined for colour images only (operators.py:122-123 hard-code three channels)")
        weights = None                                               # path_prox is ignored, as in the reference (iteration.py:34-41)
    else:
        weights = path_prox if isinstance(path_prox, DnCNNWeights) else load_weights(str(path_prox))
    if "unstable" in name:
        nb = 20 if ch == 3 else 17                                   # iteration.py:34-39
        if weights.depth != nb or weights.residual_sign > 0:
            raise RuntimeError(f"'{name}' expects a KAIR DnCNN checkpoint with nb={nb}")
    plist = [params] * B if isinstance(params, dict) else list(params)
    if len(plist) != B:
        raise ValueError("need one parameter dict per item")
    items = [item_params(mid, n, p["gamma1"], p["gamma2"], p.get("alpha_s", 1), p.get("alpha_n", 1), p.get("myLambda", 1),
                         p.get("gaussian_nl", 0), p.get("sp_nl", 0), p.get("poisson_alpha", 300), p.get("r", phi.r if kind == "random_sampling" else 1))
             for p in plist]
    key, eng = _acquire_engine(B, C, H, W, mid, kind, phi, max(1, int(max_iter)), conv_engine, device, denoiser_chunk, weights)
    try:
        eng.set_params(items)
        if mid in ("ADMM_B2", "ADMM_C", "RED_C"):
            eng.set_admm(m1, m2, gammaInADMMStep1)
        eng.set_ssim(ssim if x_true is not None else "none")
        import torch
        launches0 = eng.kernel_launches
        t0 = time.perf_counter()
        sparse = mid in ("B", "ADMM_B2", "TV_B3")                     # the methods that carry s; for the others s == 0 (iteration.py:24)
        if eng.staging() is not None:
            hx0, hobs = eng.stage("x0", x_0), eng.stage("obs", x_obsrv)
            htrue = None if x_true is None else eng.stage("true", x_true)
            st = eng.staging()
            x, s, tr = eng.restore_host(hx0, hobs, htrue, int(max_iter), want_s=sparse, out=st["x"].numpy(), s_out=st["s"].numpy())
            x = x.copy()                                               # the staging buffers belong to the (cached) engine
            s = s.copy() if sparse else None
        else:
            stack = lambda a: None if a is None else (np.stack([np.asarray(v) for v in a]) if isinstance(a, (list, tuple)) else a)
            x, s, tr = eng.restore_host(stack(x_0), stack(x_obsrv), stack(x_true), int(max_iter), want_s=sparse)
        if s is None:
            s = np.zeros(batch_shape, dtype=np.float32)
        torch.cuda.synchronize(eng.device)
        wall = time.perf_counter() - t0
        launches = eng.kernel_launches - launches0
    except Exception:
        eng.close()
        raise
    _release_engine(key, eng)
    c, psnr = metrics_from_traces(tr, n)                              # [it, B]
    x = x.reshape(batch_shape)
    s = s.reshape(batch_shape)
    ssim_data = ssim_from_traces(tr, C, H, W)                          # [it, B]; evaluated on the device
    return dict(x=x, s=s, c=c, psnr=psnr, ssim=ssim_data, time_per_iter=wall / max(1, int(max_iter)), traces=tr,
                launches=launches)


def test_iter(x_0, x_obsrv, x_true, phi, adj_phi, gamma1, gamma2, alpha_s, alpha_n, myLambda, m1, m2, gammaInADMMStep1,
              gaussian_nl, sp_nl, poisson_alpha, path_prox, max_iter, method="A-Proposed", ch=3, r=1):
    """Same arguments (and argument order: ..., gamma2, alpha_s, alpha_n, ...) and result tuple as the
    reference (iteration.py:10,196): (x_n, s_n + 0.5, c, psnr_data, ssim_data, average_time).

    Differences, by design: x_n is float32 (as the reference's denoiser output is); ssim_data is evaluated on the
    device every iteration like the reference (iteration.py:189; run_batch(ssim="final") keeps only the last
    entry and leaves NaN before it); average_time is wall seconds
    per iteration including the H2D/D2H copies, not process CPU seconds; an unknown method raises
    ValueError instead of printing and crashing on an unbound variable (iteration.py:183-185)."""
    p = dict(gamma1=gamma1, gamma2=gamma2, alpha_s=alpha_s, alpha_n=alpha_n, myLambda=myLambda, gaussian_nl=gaussian_nl,
             sp_nl=sp_nl, poisson_alpha=poisson_alpha, r=r)
    x_0 = np.asarray(x_0)
    res = run_batch(x_0[None], np.asarray(x_obsrv)[None], None if x_true is None else np.asarray(x_true)[None], phi, adj_phi,
                    p, path_prox, max_iter, method, ch, m1=m1, m2=m2, gammaInADMMStep1=gammaInADMMStep1, ssim="all")
    return (res["x"][0], res["s"][0].astype(np.float64) + 0.5, res["c"][:, 0], res["psnr"][:, 0], res["ssim"][:, 0],
            res["time_per_iter"])


test_iter.__test__ = False  # not a pytest test
