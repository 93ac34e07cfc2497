"""CPU emulation of tensor-core operand-split schemes for the DnCNN body layers (design probe, not product).

Runs the ours-A blur loop of the oracle at 64x64 with the denoiser's 64->64 layers computed from rounded
operands (exact products, float64 accumulation) and reports the relative L2 distance of the final iterate to the
fp32 denoiser — the quantity the 1e-4 parity gate is written on.
"""
import sys, os
import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pds_oracle as po
import importlib
weights = importlib.import_module("pnp_pds_b200.models.weights")

def q16(t): return t.to(torch.float16).to(torch.float64)
def q8(t): return t.to(torch.float32).to(torch.float8_e4m3fn).to(torch.float64)
def qbf(t): return t.to(torch.bfloat16).to(torch.float64)

_E2M1 = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0], dtype=torch.float64)


def q4_block(t, dim):
    """e2m1 with one power-of-two scale per block of 32 elements along `dim` (mxfp4-style, ue8m0 scales)."""
    t = t.movedim(dim, -1)
    shp = t.shape
    b = t.reshape(*shp[:-1], shp[-1] // 32, 32)
    amax = b.abs().amax(-1, keepdim=True).clamp_min(1e-300)
    scale = torch.exp2(torch.ceil(torch.log2(amax / 6.0)))
    x = (b / scale).abs().clamp(max=6.0)
    idx = (x.unsqueeze(-1) - _E2M1).abs().argmin(-1)
    q = _E2M1[idx] * torch.sign(b) * scale
    return q.reshape(shp).movedim(-1, dim)


def make_denoise(layers, scheme, slope=0.01, sign=1.0, clamp=True):
    L = [(torch.from_numpy(w).double(), torch.from_numpy(b).double()) for w, b in layers]
    def conv(a, w): return F.conv2d(a, w, None, padding=1)
    def body(a, w):
        if scheme == "fp32":
            return conv(a, w)
        a_hi = q16(a); a_lo = a - a_hi
        w_hi = q16(w); w_lo = w - w_hi
        if scheme == "fp16x1":
            return conv(a_hi, w_hi)
        if scheme == "fp16x3":
            return conv(a_hi, w_hi) + conv(a_hi, q16(w_lo)) + conv(q16(a_lo), w_hi)
        if scheme == "fp16x2w":   # weights split only
            return conv(a_hi, w_hi) + conv(a_hi, q16(w_lo))
        if scheme == "fp8lo":      # the product's scheme (dncnn_tc.cu / pds_api.cu tc_split_scales)
            import math
            e = math.frexp(float(w.abs().max()))[1]
            S = 18 - e
            t = conv(q8(a), q8(w_lo * 2.0**S)) + conv(q8(a_lo * 2.0**10), q8(w_hi * 2.0**(S - 10)))
            return conv(a_hi, w_hi) + (t * 2.0**-S).float().double()
        if scheme == "fp8w":       # weights exact (w_hi + e4m3 w_lo), activations rounded to fp16 once: 1.5 MMA times, no a_lo plane
            import math
            e = math.frexp(float(w.abs().max()))[1]
            S = 18 - e
            t = conv(q8(a), q8(w_lo * 2.0**S))
            return conv(a_hi, w_hi) + (t * 2.0**-S).float().double()
        if scheme == "fp4lo":      # what a block-scaled fp4 correction operand would give (1.5 MMA times per layer)
            t = conv(q4_block(a, 1), q4_block(w_lo, 1)) + conv(q4_block(a_lo, 1), q4_block(w_hi, 1))
            return conv(a_hi, w_hi) + t.float().double()
        raise ValueError(scheme)
    def denoise(x):
        squeeze = x.ndim == 2
        a = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).double()
        a = a[None, None] if squeeze else a[None]
        if clamp:
            a = a.clamp(0, 1)
        x_in = a
        n = len(L)
        for i, (w, b) in enumerate(L):
            if 0 < i < n - 1 or scheme == "fp32":
                a = body(a, w) + b.view(1, -1, 1, 1)
            else:
                # first / last layer: 3-term fp16 split (as in the product)
                a_hi = q16(a); a_lo = q16(a - a_hi); w_hi = q16(w); w_lo = q16(w - w_hi)
                a = conv(a_hi, w_hi) + conv(a_hi, w_lo) + conv(a_lo, w_hi) + b.view(1, -1, 1, 1)
            a = a.float().double()          # fp32 accumulator read-out
            if i != n - 1:
                a = F.leaky_relu(a, slope).float().double()
        out = (x_in + a) if sign > 0 else (x_in - a)
        if clamp:
            out = out.clamp(0, 1)
        out = out.float()[0].numpy()
        return out[0] if squeeze else out
    return denoise

def forward_errors():
    """Single-forward max abs error of every checkpoint against the reference outputs in tests/golden/denoiser.npz."""
    g = np.load(os.path.join(ROOT, "tests/golden/denoiser.npz"))
    for arch in ["DnCNN_nobn_nch_1_nlev_0.01", "DnCNN_nobn_nch_3_nlev_0.01", "DnCNN_nobn_nch_1_nlev_0.009", "dncnn_15", "dncnn_color_blind", "dncnn3"]:
        w = weights.load_weights(os.path.join(ROOT, f"tests/golden/weights/{arch}.pdsw"))
        x, y = g[f"{arch}_x"], g[f"{arch}_y"]
        for scheme in ("fp16x1", "fp16x3", "fp8lo"):
            out = make_denoise(w.layers, scheme, w.slope, w.residual_sign, w.clamp)(x)
            print(f"{arch:30s} {scheme:7s} max abs err {np.max(np.abs(out.reshape(y.shape) - y)):.3e}  rel L2 {np.linalg.norm(out.reshape(y.shape) - y) / np.linalg.norm(y):.3e}", flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "forward":
        return forward_errors()
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    its = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    wts = weights.load_weights(os.path.join(ROOT, "tests/golden/weights/DnCNN_nobn_nch_1_nlev_0.01.pdsw"))
    layers = wts.layers if hasattr(wts, "layers") else wts
    g = np.load(os.path.join(ROOT, "tests/golden/assets.npz"))
    h = g["blur_1"] if "blur_1" in g else g[[k for k in g.files if "blur" in k][0]]
    img = po.synthetic_image(0, 1, N, N)
    phi, adj = po.make_operators("blur", h, 1.0)
    x0, obs = po.synthesize_observation(img, "blur", h, 1.0, 0.01, 0.0, False, 100)
    res = {}
    for scheme in sys.argv[3:] or ["fp32", "fp16x1", "fp16x3", "fp8lo", "fp16x2w", "fp4lo"]:
        den = make_denoise(layers, scheme)
        x, s, c, psnr, _ = po.pds_iterations(x0, obs, img, phi, adj, den, 0.99, 0.99, 0.9, 0.95, 1.0, 0.01, 0.0, 100, its, "A-Proposed")
        res[scheme] = (x, psnr[-1])
        if scheme != "fp32" and "fp32" in res:
            ref = res["fp32"][0]
            print(f"{scheme:8s} relL2 {np.linalg.norm(x - ref) / np.linalg.norm(ref):.3e}  dPSNR {psnr[-1] - res['fp32'][1]:+.5f}", flush=True)
        else:
            print(scheme, "psnr", psnr[-1], flush=True)

if __name__ == "__main__":
    main()
