"""How much of pds_restore_host's transfer time is exposed: cfg4 shape (64 x 3 x 1024^2, ours-A, blur), 2 iterations per call so that
the 2.4 GB up / 0.8 GB down are a visible share; page-locked buffers (chunk-wise upload and download) against a pageable output
buffer (download in one piece at the end) and against the device-resident loop of the same length."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pnp_pds_b200.engine import Engine
from pnp_pds_b200.models.weights import load_weights
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
B, C, H, W, N = 64, 3, 1024, 1024, 2
w = load_weights(os.path.join(G, "weights", "DnCNN_nobn_nch_3_nlev_0.01.pdsw"))
hk = np.load(os.path.join(G, "assets.npz"))["blur_1"]
rng = np.random.default_rng(0)
pin = lambda: torch.empty((B, C, H, W), dtype=torch.float32).pin_memory()
x0, obs, xt, out = pin(), pin(), pin(), pin()
x0.copy_(torch.from_numpy(rng.random((B, C, H, W), dtype=np.float32))); obs.copy_(x0); xt.copy_(x0)
res = {}
with Engine(B, C, H, W, method="A", deg_op="blur", max_iter=N) as e:
    e.set_blur_kernel(hk); e.load_dncnn(w)
    e.set_params(dict(gamma1=0.99, gamma2=0.99, epsilon=0.95 * 0.01 * (C * H * W) ** 0.5, eta=0.0, lam=1.0, alpha=1.0))
    pageable = np.empty((B, C, H, W), dtype=np.float32)
    def timed(fn, reps=4):
        fn(); torch.cuda.synchronize(); best = 1e9
        for _ in range(reps):
            t = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
        return best
    res["restore_host_pinned_s"] = timed(lambda: e.restore_host(x0.numpy(), obs.numpy(), xt.numpy(), N, want_s=False, out=out.numpy()))
    res["restore_host_pageable_out_s"] = timed(lambda: e.restore_host(x0.numpy(), obs.numpy(), xt.numpy(), N, want_s=False, out=pageable))
    d0, dobs, dt_ = x0.cuda(), obs.cuda(), xt.cuda()
    def resident():
        e.set_problem(d0, dobs, dt_); e.run(N)
    res["resident_loop_s"] = timed(resident)
res["exposed_transfer_pinned_ms"] = 1e3 * (res["restore_host_pinned_s"] - res["resident_loop_s"])
res["exposed_transfer_pageable_out_ms"] = 1e3 * (res["restore_host_pageable_out_s"] - res["resident_loop_s"])
print(json.dumps(res))
