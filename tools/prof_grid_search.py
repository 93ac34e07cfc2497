"""cProfile of one rank's share of the cfg5 grid search (32 images x 8 grid points of 256 x 256, 20 iterations) through main.grid_search:
where the host-side time of a short job goes."""
import cProfile, io, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import pds_oracle as O
from pnp_pds_b200 import main as pmain
from pnp_pds_b200.models.weights import load_weights
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
hk = np.load(os.path.join(G, "assets.npz"))["blur_1"]
w = load_weights(os.path.join(G, "weights", "DnCNN_nobn_nch_1_nlev_0.01.pdsw"))
images = [O.synthetic_image(i, 1, 256, 256) for i in range(32)]
grid = [dict(alpha_n=0.82 + 0.02 * k) for k in range(8)]
settings = dict(gaussian_nl=0.01, sp_nl=0.0, poisson_noise=False, poisson_alpha=300, deg_op="blur", r=1.0)
common = dict(method="ours-A", architecture="DnCNN_nobn_nch_1_nlev_0.01", max_iter=20, gamma1=0.99, gamma2=0.99, alpha_s=0.95, myLambda=1.0)
run = lambda: pmain.grid_search(images, grid, settings, common, 1, hk, w, batch_size=256)
run(); torch.cuda.synchronize()
t = time.perf_counter(); run(); torch.cuda.synchronize(); print("wall", time.perf_counter() - t)
pr = cProfile.Profile(); pr.enable(); run(); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[-3800:])
