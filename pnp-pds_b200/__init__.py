"""pnp_pds_b200 — B200-native PnP-PDS (primal-dual splitting with a DnCNN denoiser) hot path.

Drop-in for the reference's operator API on that path (yodai49/PnP-PDS):
  operators.get_observation_operators / proj_l2_ball / proj_l1_ball / prox_GKL / denoise
  iteration.test_iter
  main.test_all_images / eval_restoration
  models.denoiser.Denoiser, models.network_dncnn.DnCNN
All compute runs in hand-written sm_100a kernels behind the C ABI in include/pnp_pds.h
(libpnp_pds.so); there is no CPU fallback.
"""
__version__ = "0.1.0"
