"""lc_b200 — B200-native (sm_100a) implementation of the LC-loss hot path of fulliu/lc.

Public operators (same names and contracts as the reference modules they replace):

    lc_b200.cov_mixed.Loss_cov_mixed                 <- lib/cov_mixed.py
    lc_b200.nll.pnp_auto.weighted_pnp_jac_wrt_pts2d  <- lib/nll/pnp_auto.py
    lc_b200.nll.pnp_auto.diff_pnp_perturb
    lc_b200.pnp.cer_solver.solve                     <- lib/pnp/cer_solver.py
    lc_b200.fused.solve_and_loss                     (solve -> loss -> grads in one launch)
    lc_b200.dense.dense_pose_loss                    <- the glue of losses.py:355-386 fused with the loss (row f1)
    lc_b200.sharded.sharded_mean_loss                (batch-sharded mean with one scalar all-reduce)

All of them call hand-written CUDA kernels through the C ABI in include/lc_b200.h.
There is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
