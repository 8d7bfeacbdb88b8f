"""Component timings of the test-time chain at the zycbv test shape (128x128, dense_sample 1): selection, initialiser, LM on the
padded batch vs LM on a batch trimmed to max(n_points).  Run on the GPU box: python tools/bench_chain_components.py"""
import sys, torch, time
sys.path.insert(0, '/root/repo')
from lc_b200.synth import make_dense_outputs
from lc_b200.select import solve_pnp_dense, dense_point_select
from lc_b200.pnp import cer_solver, init_solver
for B, rows_out in ((32, 40), (256, 40), (32, 88), (256, 88)):      # object mask = 69 % / 31 % of the crop
    d = make_dense_outputs(8, 128, 128, 80)
    cu = {k: torch.cat([v] * (B // 8)).cuda() for k, v in d.items()}
    ml = torch.full((B, 1, 128, 128), 3.0, device="cuda"); ml[:, :, :rows_out] = -3
    xyz = cu["xyz_noc"].permute(0, 2, 3, 1)
    def t(fn, n=20):
        for _ in range(3): fn()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
    sel = dense_point_select(xyz, ml, xyz_weight_logits=cu["logits"], xyz_weights_scale=cu["scale"], noc_scale=cu["noc_scale"], sample=1)
    print(B, 'n_points', sel['n_points'][:4].tolist(), 'Nmax', sel['pts3d'].shape[1])
    inv, start, inl = init_solver.solve(cu["K"], sel["pts3d"], sel["pts2d"], weights=sel["inv_cov"], n_points=sel["n_points"])
    print(' select us', t(lambda: dense_point_select(xyz, ml, xyz_weight_logits=cu["logits"], xyz_weights_scale=cu["scale"], noc_scale=cu["noc_scale"], sample=1)))
    print(' init us', t(lambda: init_solver.solve(cu["K"], sel["pts3d"], sel["pts2d"], weights=sel["inv_cov"], n_points=sel["n_points"])))
    print(' lm (padded N=16384) us', t(lambda: cer_solver.solve(cu["K"], sel["pts3d"], sel["pts2d"], sel["inv_cov"], start, sel["n_points"], filter_input_nan=True)))
    nmax = int(sel["n_points"].max())
    p3, p2, ic = sel["pts3d"][:, :nmax].contiguous(), sel["pts2d"][:, :nmax].contiguous(), sel["inv_cov"][:, :nmax].contiguous()
    print(' lm (trimmed N=%d) us' % nmax, t(lambda: cer_solver.solve(cu["K"], p3, p2, ic, start, sel["n_points"], filter_input_nan=True)))
