// lc_b200 — tiny-N path (sparse keypoints: N = 8 / 16, configs/gsplmo.yaml:7 `sparse_cnt: 16`, losses.py:329-334): ONE THREAD PER POSE.
//
// With a handful of correspondences a pose is almost pure O(1) work: the trust-region steps (6x6 Cholesky, sincos, bookkeeping)
// and the 6x6 forward / reverse algebra of the loss outweigh the point loops.  The CTA-per-pose kernels run those sections on one
// lane while 31 idle; here every lane of a warp owns a pose and runs the whole pipeline — LM solve (lm_advance, the same
// Ceres-faithful code), LC loss forward and backward — sequentially in fp64, so the serial sections of 32 poses execute in
// parallel.  Lanes whose solve needs more iterations simply stay active longer.
//
// Any element type (fp32 / fp64), any strides, every weight mode (incl. full 2x2), ragged n_points.  Per-thread state lives in
// registers and a few hundred bytes of local memory (L1): the packed sums H', G', b', C' = H'^-1 and the reverse coefficients;
// the 24 bbox-corner rows are recomputed where they are used (lc_six_fast's formulation: y = C' r, g = G' y, z = C' g).
// Math: SURVEY.md §8a / §8c; identical to lc_stream.cu (left-perturbation accumulation basis, no depth decoupling: fp64).
#include "lc_point.cuh"

namespace lc {

// resident 64-thread CTAs per SM the compiler must allow (register cap = 65536 / (64 * this)): 4 -> 255 registers, 8 -> 128
#ifndef LC_TINY_MIN_BLOCKS
#define LC_TINY_MIN_BLOCKS 4
#endif

struct TinyPose {          // what point_error / point_jac_left / lm_write_result need
    double K[9], pose[7], R[9], t[3];
    LmState lm;
};

// bbox-corner row (corner c3, axis ax) of jac_update2alter (cov_mixed.py:52-65) in the left basis: [Rb(-[c]x) R^-1 | e_ax]
__device__ __forceinline__ void tiny_row(const double* Rb, const double* Ri, const double* c3, int ax, double (&row)[6]) {
    const double nC[9] = {0, c3[2], -c3[1], -c3[2], 0, c3[0], c3[1], -c3[0], 0};
    double A3[3];
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) A3[cc] = Rb[ax * 3] * nC[cc] + Rb[ax * 3 + 1] * nC[3 + cc] + Rb[ax * 3 + 2] * nC[6 + cc];
#pragma unroll
    for (int m = 0; m < 3; ++m) row[m] = A3[0] * Ri[m] + A3[1] * Ri[3 + m] + A3[2] * Ri[6 + m];
    row[3] = ax == 0 ? 1.0 : 0.0; row[4] = ax == 1 ? 1.0 : 0.0; row[5] = ax == 2 ? 1.0 : 0.0;
}
__device__ __forceinline__ void tiny_symv(const double* Sp /* packed */, const double (&x)[6], double (&y)[6]) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) v = fma(Sp[i <= j ? sym_idx(i, j) : sym_idx(j, i)], x[j], v);
        y[i] = v;
    }
}

// LC loss forward + backward of pose b at s.pose, sequential (cov_mixed.py:100-150).  Separate (non-inlined) function: the
// solve above it keeps its own register allocation.
template <typename T>
__device__ __noinline__ void tiny_lc_phase(const lc_args& a, TinyPose& s, int b, int n) {
    double qn;
    quat_to_R_ref(s.pose, s.R, &qn);
    s.t[0] = s.pose[4]; s.t[1] = s.pose[5]; s.t[2] = s.pose[6];
    double Rb[9], Ri[9], bbox[24];
#pragma unroll
    for (int k = 0; k < 9; ++k) { const double id = (k % 4 == 0) ? 1.0 : 0.0; Rb[k] = qn * id + (s.R[k] - id); }
    {
        const double* R = s.R;
        const double c00 = R[4] * R[8] - R[5] * R[7], c01 = R[5] * R[6] - R[3] * R[8], c02 = R[3] * R[7] - R[4] * R[6];
        const double idet = 1.0 / (R[0] * c00 + R[1] * c01 + R[2] * c02);
        Ri[0] = c00 * idet; Ri[1] = (R[2] * R[7] - R[1] * R[8]) * idet; Ri[2] = (R[1] * R[5] - R[2] * R[4]) * idet;
        Ri[3] = c01 * idet; Ri[4] = (R[0] * R[8] - R[2] * R[6]) * idet; Ri[5] = (R[2] * R[3] - R[0] * R[5]) * idet;
        Ri[6] = c02 * idet; Ri[7] = (R[1] * R[6] - R[0] * R[7]) * idet; Ri[8] = (R[0] * R[4] - R[1] * R[3]) * idet;
    }
    for (int k = 0; k < 24; ++k) bbox[k] = ld<T>(a.bbox, b * a.bbox.stride[0] + (k / 3) * a.bbox.stride[1] + (k % 3) * a.bbox.stride[2]);
    int flag = 0;
    const double Lmax = a.max_err_len;

    // pass 1: sum_i valid_i |ec_ia|, sum_i valid_i      (cov_mixed.py:28-31)
    double a0 = 0.0, a1 = 0.0, cnt = 0.0;
    for (int i = 0; i < n; ++i) {
        PointIn p; PointErr e;
        load_point<T>(a, b, i, false, p);
        point_error(s, p, Lmax, e);
        a0 = fma(p.valid, fabs(e.ec[0]), a0); a1 = fma(p.valid, fabs(e.ec[1]), a1); cnt += p.valid;
    }
    const double vcnt = a.valid.ptr ? cnt : static_cast<double>(n);
    const double d0 = a.rel_thresh * (a0 / vcnt), d1 = a.rel_thresh * (a1 / vcnt);
    // pass 2: q_a = mean valid s^2 sigma    (cov_mixed.py:32-36)
    double q0 = 0.0, q1 = 0.0;
    for (int i = 0; i < n; ++i) {
        PointIn p; PointErr e;
        load_point<T>(a, b, i, true, p);
        point_error(s, p, Lmax, e);
        const double b0 = fabs(e.ec[0]), b1 = fabs(e.ec[1]);
        const double sg0 = b0 > d0 ? d0 * (2.0 * b0 - d0) : b0 * b0;
        const double sg1 = b1 > d1 ? d1 * (2.0 * b1 - d1) : b1 * b1;
        q0 = fma(p.valid * (p.s[0] * p.s[0]), sg0, q0);
        q1 = fma(p.valid * (p.s[1] * p.s[1]), sg1, q1);
    }
    // delta_k = sqrt(we q_a / (sigma_k + 1e-6)) = sq_a * rsqrt(sigma_k + 1e-6)
    const double sq0 = sqrt((q0 / vcnt) * a.w_e_thresh), sq1 = sqrt((q1 / vcnt) * a.w_e_thresh);
    // pass 3: H' = sum W J'J'^T, G' = sum W^2 sigma J'J'^T, b' = sum W ec J'   (left basis)
    double fin[48];
#pragma unroll
    for (int k = 0; k < 48; ++k) fin[k] = 0.0;
    for (int i = 0; i < n; ++i) {
        PointIn p; PointErr e;
        load_point<T>(a, b, i, true, p);
        point_error(s, p, Lmax, e);
        double J[2][6];
        point_jac_left(s, e.P, J);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const double dc = c ? d1 : d0, sq = c ? sq1 : sq0;
            const double av = fabs(e.ec[c]);
            const double sg = av > dc ? dc * (2.0 * av - dc) : av * av;
            const double del = sq * rsqrt(sg + 1e-6);
            const double sk = p.s[c];
            const double w = sk > del ? del * (2.0 * sk - del) : sk * sk;
            acc_outer<0>(fin, w, J[c]);
            acc_outer<21>(fin, w * w * sg, J[c]);
            const double wb = w * e.ec[c];
#pragma unroll
            for (int r = 0; r < 6; ++r) fin[42 + r] = fma(wb, J[c][r], fin[42 + r]);
        }
    }

    // ---- 6x6 forward (lc_six_fast's formulation, sequential) ----
    double Cp[kSym];   // C' = H'^-1 packed
    {
        double Hs[36], C[36];
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) Hs[r * 6 + c] = fin[r <= c ? sym_idx(r, c) : sym_idx(c, r)];
        if (chol6_inverse(Hs, C) != 0) {
            // safe_cholesky: non-SPD -> H_ref := I (pnp_utils.py:140-167), i.e. C' = T T^T with T = blockdiag(R, I)
            flag |= LC_ST_HESS_NOT_SPD;
            for (int k = 0; k < 36; ++k) C[k] = 0.0;
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) C[r * 6 + c] = s.R[r * 3] * s.R[c * 3] + s.R[r * 3 + 1] * s.R[c * 3 + 1] + s.R[r * 3 + 2] * s.R[c * 3 + 2];
            C[21] = 1.0; C[28] = 1.0; C[35] = 1.0;
        }
        for (int r = 0; r < 6; ++r)
            for (int c = r; c < 6; ++c) Cp[sym_idx(r, c)] = C[r * 6 + c];
    }
    double dth[6];
    {
        const double bv[6] = {fin[42], fin[43], fin[44], fin[45], fin[46], fin[47]};
        tiny_symv(Cp, bv, dth);
    }
    double sC[8], sM[8], sU[8];
    bool goodC = true, goodM = true;
    for (int j = 0; j < 8; ++j) {
        double c_ = 0.0, m_ = 0.0, u_ = 0.0;
        for (int ax = 0; ax < 3; ++ax) {
            double row[6], y[6], g[6];
            tiny_row(Rb, Ri, bbox + 3 * j, ax, row);
            tiny_symv(Cp, row, y);
            tiny_symv(fin + 21, y, g);
            double vc = 0.0, vm = 0.0, uu = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) { vc = fma(row[i], y[i], vc); vm = fma(y[i], g[i], vm); uu = fma(row[i], dth[i], uu); }
            goodC = goodC && vc > 0.0; goodM = goodM && vm > 0.0;
            c_ += vc; m_ += vm; u_ = fma(uu, uu, u_);
        }
        sC[j] = c_; sM[j] = m_; sU[j] = u_;
    }
    if (!goodC) flag |= LC_ST_PRIOR_NOT_GOOD;
    if (!goodM) flag |= LC_ST_COV_NOT_GOOD;
    double prior = 0.0, cov_err = 0.0, lin = 0.0;
    for (int j = 0; j < 8; ++j) { prior += goodC ? sqrt(sC[j]) : 1.0; cov_err += goodM ? sqrt(sM[j]) : 1.0; lin += sqrt(sU[j]); }
    prior *= 0.125; cov_err *= 0.125; lin *= 0.125;
    const double ip = 1.0 / prior;
    const double loss = log(prior) + 0.5 * (cov_err + lin) * ip;
    if (a.loss.ptr) st<T>(a.loss, b * a.loss.stride[0], loss);
    if (a.lc_flags) a.lc_flags[b] = flag;
    if (a.loss_sum) { atomicAdd(a.loss_sum, loss); atomicAdd(a.loss_sum + 1, 1.0); }
    if (a.cov.ptr || a.update_cov.ptr) {
        // reference-basis covariances on request: S_ref = Ti S' Ti^T with Ti = blockdiag(R^-1, I), M' = C' G' C'
        double C[36], M[36], T1[36];
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) C[r * 6 + c] = Cp[r <= c ? sym_idx(r, c) : sym_idx(c, r)];
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) {
                double v = 0.0;
                for (int k = 0; k < 6; ++k) v = fma(C[r * 6 + k], fin[21 + (k <= c ? sym_idx(k, c) : sym_idx(c, k))], v);
                T1[r * 6 + c] = v;
            }
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) {
                double v = 0.0;
                for (int k = 0; k < 6; ++k) v = fma(T1[r * 6 + k], C[k * 6 + c], v);
                M[r * 6 + c] = v;
            }
        auto ti = [&](int r, int c) { return (r < 3 && c < 3) ? Ri[r * 3 + c] : (r == c ? 1.0 : 0.0); };
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) {
                double vC = 0.0, vM = 0.0;
                for (int i = 0; i < 6; ++i) {
                    double wc = 0.0, wm = 0.0;
                    for (int k = 0; k < 6; ++k) { wc = fma(C[i * 6 + k], ti(c, k), wc); wm = fma(0.5 * (M[i * 6 + k] + M[k * 6 + i]), ti(c, k), wm); }
                    vC = fma(ti(r, i), wc, vC); vM = fma(ti(r, i), wm, vM);
                }
                if (a.cov.ptr) st<T>(a.cov, b * a.cov.stride[0] + r * a.cov.stride[1] + c * a.cov.stride[2], vC);
                if (a.update_cov.ptr) st<T>(a.update_cov, b * a.update_cov.stride[0] + r * a.update_cov.stride[1] + c * a.update_cov.stride[2], vM);
            }
    }
    if (!(a.g_pts3d.ptr || a.g_pts2d.ptr || a.g_weights.ptr)) return;

    // ---- 6x6 reverse: 24-term sums of outer products of y = C' r, z = C' G' y ----
    const double go = a.grad_scale * (a.grad_out.ptr ? ld<T>(a.grad_out, b * a.grad_out.stride[0]) : 1.0);
    const double g_p = go * (ip - 0.5 * (cov_err + lin) * ip * ip);
    const double g_c = go * 0.5 * ip;
    double cHL[kSym], cGL[kSym], bL[6];
#pragma unroll
    for (int k = 0; k < kSym; ++k) { cHL[k] = 0.0; cGL[k] = 0.0; }
#pragma unroll
    for (int k = 0; k < 6; ++k) bL[k] = 0.0;
    for (int j = 0; j < 8; ++j) {
        const double wC = goodC ? g_p * 0.0625 / sqrt(sC[j]) : 0.0;
        const double wM = goodM ? g_c * 0.0625 / sqrt(sM[j]) : 0.0;
        const double wU = sU[j] > 0.0 ? g_c * 0.125 / sqrt(sU[j]) : 0.0;
        for (int ax = 0; ax < 3; ++ax) {
            double row[6], y[6], g[6], z[6];
            tiny_row(Rb, Ri, bbox + 3 * j, ax, row);
            tiny_symv(Cp, row, y);
            tiny_symv(fin + 21, y, g);
            tiny_symv(Cp, g, z);
            double uu = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) uu = fma(row[i], dth[i], uu);
            int k = 0;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                bL[i] = fma(wU * uu, y[i], bL[i]);
#pragma unroll
                for (int jj = i; jj < 6; ++jj) {
                    const double f = i == jj ? 1.0 : 2.0;
                    const double yy = y[i] * y[jj];
                    cGL[k] = fma(f * wM, yy, cGL[k]);
                    cHL[k] = fma(f, fma(wC, yy, wM * fma(y[i], z[jj], z[i] * y[jj])), cHL[k]);
                    ++k;
                }
            }
        }
    }
    {
        const bool spd = !(flag & LC_ST_HESS_NOT_SPD);   // torch.where(cond, eye, H): no gradient into H
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int jj = i; jj < 6; ++jj) {
                const double bd = i == jj ? bL[i] * dth[i] : fma(bL[i], dth[jj], bL[jj] * dth[i]);
                cHL[k] = spd ? -(cHL[k] + bd) : 0.0;
                ++k;
            }
    }

    // pass 4: per-coordinate adjoints  
    for (int i = 0; i < a.N; ++i) {
        if (i >= n) {   // padding beyond n_points: defined, zero
            for (int c = 0; c < 2; ++c) {
                if (a.g_weights.ptr) st<T>(a.g_weights, b * a.g_weights.stride[0] + i * a.g_weights.stride[1] + c * a.g_weights.stride[2], 0.0);
                if (a.g_pts2d.ptr) st<T>(a.g_pts2d, b * a.g_pts2d.stride[0] + i * a.g_pts2d.stride[1] + c * a.g_pts2d.stride[2], 0.0);
            }
            if (a.g_pts3d.ptr)
                for (int c = 0; c < 3; ++c) st<T>(a.g_pts3d, b * a.g_pts3d.stride[0] + i * a.g_pts3d.stride[1] + c * a.g_pts3d.stride[2], 0.0);
            continue;
        }
        PointIn p; PointErr e;
        load_point<T>(a, b, i, true, p);
        point_error(s, p, Lmax, e);
        double J[2][6];
        point_jac_left(s, e.P, J);
        double ecb[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const double dc = c ? d1 : d0, sq = c ? sq1 : sq0;
            const double av = fabs(e.ec[c]);
            const bool big = av > dc;
            const double sg = big ? dc * (2.0 * av - dc) : av * av;
            const double del = sq * rsqrt(sg + 1e-6);
            const double sk = p.s[c];
            const bool wbig = sk > del;
            const double w = wbig ? del * (2.0 * sk - del) : sk * sk;
            double qh = 0.0, qg = 0.0, lb = 0.0;
            int k = 0;
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                lb = fma(J[c][r], bL[r], lb);
#pragma unroll
                for (int cc = r; cc < 6; ++cc) {
                    const double pp = J[c][r] * J[c][cc];
                    qh = fma(cHL[k], pp, qh);
                    qg = fma(cGL[k], pp, qg);
                    ++k;
                }
            }
            const double Wbar = qh + 2.0 * w * sg * qg + e.ec[c] * lb;
            const double sigbar = w * w * qg;
            if (a.g_weights.ptr)
                st<T>(a.g_weights, b * a.g_weights.stride[0] + i * a.g_weights.stride[1] + c * a.g_weights.stride[2], Wbar * (wbig ? 2.0 * del : 2.0 * sk));
            const double sgn = (e.ec[c] > 0.0) ? 1.0 : ((e.ec[c] < 0.0) ? -1.0 : 0.0);
            ecb[c] = sigbar * (big ? 2.0 * dc : 2.0 * av) * sgn;
            if (a.g_pts2d.ptr) st<T>(a.g_pts2d, b * a.g_pts2d.stride[0] + i * a.g_pts2d.stride[1] + c * a.g_pts2d.stride[2], ecb[c]);
        }
        if (a.g_pts3d.ptr) {
            // gX = -R^T (dproj/dP)^T ecbar,  dproj/dP = (K[:2,:] - proj (x) K[2,:] [z >= 0.1]) / max(z, 0.1)
            const double iz = 1.0 / e.zc;
            double gP[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double k2 = e.z_active ? s.K[6 + c] : 0.0;
                gP[c] = (fma(-e.proj[0], k2, s.K[c]) * ecb[0] + fma(-e.proj[1], k2, s.K[3 + c]) * ecb[1]) * iz;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
                st<T>(a.g_pts3d, b * a.g_pts3d.stride[0] + i * a.g_pts3d.stride[1] + c * a.g_pts3d.stride[2],
                      -(s.R[c] * gP[0] + s.R[3 + c] * gP[1] + s.R[6 + c] * gP[2]));
        }
    }
}

template <typename T, int MODE>
__global__ void __launch_bounds__(64, LC_TINY_MIN_BLOCKS) lc_tiny_kernel(const lc_args a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const int n = a.n_points ? min(max(a.n_points[b], 0), a.N) : a.N;
    const bool sanitize = (MODE & MODE_LM) && (a.flags & LC_FLAG_NAN_TO_NUM) != 0;
    TinyPose s;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const double v = ld<T>(a.K, b * a.K.stride[0] + (k / 3) * a.K.stride[1] + (k % 3) * a.K.stride[2]);
        s.K[k] = sanitize ? nan_to_num<T>(v) : v;
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const double v = ld<T>(a.pose, b * a.pose.stride[0] + k * a.pose.stride[1]);
        s.pose[k] = sanitize ? nan_to_num<T>(v) : v;
    }

    // =========================== LM solve ===========================
    if (MODE & MODE_LM) {
        LmState& L = s.lm;
        double* trace = a.trace ? a.trace + static_cast<int64_t>(b) * (a.max_iter + 2) * 4 : nullptr;
        bool solved = false;
        if (n >= 3) {
            quat_to_angle_axis(s.pose, L.x);
            L.x[3] = s.pose[4]; L.x[4] = s.pose[5]; L.x[5] = s.pose[6];
            lm_set_eval_point(L, L.x);
            L.ctl = CTL_EVAL_FULL;
            const double Kc[6] = {s.K[0], s.K[1], s.K[3], s.K[4], s.K[2], s.K[5]};
            bool first = true;
            for (;;) {
                const int kind = L.ctl;
                double acc[28];
#pragma unroll
                for (int k = 0; k < 28; ++k) acc[k] = 0.0;
                if (kind == CTL_EVAL_COST) for (int i = 0; i < n; ++i) lm_point_accum<T, false>(a, Kc, L.Rm, L.te, b, i, sanitize, acc);
                else for (int i = 0; i < n; ++i) lm_point_accum<T, true>(a, Kc, L.Rm, L.te, b, i, sanitize, acc);
                lm_advance(L, acc, kind, first, a.max_iter, a.function_tolerance, (a.flags & LC_FLAG_TOL_NEEDS_SUCCESS) != 0, trace);
                first = false;
                if (L.ctl == CTL_STOP) break;
            }
            solved = L.term == TERM_CONVERGENCE;
        }
        lm_write_result<T>(a, s, b, n, solved);
    }
    if (MODE & MODE_LC) tiny_lc_phase<T>(a, s, b, n);
}

template <typename T, int MODE>
static int launch_tiny_t(const lc_args& a, cudaStream_t st) {
    constexpr int NT = 64;
    lc_tiny_kernel<T, MODE><<<(a.B + NT - 1) / NT, NT, 0, st>>>(a);
    note_kernel("lc::lc_tiny_kernel<%s,%s>", sizeof(T) == 4 ? "float" : "double", MODE == MODE_LM ? "LM" : (MODE == MODE_LC ? "LC" : "LM|LC"));
    return static_cast<int>(cudaGetLastError());
}

int launch_tiny_pose(const lc_args& a, int mode, cudaStream_t st) {
    const bool f32 = a.dtype == LC_F32;
    switch (mode) {
        case MODE_LM: return f32 ? launch_tiny_t<float, MODE_LM>(a, st) : launch_tiny_t<double, MODE_LM>(a, st);
        case MODE_LC: return f32 ? launch_tiny_t<float, MODE_LC>(a, st) : launch_tiny_t<double, MODE_LC>(a, st);
        default: return f32 ? launch_tiny_t<float, MODE_LM | MODE_LC>(a, st) : launch_tiny_t<double, MODE_LM | MODE_LC>(a, st);
    }
}

}  // namespace lc
