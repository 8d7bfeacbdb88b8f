// lc_b200 — sm_100a kernels for the LC hot path and their C ABI (see include/lc_b200.h).
//
// One CTA per pose.  A pose's N correspondences are streamed (coalesced for the planar layout
// the dense call site produces) by the CTA's threads; every pass ends in one multi-value CTA
// reduction of at most 48 doubles, followed by a short 6x6 section.  Nothing per-point is ever
// written to HBM except the requested gradients, so DRAM traffic is the algorithmic
// 48*N + O(1) bytes per pose (re-reads between passes hit L1/L2: one pose is <= 112 KB).
//
//   lc_pose_kernel<T, NT, MODE>
//     MODE & 1  LM   : Ceres-faithful Levenberg-Marquardt (replaces ceres.cpp:72-145), fp64
//     MODE & 2  LC   : Loss_cov_mixed forward + backward (replaces cov_mixed.py:100-150), fp64
//   lc_jac_kernel / lc_jac_bwd_kernel : weighted_pnp_jac_wrt_pts2d forward / double-backward
//
// Math reference: SURVEY.md §8a ("LC math") and §8c (solver spec); oracle/*.c are the CPU checkers.
#include <cfloat>
#include <cstdio>
#include <cstring>

#include "lc_device.cuh"

namespace lc {

enum { MODE_LM = 1, MODE_LC = 2 };
enum { TERM_CONVERGENCE = 0, TERM_NO_CONVERGENCE = 1, TERM_FAILURE = 2 };
enum { CTL_CONTINUE = 0, CTL_STOP = 1 };

struct LmState {
    double x[6], xc[6];        // accepted point / candidate, [angle-axis, t]
    double A[kSym], gs[6];     // scaled J^T J (packed) and scaled gradient at x
    double scale[6], diag[6];
    double cost, radius, dec, xnorm, gmax, model_change, reported_radius;
    double Rm[9], Jl[9], te[3];  // rotation, left Jacobian and translation of the evaluation point
    int reuse_diag, n_invalid, it, step_ok, any_success, ctl, term;
};

struct PoseShared {
    double K[9], pose[7], R[9], Rb[9], t[3], bbox[24];
    double red[kMaxWarps * 48];
    double fin[48];
    double H[36], G[36], C[36], M[36], T1[36], T2[36], Cbar[36], Mbar[36], Gbar[36], Hbar[36];
    double bv[6], dth[6], dthbar[6], bbar[6];
    double rows[24 * 6], vC[24], vM[24], u[24];
    double wC[8], wM[8], wU[8];
    double cHL[kSym], cGL[kSym], bL[6];  // backward coefficients, left basis, packed (off-diagonals doubled)
    double stat[8];                      // d0, d1, q0*we, q1*we, go
    int flag;
    LmState lm;
};

// ---------------------------------------------------------------------------------------------
// per-point geometry
// ---------------------------------------------------------------------------------------------
struct PointIn {
    double X[3], x[2], s[2], valid;
};

template <typename T>
__device__ __forceinline__ void load_point(const lc_args& a, int b, int i, bool want_s, PointIn& p) {
    const int64_t o3 = b * a.pts3d.stride[0] + i * a.pts3d.stride[1];
    p.X[0] = ld<T>(a.pts3d, o3);
    p.X[1] = ld<T>(a.pts3d, o3 + a.pts3d.stride[2]);
    p.X[2] = ld<T>(a.pts3d, o3 + 2 * a.pts3d.stride[2]);
    const int64_t o2 = b * a.pts2d.stride[0] + i * a.pts2d.stride[1];
    p.x[0] = ld<T>(a.pts2d, o2);
    p.x[1] = ld<T>(a.pts2d, o2 + a.pts2d.stride[2]);
    if (want_s) {
        const int64_t os = b * a.weights.stride[0] + i * a.weights.stride[1];
        p.s[0] = ld<T>(a.weights, os);
        p.s[1] = ld<T>(a.weights, os + a.weights.stride[2]);
    }
    p.valid = a.valid.ptr ? ld<T>(a.valid, b * a.valid.stride[0] + i * a.valid.stride[1]) : 1.0;
}

// project_apply + clamp_error (transforms.py:47-63, cov_mixed.py:16-24).  P = R X + t is returned for the Jacobian.
struct PointErr {
    double P[3], proj[2], ec[2], zc;
    bool z_active;
};

__device__ __forceinline__ void point_error(const PoseShared& s, const PointIn& p, double Lmax, PointErr& e) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
        e.P[r] = fma(s.R[r * 3], p.X[0], fma(s.R[r * 3 + 1], p.X[1], fma(s.R[r * 3 + 2], p.X[2], s.t[r])));
    double KP[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) KP[r] = fma(s.K[r * 3], e.P[0], fma(s.K[r * 3 + 1], e.P[1], s.K[r * 3 + 2] * e.P[2]));
    e.z_active = KP[2] >= 0.1;
    e.zc = KP[2] > 0.1 ? KP[2] : 0.1;
    const double iz = 1.0 / e.zc;
    e.proj[0] = KP[0] * iz;
    e.proj[1] = KP[1] * iz;
    double e0 = p.x[0] - e.proj[0], e1 = p.x[1] - e.proj[1];
    const double len = sqrt(fma(e0, e0, e1 * e1)) + 1e-6;
    if (len > Lmax) {
        const double f = (len - Lmax) / len;
        e0 = fma(-f, e0, e0);
        e1 = fma(-f, e1, e1);
    }
    e.ec[0] = e0;
    e.ec[1] = e1;
}

// Rows of the 2x6 Jacobian of residual_with_jac6d (pnp_auto.py:33-54) in the LEFT basis:
//   J'_a = [ q x D_a | D_a ],  q = R X,  D_a = (1/z) [K_a0, K_a1, -(K_a0 u0 + K_a1 v0)]
// The reference's right-perturbation Jacobian is J_a = J'_a . blockdiag(R, I); the 6x6 sums are
// transformed once per pose instead of once per point.
__device__ __forceinline__ void point_jac_left(const PoseShared& s, const double* P, double (&J)[2][6]) {
    const double iz = 1.0 / P[2];
    const double u0 = P[0] * iz, v0 = P[1] * iz;
    const double q0 = P[0] - s.t[0], q1 = P[1] - s.t[1], q2 = P[2] - s.t[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const double k0 = s.K[a * 3], k1 = s.K[a * 3 + 1];
        const double d0 = k0 * iz, d1 = k1 * iz, d2 = -fma(k0, u0, k1 * v0) * iz;
        J[a][0] = fma(q1, d2, -q2 * d1);
        J[a][1] = fma(q2, d0, -q0 * d2);
        J[a][2] = fma(q0, d1, -q1 * d0);
        J[a][3] = d0;
        J[a][4] = d1;
        J[a][5] = d2;
    }
}

// acc[OFF .. OFF+21) += w * J J^T (packed upper)
template <int OFF, int V>
__device__ __forceinline__ void acc_outer(double (&acc)[V], double w, const double (&J)[6]) {
    double wJ[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) wJ[i] = w * J[i];
    int k = OFF;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = i; j < 6; ++j) {
            acc[k] = fma(wJ[i], J[j], acc[k]);
            ++k;
        }
}

// ---------------------------------------------------------------------------------------------
// basis change helpers (threads 0..35 / single thread), T = blockdiag(Rm, I)
// ---------------------------------------------------------------------------------------------
// out (6x6 full) = T^T S T with S packed-symmetric in the left basis
__device__ inline double tts_entry(const double* Sp, const double* Rm, int r, int c) {
    // (T^T S T)_rc = sum_ij T_ir S_ij T_jc ; T = [[Rm,0],[0,I]]
    double s = 0.0;
    for (int i = 0; i < 6; ++i) {
        const double tir = (i < 3) ? (r < 3 ? Rm[i * 3 + r] : 0.0) : (i == r ? 1.0 : 0.0);
        if (tir == 0.0) continue;
        for (int j = 0; j < 6; ++j) {
            const double tjc = (j < 3) ? (c < 3 ? Rm[j * 3 + c] : 0.0) : (j == c ? 1.0 : 0.0);
            if (tjc == 0.0) continue;
            const double sij = Sp[i <= j ? sym_idx(i, j) : sym_idx(j, i)];
            s = fma(tir * sij, tjc, s);
        }
    }
    return s;
}
// (T M T^T)_rc for a full 6x6 M (right basis) -> left basis
__device__ inline double tmt_entry(const double* M, const double* Rm, int r, int c) {
    double s = 0.0;
    for (int i = 0; i < 6; ++i) {
        const double tri = (r < 3) ? (i < 3 ? Rm[r * 3 + i] : 0.0) : (i == r ? 1.0 : 0.0);
        if (tri == 0.0) continue;
        for (int j = 0; j < 6; ++j) {
            const double tcj = (c < 3) ? (j < 3 ? Rm[c * 3 + j] : 0.0) : (j == c ? 1.0 : 0.0);
            if (tcj == 0.0) continue;
            s = fma(tri * M[i * 6 + j], tcj, s);
        }
    }
    return s;
}

// ---------------------------------------------------------------------------------------------
// LM: Ceres 2.1.0 TrustRegionMinimizer + LevenbergMarquardtStrategy (see oracle/lm_oracle.c)
// ---------------------------------------------------------------------------------------------
__device__ inline void lm_set_eval_point(LmState& L, const double* x) {
    // ceres/rotation.h AngleAxisRotatePoint as a matrix, and the left Jacobian of SO(3):
    //   d(R(w) X)/dw = -[R X]x Jl(w)
    const double w0 = x[0], w1 = x[1], w2 = x[2];
    const double th2 = w0 * w0 + w1 * w1 + w2 * w2;
    double a, bq, cq;  // R = I + a [w]x + bq [w]x^2 ; Jl = I + bq [w]x + cq [w]x^2
    if (th2 > DBL_EPSILON) {
        const double th = sqrt(th2);
        double sn, cs;
        sincos(th, &sn, &cs);
        a = sn / th;
        if (th2 > 1e-6) {
            bq = (1.0 - cs) / th2;
            cq = (th - sn) / (th2 * th);
        } else {
            bq = 0.5 - th2 / 24.0;
            cq = 1.0 / 6.0 - th2 / 120.0;
        }
    } else {
        a = 1.0; bq = 0.0; cq = 0.0;  // R = I + [w]x (first-order branch of AngleAxisRotatePoint)
    }
    const double W[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
    double W2[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) W2[r * 3 + c] = W[r * 3] * W[c] + W[r * 3 + 1] * W[3 + c] + W[r * 3 + 2] * W[6 + c];
    const double bR = (th2 > DBL_EPSILON) ? bq : 0.0;
    for (int k = 0; k < 9; ++k) {
        const double id = (k % 4 == 0) ? 1.0 : 0.0;
        L.Rm[k] = id + a * W[k] + bR * W2[k];
        L.Jl[k] = id + bq * W[k] + cq * W2[k];
    }
    L.te[0] = x[3]; L.te[1] = x[4]; L.te[2] = x[5];
}

// fin[0..21) = S (left basis, packed), fin[21..27) = J'^T r, fin[27] = cost at the evaluation point.
// Builds the scaled normal matrix / gradient at that point into L.A / L.gs, returns false if not finite.
__device__ inline bool lm_take_normal_eq(LmState& L, const double* fin, bool first) {
    double JtJ[kSym], g[6];
    // T = blockdiag(Jl, I): JtJ = T^T S T, g = T^T gv
    for (int r = 0; r < 6; ++r)
        for (int c = r; c < 6; ++c) JtJ[sym_idx(r, c)] = tts_entry(fin, L.Jl, r, c);
    for (int r = 0; r < 3; ++r) g[r] = L.Jl[r] * fin[21] + L.Jl[3 + r] * fin[22] + L.Jl[6 + r] * fin[23];
    for (int r = 3; r < 6; ++r) g[r] = fin[21 + r];
    bool finite = isfinite(fin[27]);
    for (int k = 0; k < kSym; ++k) finite = finite && isfinite(JtJ[k]);
    for (int k = 0; k < 6; ++k) finite = finite && isfinite(g[k]);
    if (!finite) return false;
    if (first)
        for (int k = 0; k < 6; ++k) L.scale[k] = 1.0 / (1.0 + sqrt(JtJ[sym_idx(k, k)]));
    for (int r = 0; r < 6; ++r)
        for (int c = r; c < 6; ++c) L.A[sym_idx(r, c)] = JtJ[sym_idx(r, c)] * L.scale[r] * L.scale[c];
    double gmax = 0.0;
    for (int k = 0; k < 6; ++k) {
        L.gs[k] = g[k] * L.scale[k];
        // Ceres: |x - Plus(x, -g)|_inf, evaluated in floating point
        const double xs = __dadd_rn(L.x[k], -g[k]);
        gmax = fmax(gmax, fabs(__dsub_rn(L.x[k], xs)));
    }
    L.gmax = gmax;
    L.cost = fin[27];
    double xn = 0.0;
    for (int k = 0; k < 6; ++k) xn = fma(L.x[k], L.x[k], xn);
    L.xnorm = sqrt(xn);
    return true;
}

// Thread 0: consume the evaluation in fin[], advance the trust-region loop until the next
// evaluation point is known (L.ctl = CTL_CONTINUE) or the solve has terminated (CTL_STOP).
__device__ inline void lm_advance(LmState& L, const double* fin, bool first, int max_iter, double ftol, bool tol_guard,
                                  double* trace) {
    const double gtol = 1e-10, ptol = 1e-8, min_rel_dec = 1e-3;
    const double min_radius = 1e-32, max_radius = 1e16, min_diag = 1e-6, max_diag = 1e32;
    if (first) {
        L.radius = 1e4; L.dec = 2.0; L.reuse_diag = 0; L.n_invalid = 0; L.it = 0; L.any_success = 0;
        L.reported_radius = L.radius; L.term = TERM_FAILURE; L.model_change = 0.0;
        if (!lm_take_normal_eq(L, fin, true)) { L.ctl = CTL_STOP; return; }
        L.step_ok = 1;
    } else {
        const double cost_c = isfinite(fin[27]) ? fin[27] : DBL_MAX;
        const bool armed = !tol_guard || L.any_success;
        double sn = 0.0;
        for (int k = 0; k < 6; ++k) { const double d = L.x[k] - L.xc[k]; sn = fma(d, d, sn); }
        sn = sqrt(sn);
        if (armed && sn <= ptol * (L.xnorm + ptol)) { L.term = TERM_CONVERGENCE; L.ctl = CTL_STOP; return; }
        if (armed && fabs(L.cost - cost_c) <= ftol * L.cost) { L.term = TERM_CONVERGENCE; L.ctl = CTL_STOP; return; }
        const double rho = cost_c >= DBL_MAX ? -DBL_MAX : (L.cost - cost_c) / L.model_change;
        if (rho > min_rel_dec) {
            for (int k = 0; k < 6; ++k) L.x[k] = L.xc[k];
            if (!lm_take_normal_eq(L, fin, false)) { L.term = TERM_FAILURE; L.ctl = CTL_STOP; return; }
            const double t = 2.0 * rho - 1.0;
            L.radius = fmin(max_radius, L.radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
            L.dec = 2.0; L.reuse_diag = 0; L.step_ok = 1; L.any_success = 1;
        } else {
            L.radius = L.radius / L.dec; L.dec *= 2.0; L.reuse_diag = 1; L.step_ok = 0;
        }
    }
    for (;;) {
        // FinalizeIterationAndCheckIfMinimizerCanContinue
        L.reported_radius = L.radius;
        if (trace) { double* tr = trace + 4 * L.it; tr[0] = L.cost; tr[1] = L.radius; tr[2] = L.step_ok; tr[3] = L.gmax; }
        if (L.it >= max_iter) { L.term = TERM_NO_CONVERGENCE; L.ctl = CTL_STOP; return; }
        if (L.step_ok && L.gmax <= gtol) { L.term = TERM_CONVERGENCE; L.ctl = CTL_STOP; return; }
        if (L.radius <= min_radius) { L.term = TERM_CONVERGENCE; L.ctl = CTL_STOP; return; }
        ++L.it;
        // LevenbergMarquardtStrategy::ComputeStep; (J^T J + D^2) y = J^T r replaces DENSE_QR on [J; D]
        if (!L.reuse_diag)
            for (int k = 0; k < 6; ++k) L.diag[k] = fmin(fmax(L.A[sym_idx(k, k)], min_diag), max_diag);
        double dd[6], y[6];
        for (int k = 0; k < 6; ++k) dd[k] = L.diag[k] / L.radius;
        bool valid = chol6_solve_packed(L.A, dd, L.gs, y);
        L.reuse_diag = 1;
        double mc = 0.0;
        if (valid) {
            for (int k = 0; k < 6; ++k) { y[k] = -y[k]; valid = valid && isfinite(y[k]); }
        }
        if (valid) {
            // model_cost_change = -(J s)'(r + J s / 2) = -s'g - s'As/2
            double sg = 0.0, sAs = 0.0;
            for (int r = 0; r < 6; ++r) {
                sg = fma(y[r], L.gs[r], sg);
                double row = 0.0;
                for (int c = 0; c < 6; ++c) row = fma(L.A[r <= c ? sym_idx(r, c) : sym_idx(c, r)], y[c], row);
                sAs = fma(y[r], row, sAs);
            }
            mc = -sg - 0.5 * sAs;
            valid = mc > 0.0;
        }
        if (!valid) {
            if (++L.n_invalid >= 5) { L.term = TERM_FAILURE; L.ctl = CTL_STOP; return; }
            L.radius *= 0.5; L.reuse_diag = 1; L.step_ok = 0;  // StepIsInvalid
            continue;
        }
        L.n_invalid = 0;
        L.model_change = mc;
        for (int k = 0; k < 6; ++k) L.xc[k] = L.x[k] + y[k] * L.scale[k];
        lm_set_eval_point(L, L.xc);
        L.ctl = CTL_CONTINUE;
        return;
    }
}

__device__ inline void quat_to_angle_axis(const double* q, double* aa) {
    const double s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    if (s2 > 0.0) {
        const double sn = sqrt(s2), cs = q[0];
        const double two_theta = 2.0 * ((cs < 0.0) ? atan2(-sn, -cs) : atan2(sn, cs));
        const double k = two_theta / sn;
        aa[0] = q[1] * k; aa[1] = q[2] * k; aa[2] = q[3] * k;
    } else {
        aa[0] = q[1] * 2.0; aa[1] = q[2] * 2.0; aa[2] = q[3] * 2.0;
    }
}
__device__ inline void angle_axis_to_quat(const double* aa, double* q) {
    const double th2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
    if (th2 > 0.0) {
        const double th = sqrt(th2), h = th * 0.5;
        double sn, cs;
        sincos(h, &sn, &cs);
        const double k = sn / th;
        q[0] = cs; q[1] = aa[0] * k; q[2] = aa[1] * k; q[3] = aa[2] * k;
    } else {
        q[0] = 1.0; q[1] = aa[0] * 0.5; q[2] = aa[1] * 0.5; q[3] = aa[2] * 0.5;
    }
}

// sqrt-information factor L (a = L00, b = L10, c = L11) of one correspondence, with the fp32/fp64
// rounding the reference's host-side prologue applies (cer_solver.py:37-40, test.py:54,95)
template <typename T>
__device__ __forceinline__ void load_sqrt_info(const lc_args& a, int b, int i, bool sanitize, double& la, double& lb, double& lc_) {
    const int64_t o = b * a.weights.stride[0] + i * a.weights.stride[1];
    const T* w = static_cast<const T*>(a.weights.ptr);
    auto get = [&](int64_t off) -> T {
        T v = w[off];
        if (sanitize) v = static_cast<T>(nan_to_num<T>(static_cast<double>(v)));
        return v;
    };
    if (a.weight_mode == LC_W_ICOV_DIAG) {
        la = static_cast<double>(static_cast<T>(sqrt(static_cast<T>(get(o)))));
        lc_ = static_cast<double>(static_cast<T>(sqrt(static_cast<T>(get(o + a.weights.stride[2])))));
        lb = 0.0;
    } else if (a.weight_mode == LC_W_INV_STD) {
        const T s0 = get(o), s1 = get(o + a.weights.stride[2]);
        const T i0 = s0 * s0, i1 = s1 * s1;  // inv_cov2d = inv_std ** 2 in T
        la = static_cast<double>(static_cast<T>(sqrt(i0)));
        lc_ = static_cast<double>(static_cast<T>(sqrt(i1)));
        lb = 0.0;
    } else if (a.weight_mode == LC_W_ICOV_FULL) {
        // torch.linalg.cholesky_ex on a 2x2 in T (lower): l00 = sqrt(i00), l10 = i10 / l00, l11 = sqrt(i11 - l10^2)
        const T i00 = get(o), i10 = get(o + a.weights.stride[2]), i11 = get(o + a.weights.stride[2] + a.weights.stride[3]);
        const T l00 = static_cast<T>(sqrt(i00));
        const T l10 = i10 / l00;
        const T l11 = static_cast<T>(sqrt(static_cast<T>(i11 - l10 * l10)));
        la = l00; lb = l10; lc_ = l11;
    } else {  // LC_W_SQRT_L
        la = get(o);
        lb = get(o + a.weights.stride[2]);
        lc_ = get(o + a.weights.stride[2] + a.weights.stride[3]);
    }
}

// One evaluation pass of the reprojection cost (ceres.cpp:30-55) at the point held in L.Rm/L.te:
// cost, J'^T J' and J'^T r in the left basis (J = J' blockdiag(Jl, I)).
template <typename T, int NT>
__device__ __forceinline__ void lm_eval_pass(const lc_args& a, PoseShared& s, int b, int n, bool sanitize) {
    const LmState& L = s.lm;
    double acc[28];
#pragma unroll
    for (int k = 0; k < 28; ++k) acc[k] = 0.0;
    const double k00 = s.K[0], k01 = s.K[1], k10 = s.K[3], k11 = s.K[4], cx = s.K[2], cy = s.K[5];
    for (int i = threadIdx.x; i < n; i += NT) {
        const int64_t o3 = b * a.pts3d.stride[0] + i * a.pts3d.stride[1];
        double X0 = ld<T>(a.pts3d, o3), X1 = ld<T>(a.pts3d, o3 + a.pts3d.stride[2]), X2 = ld<T>(a.pts3d, o3 + 2 * a.pts3d.stride[2]);
        const int64_t o2 = b * a.pts2d.stride[0] + i * a.pts2d.stride[1];
        double px = ld<T>(a.pts2d, o2), py = ld<T>(a.pts2d, o2 + a.pts2d.stride[2]);
        if (sanitize) {
            X0 = nan_to_num<T>(X0); X1 = nan_to_num<T>(X1); X2 = nan_to_num<T>(X2);
            px = nan_to_num<T>(px); py = nan_to_num<T>(py);
        }
        double la, lb, lc_;
        load_sqrt_info<T>(a, b, i, sanitize, la, lb, lc_);
        const double q0 = fma(L.Rm[0], X0, fma(L.Rm[1], X1, L.Rm[2] * X2));
        const double q1 = fma(L.Rm[3], X0, fma(L.Rm[4], X1, L.Rm[5] * X2));
        const double q2 = fma(L.Rm[6], X0, fma(L.Rm[7], X1, L.Rm[8] * X2));
        const double p0 = q0 + L.te[0], p1 = q1 + L.te[1], p2 = q2 + L.te[2];
        const double iz = 1.0 / p2;
        const double up = fma(p0, k00, p1 * k01) * iz, vp = fma(p0, k10, p1 * k11) * iz;
        const double du = up - (px - cx), dv = vp - (py - cy);
        const double r0 = fma(du, la, dv * lb), r1 = dv * lc_;
        // d(up,vp)/dp rows, then L^T applied: row0 = a D0 + b D1, row1 = c D1
        const double D0[3] = {k00 * iz, k01 * iz, -up * iz}, D1[3] = {k10 * iz, k11 * iz, -vp * iz};
        double J0[6], J1[6];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            J0[3 + c] = fma(la, D0[c], lb * D1[c]);
            J1[3 + c] = lc_ * D1[c];
        }
        J0[0] = fma(q1, J0[5], -q2 * J0[4]); J0[1] = fma(q2, J0[3], -q0 * J0[5]); J0[2] = fma(q0, J0[4], -q1 * J0[3]);
        J1[0] = fma(q1, J1[5], -q2 * J1[4]); J1[1] = fma(q2, J1[3], -q0 * J1[5]); J1[2] = fma(q0, J1[4], -q1 * J1[3]);
        acc_outer<0>(acc, 1.0, J0);
        acc_outer<0>(acc, 1.0, J1);
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[21 + c] = fma(J0[c], r0, fma(J1[c], r1, acc[21 + c]));
        acc[27] += 0.5 * fma(r0, r0, r1 * r1);
    }
    block_reduce<28, NT>(acc, s.red, s.fin);
}

// ---------------------------------------------------------------------------------------------
// the fused per-pose kernel
// ---------------------------------------------------------------------------------------------
template <typename T, int NT, int MODE>
__global__ void __launch_bounds__(NT) lc_pose_kernel(const lc_args a) {
    __shared__ PoseShared s;
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int n = a.n_points ? min(max(a.n_points[b], 0), a.N) : a.N;
    const bool sanitize = (a.flags & LC_FLAG_NAN_TO_NUM) != 0;

    // ---- pose constants ----
    if (tid < 9) {
        double v = ld<T>(a.K, b * a.K.stride[0] + (tid / 3) * a.K.stride[1] + (tid % 3) * a.K.stride[2]);
        s.K[tid] = (MODE & MODE_LM) && sanitize ? nan_to_num<T>(v) : v;
    } else if (tid < 16) {
        double v = ld<T>(a.pose, b * a.pose.stride[0] + (tid - 9) * a.pose.stride[1]);
        s.pose[tid - 9] = (MODE & MODE_LM) && sanitize ? nan_to_num<T>(v) : v;
    }
    if (MODE & MODE_LC) {
        for (int k = tid; k < 24; k += NT)
            s.bbox[k] = ld<T>(a.bbox, b * a.bbox.stride[0] + (k / 3) * a.bbox.stride[1] + (k % 3) * a.bbox.stride[2]);
    }
    __syncthreads();

    // =========================== LM solve ===========================
    if (MODE & MODE_LM) {
        LmState& L = s.lm;
        double* trace = a.trace ? a.trace + (int64_t)b * (a.max_iter + 2) * 4 : nullptr;
        bool solved = false;
        if (n >= 3) {
            if (tid == 0) {
                quat_to_angle_axis(s.pose, L.x);
                L.x[3] = s.pose[4]; L.x[4] = s.pose[5]; L.x[5] = s.pose[6];
                lm_set_eval_point(L, L.x);
            }
            __syncthreads();
            bool first = true;
            for (;;) {
                lm_eval_pass<T, NT>(a, s, b, n, sanitize);
                if (tid == 0)
                    lm_advance(L, s.fin, first, a.max_iter, a.function_tolerance, (a.flags & LC_FLAG_TOL_NEEDS_SUCCESS) != 0, trace);
                first = false;
                __syncthreads();
                if (L.ctl == CTL_STOP) break;
            }
            solved = L.term == TERM_CONVERGENCE;
        }
        if (tid == 0) {
            // ceres.cpp:134-144: state written back only when valid; cer_solver.py:51-52 keeps `start` otherwise
            if (solved) {
                double q[4];
                angle_axis_to_quat(L.x, q);
                for (int k = 0; k < 4; ++k) s.pose[k] = static_cast<double>(static_cast<T>(q[k]));
                for (int k = 0; k < 3; ++k) s.pose[4 + k] = static_cast<double>(static_cast<T>(L.x[3 + k]));
            }
            if (a.state.ptr)
                for (int k = 0; k < 7; ++k) st<T>(a.state, b * a.state.stride[0] + k * a.state.stride[1], s.pose[k]);
            if (a.radius.ptr) st<T>(a.radius, b * a.radius.stride[0], n >= 3 ? L.reported_radius : 1.0);
            if (a.invalid) a.invalid[b] = solved ? 0 : 1;
            if (a.iters) a.iters[b] = n >= 3 ? L.it : 0;
        }
        __syncthreads();
    }
    if (!(MODE & MODE_LC)) return;

    // =========================== LC loss forward ===========================
    if (tid == 0) {
        double qn;
        quat_to_R_ref(s.pose, s.R, &qn);
        // derivative of the reference's quaternion_to_matrix(q (x) dq) wrt the right perturbation: n * R_true
        for (int k = 0; k < 9; ++k) { const double id = (k % 4 == 0) ? 1.0 : 0.0; s.Rb[k] = qn * id + (s.R[k] - id); }
        s.t[0] = s.pose[4]; s.t[1] = s.pose[5]; s.t[2] = s.pose[6];
        s.flag = 0;
    }
    __syncthreads();

    const double Lmax = a.max_err_len;
    // pass 1: sum_i valid_i |ec_ia|, sum_i valid_i      (cov_mixed.py:28-31)
    {
        double acc[3] = {0.0, 0.0, 0.0};
        for (int i = tid; i < n; i += NT) {
            PointIn p; PointErr e;
            load_point<T>(a, b, i, false, p);
            point_error(s, p, Lmax, e);
            acc[0] = fma(p.valid, fabs(e.ec[0]), acc[0]);
            acc[1] = fma(p.valid, fabs(e.ec[1]), acc[1]);
            acc[2] += p.valid;
        }
        block_reduce<3, NT>(acc, s.red, s.fin);
    }
    const double vcnt = a.valid.ptr ? s.fin[2] : static_cast<double>(n);
    const double d0 = a.rel_thresh * (s.fin[0] / vcnt), d1 = a.rel_thresh * (s.fin[1] / vcnt);
    __syncthreads();  // fin is reused by the next reduction
    // pass 2: q_a = mean valid s^2 sigma    (cov_mixed.py:32-36)
    {
        double acc[2] = {0.0, 0.0};
        for (int i = tid; i < n; i += NT) {
            PointIn p; PointErr e;
            load_point<T>(a, b, i, true, p);
            point_error(s, p, Lmax, e);
            const double a0 = fabs(e.ec[0]), a1 = fabs(e.ec[1]);
            const double sg0 = a0 > d0 ? d0 * (2.0 * a0 - d0) : a0 * a0;
            const double sg1 = a1 > d1 ? d1 * (2.0 * a1 - d1) : a1 * a1;
            acc[0] = fma(p.valid * (p.s[0] * p.s[0]), sg0, acc[0]);
            acc[1] = fma(p.valid * (p.s[1] * p.s[1]), sg1, acc[1]);
        }
        block_reduce<2, NT>(acc, s.red, s.fin);
    }
    const double qw0 = (s.fin[0] / vcnt) * a.w_e_thresh, qw1 = (s.fin[1] / vcnt) * a.w_e_thresh;
    __syncthreads();
    // pass 3: H' = sum W J'J'^T, G' = sum W^2 sigma J'J'^T, b' = sum W ec J'   (left basis)
    {
        double acc[48];
#pragma unroll
        for (int k = 0; k < 48; ++k) acc[k] = 0.0;
        for (int i = tid; i < n; i += NT) {
            PointIn p; PointErr e;
            load_point<T>(a, b, i, true, p);
            point_error(s, p, Lmax, e);
            double J[2][6];
            point_jac_left(s, e.P, J);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const double dc = c ? d1 : d0, qw = c ? qw1 : qw0;
                const double av = fabs(e.ec[c]);
                const double sg = av > dc ? dc * (2.0 * av - dc) : av * av;
                const double del = sqrt(qw / (sg + 1e-6));
                const double sk = p.s[c];
                const double w = sk > del ? del * (2.0 * sk - del) : sk * sk;
                acc_outer<0>(acc, w, J[c]);
                acc_outer<21>(acc, w * w * sg, J[c]);
                const double wb = w * e.ec[c];
#pragma unroll
                for (int r = 0; r < 6; ++r) acc[42 + r] = fma(wb, J[c][r], acc[42 + r]);
            }
        }
        block_reduce<48, NT>(acc, s.red, s.fin);
    }

    // ---- 6x6 section (right basis, exactly the reference's quantities) ----
    for (int e = tid; e < 36; e += NT) {
        const int r = e / 6, c = e % 6;
        s.H[e] = tts_entry(s.fin, s.R, r, c);
        s.G[e] = tts_entry(s.fin + 21, s.R, r, c);
    }
    if (tid < 6) {
        const int r = tid;
        s.bv[r] = r < 3 ? s.R[r] * s.fin[42] + s.R[3 + r] * s.fin[43] + s.R[6 + r] * s.fin[44] : s.fin[42 + r];
    }
    // bbox corner Jacobian rows [Rb(-[c_j]x) | I]  (cov_mixed.py:52-65)
    if (tid < 24) {
        const int j = tid / 3, r = tid % 3;
        const double* c = s.bbox + 3 * j;
        const double nC[9] = {0, c[2], -c[1], -c[2], 0, c[0], c[1], -c[0], 0};
        for (int cc = 0; cc < 3; ++cc) {
            s.rows[tid * 6 + cc] = s.Rb[r * 3] * nC[cc] + s.Rb[r * 3 + 1] * nC[3 + cc] + s.Rb[r * 3 + 2] * nC[6 + cc];
            s.rows[tid * 6 + 3 + cc] = (r == cc) ? 1.0 : 0.0;
        }
    }
    __syncthreads();
    if (tid == 0) {
        // safe_cholesky: non-SPD -> identity (pnp_utils.py:140-167)
        double Hs[36];
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) Hs[r * 6 + c] = 0.5 * (s.H[r * 6 + c] + s.H[c * 6 + r]);
        if (chol6_inverse(Hs, s.C) != 0) {
            s.flag |= LC_ST_HESS_NOT_SPD;
            for (int k = 0; k < 36; ++k) s.C[k] = (k % 7 == 0) ? 1.0 : 0.0;
        }
    }
    __syncthreads();
    mm6_par<NT>(s.C, s.G, s.T1);
    if (tid < 6) {
        double v = 0.0;
        for (int k = 0; k < 6; ++k) v = fma(s.C[tid * 6 + k], s.bv[k], v);
        s.dth[tid] = v;
    }
    __syncthreads();
    mm6_par<NT>(s.T1, s.C, s.M);  // M = C G C
    __syncthreads();
    if (tid < 24) {
        const double* row = s.rows + tid * 6;
        double vc = 0.0, vm = 0.0, uu = 0.0;
        for (int r = 0; r < 6; ++r) {
            double wc = 0.0, wm = 0.0;
            for (int c = 0; c < 6; ++c) { wc = fma(s.C[r * 6 + c], row[c], wc); wm = fma(s.M[r * 6 + c], row[c], wm); }
            vc = fma(row[r], wc, vc); vm = fma(row[r], wm, vm); uu = fma(row[r], s.dth[r], uu);
        }
        s.vC[tid] = vc; s.vM[tid] = vm; s.u[tid] = uu;
    }
    __syncthreads();
    if (tid == 0) {
        bool goodC = true, goodM = true;
        for (int k = 0; k < 24; ++k) { goodC = goodC && (s.vC[k] > 0.0); goodM = goodM && (s.vM[k] > 0.0); }
        double prior = 0.0, cov_err = 0.0, lin = 0.0, sC[8], sM[8], un[8];
        for (int j = 0; j < 8; ++j) {
            sC[j] = s.vC[3 * j] + s.vC[3 * j + 1] + s.vC[3 * j + 2];
            sM[j] = s.vM[3 * j] + s.vM[3 * j + 1] + s.vM[3 * j + 2];
            un[j] = sqrt(s.u[3 * j] * s.u[3 * j] + s.u[3 * j + 1] * s.u[3 * j + 1] + s.u[3 * j + 2] * s.u[3 * j + 2]);
            prior += sqrt(goodC ? sC[j] : 1.0);
            cov_err += sqrt(goodM ? sM[j] : 1.0);
            lin += un[j];
        }
        prior *= 0.125; cov_err *= 0.125; lin *= 0.125;
        const double loss = log(prior) + 0.5 * (cov_err + lin) / prior;
        if (a.loss.ptr) st<T>(a.loss, b * a.loss.stride[0], loss);
        if (!goodC) s.flag |= LC_ST_PRIOR_NOT_GOOD;
        if (!goodM) s.flag |= LC_ST_COV_NOT_GOOD;
        if (a.lc_flags) a.lc_flags[b] = s.flag;
        const double go = a.grad_scale * (a.grad_out.ptr ? ld<T>(a.grad_out, b * a.grad_out.stride[0]) : 1.0);
        const double g_p = go * (1.0 / prior - 0.5 * (cov_err + lin) / (prior * prior));
        const double g_c = go * 0.5 / prior;
        for (int j = 0; j < 8; ++j) {
            s.wC[j] = goodC ? g_p / (16.0 * sqrt(sC[j])) : 0.0;
            s.wM[j] = goodM ? g_c / (16.0 * sqrt(sM[j])) : 0.0;
            s.wU[j] = un[j] > 0.0 ? g_c * 0.125 / un[j] : 0.0;
        }
    }
    if (a.cov.ptr)
        for (int e = tid; e < 36; e += NT) st<T>(a.cov, b * a.cov.stride[0] + (e / 6) * a.cov.stride[1] + (e % 6) * a.cov.stride[2], s.C[e]);
    if (a.update_cov.ptr)
        for (int e = tid; e < 36; e += NT)
            st<T>(a.update_cov, b * a.update_cov.stride[0] + (e / 6) * a.update_cov.stride[1] + (e % 6) * a.update_cov.stride[2],
                  0.5 * (s.M[e] + s.M[(e % 6) * 6 + e / 6]));
    const bool want_grads = a.g_pts3d.ptr || a.g_pts2d.ptr || a.g_weights.ptr;
    if (!want_grads) return;
    __syncthreads();

    // =========================== reverse 6x6 section ===========================
    for (int e = tid; e < 36; e += NT) {
        const int r = e / 6, c = e % 6;
        double cb = 0.0, mb = 0.0;
        for (int k = 0; k < 24; ++k) {
            const double qq = s.rows[k * 6 + r] * s.rows[k * 6 + c];
            cb = fma(s.wC[k / 3], qq, cb);
            mb = fma(s.wM[k / 3], qq, mb);
        }
        s.Cbar[e] = cb; s.Mbar[e] = mb;
    }
    if (tid < 6) {
        double v = 0.0;
        for (int k = 0; k < 24; ++k) v = fma(s.wU[k / 3] * s.u[k], s.rows[k * 6 + tid], v);
        s.dthbar[tid] = v;
    }
    __syncthreads();
    mm6_par<NT>(s.C, s.Mbar, s.T1);   // C Mbar
    mm6_par<NT>(s.Mbar, s.C, s.T2);   // Mbar C
    if (tid < 6) {
        double v = 0.0;
        for (int k = 0; k < 6; ++k) v = fma(s.C[tid * 6 + k], s.dthbar[k], v);
        s.bbar[tid] = v;
    }
    __syncthreads();
    mm6_par<NT>(s.T1, s.C, s.Gbar);   // Gbar = C Mbar C
    mm6_par<NT>(s.T2, s.G, s.Hbar);   // (Mbar C G), staged in Hbar
    __syncthreads();
    for (int e = tid; e < 36; e += NT) {
        const int r = e / 6, c = e % 6;
        s.T1[e] = s.Cbar[e] + s.Hbar[e] + s.Hbar[c * 6 + r] + s.dthbar[r] * s.bv[c];  // Cbar total
    }
    __syncthreads();
    mm6_par<NT>(s.C, s.T1, s.T2);
    __syncthreads();
    mm6_par<NT>(s.T2, s.C, s.Hbar);   // -Hbar
    __syncthreads();
    if (s.flag & LC_ST_HESS_NOT_SPD)
        for (int e = tid; e < 36; e += NT) s.Hbar[e] = 0.0;   // torch.where(cond, eye, H): no gradient into H
    __syncthreads();
    // to the left basis, symmetrised and packed with doubled off-diagonals: J^T S J = sum_{i<=j} c_ij J'_i J'_j
    for (int e = tid; e < kSym * 2; e += NT) {
        const bool isG = e >= kSym;
        const int k = isG ? e - kSym : e;
        int r = 0, c = 0;
        for (int i = 0, kk = 0; i < 6; ++i)
            for (int j = i; j < 6; ++j, ++kk)
                if (kk == k) { r = i; c = j; }
        const double* Msrc = isG ? s.Gbar : s.Hbar;
        const double sgn = isG ? 1.0 : -1.0;
        double v = sgn * tmt_entry(Msrc, s.R, r, c);
        if (r != c) v += sgn * tmt_entry(Msrc, s.R, c, r);
        (isG ? s.cGL : s.cHL)[k] = v;
    }
    if (tid < 6) {
        const int r = tid;  // bL = T bbar, T = blockdiag(R, I)
        s.bL[r] = r < 3 ? s.R[r * 3] * s.bbar[0] + s.R[r * 3 + 1] * s.bbar[1] + s.R[r * 3 + 2] * s.bbar[2] : s.bbar[r];
    }
    __syncthreads();

    // pass 4: per-coordinate adjoints  
    for (int i = tid; i < n; i += NT) {
        PointIn p; PointErr e;
        load_point<T>(a, b, i, true, p);
        point_error(s, p, Lmax, e);
        double J[2][6];
        point_jac_left(s, e.P, J);
        double ecb[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const double dc = c ? d1 : d0, qw = c ? qw1 : qw0;
            const double av = fabs(e.ec[c]);
            const bool big = av > dc;
            const double sg = big ? dc * (2.0 * av - dc) : av * av;
            const double del = sqrt(qw / (sg + 1e-6));
            const double sk = p.s[c];
            const bool wbig = sk > del;
            const double w = wbig ? del * (2.0 * sk - del) : sk * sk;
            double qh = 0.0, qg = 0.0, lb = 0.0;
            int k = 0;
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                lb = fma(J[c][r], s.bL[r], lb);
#pragma unroll
                for (int cc = r; cc < 6; ++cc) {
                    const double pp = J[c][r] * J[c][cc];
                    qh = fma(s.cHL[k], pp, qh);
                    qg = fma(s.cGL[k], pp, qg);
                    ++k;
                }
            }
            const double Wbar = qh + 2.0 * w * sg * qg + e.ec[c] * lb;
            const double sigbar = w * w * qg;
            if (a.g_weights.ptr)
                st<T>(a.g_weights, b * a.g_weights.stride[0] + i * a.g_weights.stride[1] + c * a.g_weights.stride[2],
                      Wbar * (wbig ? 2.0 * del : 2.0 * sk));
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

// ---------------------------------------------------------------------------------------------
// weighted_pnp_jac_wrt_pts2d (pnp_auto.py:111-135): jac (B,6,N,2) = W_k C J_k, cov = H^-1
// and its double-backward w.r.t. the weights.
// ---------------------------------------------------------------------------------------------
struct JacShared {
    double K[9], pose[7], R[9], t[3];
    double red[kMaxWarps * 57];
    double fin[57];
    double H[36], C[36], CT[36], Z[36], T1[36], T2[36], Hbar[36];
    double cHL[kSym];
    int flag;
};

template <typename T, int NT>
__device__ __forceinline__ void jac_setup(const lc_args& a, JacShared& s, int b) {
    const int tid = threadIdx.x;
    if (tid < 9) s.K[tid] = ld<T>(a.K, b * a.K.stride[0] + (tid / 3) * a.K.stride[1] + (tid % 3) * a.K.stride[2]);
    else if (tid < 16) s.pose[tid - 9] = ld<T>(a.pose, b * a.pose.stride[0] + (tid - 9) * a.pose.stride[1]);
    __syncthreads();
    if (tid == 0) {
        double qn;
        quat_to_R_ref(s.pose, s.R, &qn);
        s.t[0] = s.pose[4]; s.t[1] = s.pose[5]; s.t[2] = s.pose[6];
        s.flag = 0;
    }
    __syncthreads();
}

// left-basis Jacobian rows from raw inputs (no projection error needed here)
template <typename T>
__device__ __forceinline__ void jac_point(const lc_args& a, const double* R, const double* t, const double* K, int b, int i,
                                          double (&J)[2][6], double (&w)[2]) {
    const int64_t o3 = b * a.pts3d.stride[0] + i * a.pts3d.stride[1];
    const double X0 = ld<T>(a.pts3d, o3), X1 = ld<T>(a.pts3d, o3 + a.pts3d.stride[2]), X2 = ld<T>(a.pts3d, o3 + 2 * a.pts3d.stride[2]);
    const double q0 = fma(R[0], X0, fma(R[1], X1, R[2] * X2));
    const double q1 = fma(R[3], X0, fma(R[4], X1, R[5] * X2));
    const double q2 = fma(R[6], X0, fma(R[7], X1, R[8] * X2));
    const double P0 = q0 + t[0], P1 = q1 + t[1], P2 = q2 + t[2];
    const double iz = 1.0 / P2, u0 = P0 * iz, v0 = P1 * iz;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const double k0 = K[c * 3], k1 = K[c * 3 + 1];
        const double d0 = k0 * iz, d1 = k1 * iz, d2 = -fma(k0, u0, k1 * v0) * iz;
        J[c][0] = fma(q1, d2, -q2 * d1);
        J[c][1] = fma(q2, d0, -q0 * d2);
        J[c][2] = fma(q0, d1, -q1 * d0);
        J[c][3] = d0; J[c][4] = d1; J[c][5] = d2;
    }
    const int64_t ow = b * a.weights.stride[0] + i * a.weights.stride[1];
    w[0] = ld<T>(a.weights, ow);
    w[1] = ld<T>(a.weights, ow + a.weights.stride[2]);
}

template <typename T, int NT>
__device__ __forceinline__ void jac_hessian(const lc_args& a, JacShared& s, int b, int n) {
    const int tid = threadIdx.x;
    double acc[21];
#pragma unroll
    for (int k = 0; k < 21; ++k) acc[k] = 0.0;
    for (int i = tid; i < n; i += NT) {
        double J[2][6], w[2];
        jac_point<T>(a, s.R, s.t, s.K, b, i, J, w);
        acc_outer<0>(acc, w[0], J[0]);
        acc_outer<0>(acc, w[1], J[1]);
    }
    block_reduce<21, NT>(acc, s.red, s.fin);
    for (int e = tid; e < 36; e += NT) s.H[e] = tts_entry(s.fin, s.R, e / 6, e % 6);
    __syncthreads();
    if (tid == 0) {
        double Hs[36];
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) Hs[r * 6 + c] = 0.5 * (s.H[r * 6 + c] + s.H[c * 6 + r]);
        if (chol6_inverse(Hs, s.C) != 0) {
            s.flag |= LC_ST_HESS_NOT_SPD;
            for (int k = 0; k < 36; ++k) s.C[k] = (k % 7 == 0) ? 1.0 : 0.0;
        }
    }
    __syncthreads();
    // CT = C T^T (T = blockdiag(R, I)): A_k = W_k C J_k = W_k CT J'_k
    for (int e = tid; e < 36; e += NT) {
        const int r = e / 6, c = e % 6;
        s.CT[e] = c < 3 ? s.C[r * 6] * s.R[c * 3] + s.C[r * 6 + 1] * s.R[c * 3 + 1] + s.C[r * 6 + 2] * s.R[c * 3 + 2] : s.C[e];
    }
    __syncthreads();
}

template <typename T, int NT>
__global__ void __launch_bounds__(NT) lc_jac_kernel(const lc_args a) {
    __shared__ JacShared s;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = a.N;
    jac_setup<T, NT>(a, s, b);
    jac_hessian<T, NT>(a, s, b, n);
    if (a.cov.ptr)
        for (int e = tid; e < 36; e += NT) st<T>(a.cov, b * a.cov.stride[0] + (e / 6) * a.cov.stride[1] + (e % 6) * a.cov.stride[2], s.C[e]);
    if (a.lc_flags && tid == 0) a.lc_flags[b] = s.flag;
    if (!a.jac.ptr) return;
    for (int i = tid; i < n; i += NT) {
        double J[2][6], w[2];
        jac_point<T>(a, s.R, s.t, s.K, b, i, J, w);
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) v = fma(s.CT[r * 6 + k], J[c][k], v);
                st<T>(a.jac, b * a.jac.stride[0] + r * a.jac.stride[1] + i * a.jac.stride[2] + c * a.jac.stride[3], w[c] * v);
            }
        }
    }
}

// gW_k = gA_k^T C J_k + J_k^T Hbar J_k,  Hbar = -C (gC + sum_k W_k gA_k J_k^T) C
template <typename T, int NT>
__global__ void __launch_bounds__(NT) lc_jac_bwd_kernel(const lc_args a) {
    __shared__ JacShared s;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = a.N;
    jac_setup<T, NT>(a, s, b);
    jac_hessian<T, NT>(a, s, b, n);
    // Z' = sum_k W_k gA_k J'_k^T  (6x6, columns in the left basis)
    {
        double acc[36];
#pragma unroll
        for (int k = 0; k < 36; ++k) acc[k] = 0.0;
        for (int i = tid; i < n; i += NT) {
            double J[2][6], w[2];
            jac_point<T>(a, s.R, s.t, s.K, b, i, J, w);
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    const double ga = w[c] * ld<T>(a.g_jac, b * a.g_jac.stride[0] + r * a.g_jac.stride[1] + i * a.g_jac.stride[2] + c * a.g_jac.stride[3]);
#pragma unroll
                    for (int k = 0; k < 6; ++k) acc[r * 6 + k] = fma(ga, J[c][k], acc[r * 6 + k]);
                }
        }
        block_reduce<36, NT>(acc, s.red, s.fin);
    }
    // Z = Z' T (right basis columns), Cbar = gC + Z
    for (int e = tid; e < 36; e += NT) {
        const int r = e / 6, c = e % 6;
        double z = c < 3 ? s.fin[r * 6] * s.R[c] + s.fin[r * 6 + 1] * s.R[3 + c] + s.fin[r * 6 + 2] * s.R[6 + c] : s.fin[e];
        if (a.g_cov.ptr) z += ld<T>(a.g_cov, b * a.g_cov.stride[0] + r * a.g_cov.stride[1] + c * a.g_cov.stride[2]);
        s.Z[e] = z;
    }
    __syncthreads();
    mm6_par<NT>(s.C, s.Z, s.T1);
    __syncthreads();
    mm6_par<NT>(s.T1, s.C, s.Hbar);  // -Hbar
    __syncthreads();
    if (s.flag & LC_ST_HESS_NOT_SPD)
        for (int e = tid; e < 36; e += NT) s.Hbar[e] = 0.0;
    __syncthreads();
    for (int k = tid; k < kSym; k += NT) {
        int r = 0, c = 0;
        for (int i = 0, kk = 0; i < 6; ++i)
            for (int j = i; j < 6; ++j, ++kk)
                if (kk == k) { r = i; c = j; }
        double v = -tmt_entry(s.Hbar, s.R, r, c);
        if (r != c) v -= tmt_entry(s.Hbar, s.R, c, r);
        s.cHL[k] = v;
    }
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
        double J[2][6], w[2];
        jac_point<T>(a, s.R, s.t, s.K, b, i, J, w);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            double qh = 0.0;
            int k = 0;
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int cc = r; cc < 6; ++cc) { qh = fma(s.cHL[k], J[c][r] * J[c][cc], qh); ++k; }
            double lin = 0.0;
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                double v = 0.0;
#pragma unroll
                for (int kk = 0; kk < 6; ++kk) v = fma(s.CT[r * 6 + kk], J[c][kk], v);
                lin = fma(ld<T>(a.g_jac, b * a.g_jac.stride[0] + r * a.g_jac.stride[1] + i * a.g_jac.stride[2] + c * a.g_jac.stride[3]), v, lin);
            }
            st<T>(a.g_weights, b * a.g_weights.stride[0] + i * a.g_weights.stride[1] + c * a.g_weights.stride[2], qh + lin);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side: validation + dispatch
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[256] = "";
static thread_local int g_launches = 0;

static int fail(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

static int block_threads_for(int n) {
    if (n <= 48) return 32;
    if (n <= 192) return 64;
    if (n <= 1536) return 128;
    return 256;
}

template <typename T, int MODE>
static int launch_pose(const lc_args& a, cudaStream_t st) {
    const int nt = block_threads_for(a.N);
    switch (nt) {
        case 32: lc_pose_kernel<T, 32, MODE><<<a.B, 32, 0, st>>>(a); break;
        case 64: lc_pose_kernel<T, 64, MODE><<<a.B, 64, 0, st>>>(a); break;
        case 128: lc_pose_kernel<T, 128, MODE><<<a.B, 128, 0, st>>>(a); break;
        default: lc_pose_kernel<T, 256, MODE><<<a.B, 256, 0, st>>>(a); break;
    }
    ++g_launches;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(static_cast<int>(e), cudaGetErrorString(e));
    return LC_OK;
}

template <int MODE>
static int dispatch_pose(const lc_args* a, void* stream) {
    g_launches = 0;
    if (!a) return fail(LC_E_NULL, "args is NULL");
    if (a->abi_version != LC_B200_ABI_VERSION) return fail(LC_E_BADARG, "abi_version mismatch");
    if (a->B < 0 || a->N < 0) return fail(LC_E_BADARG, "B and N must be non-negative");
    if (!a->K.ptr || !a->pose.ptr || !a->pts3d.ptr || !a->pts2d.ptr || !a->weights.ptr)
        if (a->B > 0 && a->N > 0) return fail(LC_E_NULL, "K, pose, pts3d, pts2d and weights are required");
    if ((MODE & MODE_LC) && !a->bbox.ptr && a->B > 0) return fail(LC_E_NULL, "bbox is required for the loss");
    if ((MODE & MODE_LM) && a->max_iter < 0) return fail(LC_E_BADARG, "max_iter must be >= 0");
    if (a->B == 0) return LC_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (a->dtype == LC_F32) return launch_pose<float, MODE>(*a, st);
    if (a->dtype == LC_F64) return launch_pose<double, MODE>(*a, st);
    return fail(LC_E_DTYPE, "dtype must be LC_F32 or LC_F64");
}

template <typename T, bool BWD>
static int launch_jac(const lc_args& a, cudaStream_t st) {
    const int nt = block_threads_for(a.N);
#define LC_JAC_LAUNCH(NT_)                                              \
    if (BWD) lc_jac_bwd_kernel<T, NT_><<<a.B, NT_, 0, st>>>(a);         \
    else lc_jac_kernel<T, NT_><<<a.B, NT_, 0, st>>>(a)
    switch (nt) {
        case 32: LC_JAC_LAUNCH(32); break;
        case 64: LC_JAC_LAUNCH(64); break;
        case 128: LC_JAC_LAUNCH(128); break;
        default: LC_JAC_LAUNCH(256); break;
    }
#undef LC_JAC_LAUNCH
    ++g_launches;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(static_cast<int>(e), cudaGetErrorString(e));
    return LC_OK;
}

template <bool BWD>
static int dispatch_jac(const lc_args* a, void* stream) {
    g_launches = 0;
    if (!a) return fail(LC_E_NULL, "args is NULL");
    if (a->abi_version != LC_B200_ABI_VERSION) return fail(LC_E_BADARG, "abi_version mismatch");
    if (a->B < 0 || a->N < 0) return fail(LC_E_BADARG, "B and N must be non-negative");
    if (a->B == 0) return LC_OK;
    if (!a->K.ptr || !a->pose.ptr || !a->pts3d.ptr || !a->weights.ptr) return fail(LC_E_NULL, "K, pose, pts3d and weights are required");
    if (BWD && (!a->g_jac.ptr || !a->g_weights.ptr)) return fail(LC_E_NULL, "g_jac and g_weights are required");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (a->dtype == LC_F32) return launch_jac<float, BWD>(*a, st);
    if (a->dtype == LC_F64) return launch_jac<double, BWD>(*a, st);
    return fail(LC_E_DTYPE, "dtype must be LC_F32 or LC_F64");
}

}  // namespace lc

extern "C" {

int lc_b200_abi_version(void) { return LC_B200_ABI_VERSION; }
const char* lc_b200_last_error(void) { return lc::g_err; }
int lc_b200_last_launch_count(void) { return lc::g_launches; }

int lc_b200_lm_solve(const lc_args* a, void* stream) { return lc::dispatch_pose<lc::MODE_LM>(a, stream); }
int lc_b200_loss_fwd_bwd(const lc_args* a, void* stream) { return lc::dispatch_pose<lc::MODE_LC>(a, stream); }
int lc_b200_solve_loss(const lc_args* a, void* stream) { return lc::dispatch_pose<lc::MODE_LM | lc::MODE_LC>(a, stream); }
int lc_b200_pnp_jac_cov(const lc_args* a, void* stream) { return lc::dispatch_jac<false>(a, stream); }
int lc_b200_pnp_jac_cov_bwd(const lc_args* a, void* stream) { return lc::dispatch_jac<true>(a, stream); }

}  // extern "C"
