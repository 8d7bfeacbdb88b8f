// Per-point fp64 building blocks of the streaming kernels (lc_stream.cu: one CTA per pose; lc_tiny.cu: one thread per pose):
// strided loads, project_apply + clamp_error, left-basis Jacobian rows, the sqrt-information factor of one correspondence and
// the per-point contribution to the LM normal equations.
#pragma once

#include "lc_pose.cuh"

namespace lc {

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

template <class PS>
__device__ __forceinline__ void point_error(const PS& s, const PointIn& p, double Lmax, PointErr& e) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
        e.P[r] = fma(s.R[r * 3], p.X[0], fma(s.R[r * 3 + 1], p.X[1], fma(s.R[r * 3 + 2], p.X[2], s.t[r])));
    double KP[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) KP[r] = fma(s.K[r * 3], e.P[0], fma(s.K[r * 3 + 1], e.P[1], s.K[r * 3 + 2] * e.P[2]));
    e.z_active = KP[2] >= 0.1;
    e.zc = KP[2] > 0.1 ? KP[2] : 0.1;
    const double iz = fast_rcp(e.zc);
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
template <class PS>
__device__ __forceinline__ void point_jac_left(const PS& s, const double* P, double (&J)[2][6]) {
    const double iz = fast_rcp(P[2]);
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

// Contribution of correspondence i to the LM sums at the evaluation point (Rm, te): acc = [J'^T J' packed (21), J'^T r (6),
// cost] in the left basis (J = J' blockdiag(Jl, I)); ceres.cpp:30-55.  Kc = {k00, k01, k10, k11, cx, cy}.
template <typename T, bool JAC>
__device__ __forceinline__ void lm_point_accum(const lc_args& a, const double* Kc, const double* Rm, const double* te, int b, int i, bool sanitize,
                                               double (&acc)[28]) {
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
    const double q0 = fma(Rm[0], X0, fma(Rm[1], X1, Rm[2] * X2));
    const double q1 = fma(Rm[3], X0, fma(Rm[4], X1, Rm[5] * X2));
    const double q2 = fma(Rm[6], X0, fma(Rm[7], X1, Rm[8] * X2));
    const double p0 = q0 + te[0], p1 = q1 + te[1], p2 = q2 + te[2];
    const double iz = fast_rcp(p2);
    const double up = fma(p0, Kc[0], p1 * Kc[1]) * iz, vp = fma(p0, Kc[2], p1 * Kc[3]) * iz;
    const double du = up - (px - Kc[4]), dv = vp - (py - Kc[5]);
    const double r0 = fma(du, la, dv * lb), r1 = dv * lc_;
    acc[27] += 0.5 * fma(r0, r0, r1 * r1);
    if (!JAC) return;
    // d(up,vp)/dp rows, then L^T applied: row0 = a D0 + b D1, row1 = c D1
    const double D0[3] = {Kc[0] * iz, Kc[1] * iz, -up * iz}, D1[3] = {Kc[2] * iz, Kc[3] * iz, -vp * iz};
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
}

}  // namespace lc
