// lc_b200 — streaming sm_100a kernels for the LC hot path (generic fallback + small-N path).
//
// One CTA per pose; the pose's N correspondences are re-streamed from global memory (L1/L2 resident between
// passes) with arbitrary element strides and fp32 or fp64 I/O; all arithmetic is fp64.  This path takes every
// shape the ABI allows (any N, full 2x2 weights, fp64 tensors).  Large fp32 batches with diagonal weights go
// to the shared-memory resident kernel in lc_resident.cu instead (see lc_abi.cu for the dispatch rule).
//
//   lc_pose_kernel<T, NT, MODE>
//     MODE & 1  LM   : Ceres-faithful Levenberg-Marquardt (replaces ceres.cpp:72-145)
//     MODE & 2  LC   : Loss_cov_mixed forward + backward (replaces cov_mixed.py:100-150)
//   lc_jac_kernel / lc_jac_bwd_kernel : weighted_pnp_jac_wrt_pts2d forward / double-backward
//
// Math reference: SURVEY.md §8a ("LC math") and §8c (solver spec); oracle/*.c are the CPU checkers.
#include <cstdio>

#include "lc_point.cuh"

namespace lc {

// One evaluation pass of the reprojection cost (ceres.cpp:30-55) at the point held in L.Rm/L.te:
// cost, J'^T J' and J'^T r in the left basis (J = J' blockdiag(Jl, I)).
template <typename T, int NT, bool JAC>
__device__ __forceinline__ void lm_eval_pass(const lc_args& a, PoseShared& s, int b, int n, bool sanitize) {
    const LmState& L = s.lm;
    double acc[28];
#pragma unroll
    for (int k = 0; k < 28; ++k) acc[k] = 0.0;
    const double Kc[6] = {s.K[0], s.K[1], s.K[3], s.K[4], s.K[2], s.K[5]};
    for (int i = threadIdx.x; i < n; i += NT) lm_point_accum<T, JAC>(a, Kc, L.Rm, L.te, b, i, sanitize, acc);
    block_reduce<28, NT>(acc, s.red, s.fin);
}

// ---------------------------------------------------------------------------------------------
// the fused per-pose kernel (streaming)
// ---------------------------------------------------------------------------------------------
template <typename T, int NT, int MODE>
__global__ void __launch_bounds__(NT, (NT <= 64) ? (1024 / NT / 2) : 1) lc_pose_kernel(const lc_args a, int n_skip_le) {
    __shared__ PoseShared s;
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int n = a.n_points ? min(max(a.n_points[b], 0), a.N) : a.N;
    if (n <= n_skip_le) return;   // ragged batch split by n_points: this pose was handled by the resident launch (lc_abi.cu)
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
                L.ctl = CTL_EVAL_FULL;
            }
            __syncthreads();
            bool first = true;
            for (;;) {
                const int kind = L.ctl;
                if (kind == CTL_EVAL_COST) lm_eval_pass<T, NT, false>(a, s, b, n, sanitize);
                else lm_eval_pass<T, NT, true>(a, s, b, n, sanitize);
                if (tid == 0)
                    lm_advance(L, s.fin, kind, first, a.max_iter, a.function_tolerance, (a.flags & LC_FLAG_TOL_NEEDS_SUCCESS) != 0, trace);
                first = false;
                __syncthreads();
                if (L.ctl == CTL_STOP) break;
            }
            solved = L.term == TERM_CONVERGENCE;
        }
        if (tid == 0) lm_write_result<T>(a, s, b, n, solved);
        __syncthreads();
    }
    if (!(MODE & MODE_LC)) return;

    // =========================== LC loss forward ===========================
    if (tid == 0) lc_pose_setup(s, false);
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
    lc_six_forward<T, NT>(a, s, b);
    const bool want_grads = a.g_pts3d.ptr || a.g_pts2d.ptr || a.g_weights.ptr;
    if (!want_grads) return;
    lc_six_backward<NT>(s, (a.flags & LC_FLAG_COV_2D) ? 2 : 3);

    // gradient slots of the padding beyond n_points (ragged batches): defined, zero
    for (int i = n + tid; i < a.N; i += NT) {
        for (int c = 0; c < 2; ++c) {
            if (a.g_weights.ptr) st<T>(a.g_weights, b * a.g_weights.stride[0] + i * a.g_weights.stride[1] + c * a.g_weights.stride[2], 0.0);
            if (a.g_pts2d.ptr) st<T>(a.g_pts2d, b * a.g_pts2d.stride[0] + i * a.g_pts2d.stride[1] + c * a.g_pts2d.stride[2], 0.0);
        }
        if (a.g_pts3d.ptr)
            for (int c = 0; c < 3; ++c) st<T>(a.g_pts3d, b * a.g_pts3d.stride[0] + i * a.g_pts3d.stride[1] + c * a.g_pts3d.stride[2], 0.0);
    }
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

// Second-order term of hessian_6d_elem (pnp_auto.py:59-83): the reference's per-coordinate Hessian is
// J J^T + r * d2r/d(delta)^2 with the exact exponential-map second derivative.  In the left basis
//   d2r_a = -(1/z) (J'_a Z^T + Z J'_a^T) + blockdiag(S(D_a), 0),   Z = (q1, -q0, 0, 0, 0, 1),
//   S(D)_ij = (D_i q_j + D_j q_i)/2 - delta_ij (D . q)            (i, j < 3),  D_a = J'_a[3:6]
// (from d2 pi_a = -(1/z)(D e_z^T + e_z D^T) and d2(Exp(w) q)/dw_i dw_j = (e_i q_j + e_j q_i)/2 - delta_ij q).
// Returns the packed-upper entry (i, j), i <= j.
struct PointSecond { double q[3], iz, r[2]; };
__device__ __forceinline__ double second_entry(const double (&J)[6], const PointSecond& p, int i, int j) {
    const double Z[6] = {p.q[1], -p.q[0], 0.0, 0.0, 0.0, 1.0};
    double v = -p.iz * (J[i] * Z[j] + Z[i] * J[j]);
    if (i < 3 && j < 3) {
        v += 0.5 * (J[3 + i] * p.q[j] + J[3 + j] * p.q[i]);
        if (i == j) v -= J[3] * p.q[0] + J[4] * p.q[1] + J[5] * p.q[2];
    }
    return v;
}

// left-basis Jacobian rows from raw inputs (no projection error needed here)
template <typename T>
__device__ __forceinline__ void jac_point(const lc_args& a, const double* R, const double* t, const double* K, int b, int i,
                                          double (&J)[2][6], double (&w)[2], PointSecond* sec = nullptr) {
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
    if (sec) {
        // r = K[:2,:2] uv0 + K[:2,2] - pts2d   (residual_with_jac6d, pnp_auto.py:43-48)
        const int64_t o2 = b * a.pts2d.stride[0] + i * a.pts2d.stride[1];
        sec->q[0] = q0; sec->q[1] = q1; sec->q[2] = q2; sec->iz = iz;
        sec->r[0] = fma(K[0], u0, fma(K[1], v0, K[2])) - ld<T>(a.pts2d, o2);
        sec->r[1] = fma(K[3], u0, fma(K[4], v0, K[5])) - ld<T>(a.pts2d, o2 + a.pts2d.stride[2]);
    }
}

template <typename T, int NT>
__device__ __forceinline__ void jac_hessian(const lc_args& a, JacShared& s, int b, int n) {
    const int tid = threadIdx.x;
    double acc[21];
#pragma unroll
    for (int k = 0; k < 21; ++k) acc[k] = 0.0;
    const bool exact = (a.flags & LC_FLAG_EXACT_HESSIAN) && a.pts2d.ptr;
    for (int i = tid; i < n; i += NT) {
        double J[2][6], w[2];
        PointSecond sec;
        jac_point<T>(a, s.R, s.t, s.K, b, i, J, w, exact ? &sec : nullptr);
        acc_outer<0>(acc, w[0], J[0]);
        acc_outer<0>(acc, w[1], J[1]);
        if (exact) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const double wr = w[c] * sec.r[c];
                int k = 0;
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int cc = r; cc < 6; ++cc) { acc[k] = fma(wr, second_entry(J[c], sec, r, cc), acc[k]); ++k; }
            }
        }
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
    const bool exact = (a.flags & LC_FLAG_EXACT_HESSIAN) && a.pts2d.ptr;
    for (int i = tid; i < n; i += NT) {
        double J[2][6], w[2];
        PointSecond sec;
        jac_point<T>(a, s.R, s.t, s.K, b, i, J, w, exact ? &sec : nullptr);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            double qh = 0.0;
            int k = 0;
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int cc = r; cc < 6; ++cc) {
                    double hk = J[c][r] * J[c][cc];
                    if (exact) hk = fma(sec.r[c], second_entry(J[c], sec, r, cc), hk);
                    qh = fma(s.cHL[k], hk, qh);
                    ++k;
                }
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
// host side: launchers
// ---------------------------------------------------------------------------------------------
static int stream_threads_for(int n) {
    if (n <= 48) return 32;
    if (n <= 192) return 64;
    if (n <= 1536) return 128;
    return 256;
}

template <typename T, int MODE>
static int launch_pose_t(const lc_args& a, cudaStream_t st, int skip) {
    switch (stream_threads_for(a.N)) {
        case 32: lc_pose_kernel<T, 32, MODE><<<a.B, 32, 0, st>>>(a, skip); break;
        case 64: lc_pose_kernel<T, 64, MODE><<<a.B, 64, 0, st>>>(a, skip); break;
        case 128: lc_pose_kernel<T, 128, MODE><<<a.B, 128, 0, st>>>(a, skip); break;
        default: lc_pose_kernel<T, 256, MODE><<<a.B, 256, 0, st>>>(a, skip); break;
    }
    note_kernel("lc::lc_pose_kernel<%s,%d,%s>", sizeof(T) == 4 ? "float" : "double", stream_threads_for(a.N),
                MODE == MODE_LM ? "LM" : (MODE == MODE_LC ? "LC" : "LM|LC"));
    return static_cast<int>(cudaGetLastError());
}

int launch_stream_pose(const lc_args& a, int mode, cudaStream_t st, int n_skip_le) {
    const bool f32 = a.dtype == LC_F32;
    const int k = n_skip_le;
    switch (mode) {
        case MODE_LM: return f32 ? launch_pose_t<float, MODE_LM>(a, st, k) : launch_pose_t<double, MODE_LM>(a, st, k);
        case MODE_LC: return f32 ? launch_pose_t<float, MODE_LC>(a, st, k) : launch_pose_t<double, MODE_LC>(a, st, k);
        default: return f32 ? launch_pose_t<float, MODE_LM | MODE_LC>(a, st, k) : launch_pose_t<double, MODE_LM | MODE_LC>(a, st, k);
    }
}

template <typename T, bool BWD>
static int launch_jac_t(const lc_args& a, cudaStream_t st) {
#define LC_JAC_LAUNCH(NT_)                                              \
    if (BWD) lc_jac_bwd_kernel<T, NT_><<<a.B, NT_, 0, st>>>(a);         \
    else lc_jac_kernel<T, NT_><<<a.B, NT_, 0, st>>>(a)
    switch (stream_threads_for(a.N)) {
        case 32: LC_JAC_LAUNCH(32); break;
        case 64: LC_JAC_LAUNCH(64); break;
        case 128: LC_JAC_LAUNCH(128); break;
        default: LC_JAC_LAUNCH(256); break;
    }
#undef LC_JAC_LAUNCH
    note_kernel("lc::%s<%s,%d>", BWD ? "lc_jac_bwd_kernel" : "lc_jac_kernel", sizeof(T) == 4 ? "float" : "double", stream_threads_for(a.N));
    return static_cast<int>(cudaGetLastError());
}

int launch_stream_jac(const lc_args& a, bool bwd, cudaStream_t st) {
    const bool f32 = a.dtype == LC_F32;
    if (bwd) return f32 ? launch_jac_t<float, true>(a, st) : launch_jac_t<double, true>(a, st);
    return f32 ? launch_jac_t<float, false>(a, st) : launch_jac_t<double, false>(a, st);
}

}  // namespace lc
