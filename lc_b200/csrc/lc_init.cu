// lc_b200 — device-side pose initialiser (include/lc_b200.h: lc_init_args).
//
// Role of lib/pnp/cv2_solver.solve (cv2_solver.py:6-88: cv2.solvePnPRansac(EPNP, 150 iterations) per sample on the host, after
// a device->host copy and a synchronize): produce the start pose of the weighted LM solve (test.py:120-127) and the inlier set
// of the 'weighted_filtered' branch (test.py:131-134).  NOT a port of OpenCV's RANSAC (its random stream cannot be reproduced
// and is not needed: the start only has to land in the basin of the LM solve that follows).  One CTA per pose:
//   * weighted DLT in normalised coordinates (K^-1 x, Hartley-normalised X), reduced analytically to a 4x4 symmetric
//     eigenproblem: with S = sum w X~X~^T, Sx = sum w x X~X~^T, Sy, Sq = sum w (x^2+y^2) X~X~^T the rows of P = [p1;p2;p3]
//     minimise sum w [(p1.X~ - x p3.X~)^2 + (p2.X~ - y p3.X~)^2]  =>  p1 = S^-1 Sx p3, p2 = S^-1 Sy p3,
//     p3 = smallest eigenvector of Sq - Sx S^-1 Sx - Sy S^-1 Sy (inverse iteration).  40 fp64 sums per pass;
//   * R from the first two rows of P (robust in the weak-perspective regime of small objects), r3 = r1 x r2;
//   * IRLS: Cauchy weights on the pixel reprojection error with scale `reproj_thresh`, `irls_rounds` re-solves;
//   * inlier mask = reprojection error < reproj_thresh under the returned pose (the role of RANSAC's inlier set).
#include "lc_resident.cuh"

namespace lc {

constexpr int kInitNT = 256;
constexpr int kInitMaxPts = 1024;   // correspondences used for the DLT / IRLS sums (sub-sampled by stride when there are more)

// Lower Cholesky factor of the SPD 4x4 S with reciprocal diagonal (no IEEE division / sqrt: rsqrt + multiplies).
__device__ bool chol4(const double S[4][4], double L[4][4], double il[4]) {
    for (int j = 0; j < 4; ++j) {
        double d = S[j][j];
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
        if (!(d > 0.0) || isinf(d)) return false;
        il[j] = rsqrt(d);
        L[j][j] = d * il[j];
        for (int i = j + 1; i < 4; ++i) {
            double v = S[i][j];
            for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
            L[i][j] = v * il[j];
        }
    }
    return true;
}
__device__ void chol4_backsolve(const double L[4][4], const double il[4], const double* rhs, double* z) {
    double y[4];
    for (int i = 0; i < 4; ++i) { double v = rhs[i]; for (int k = 0; k < i; ++k) v -= L[i][k] * y[k]; y[i] = v * il[i]; }
    for (int i = 3; i >= 0; --i) { double v = y[i]; for (int k = i + 1; k < 4; ++k) v -= L[k][i] * z[k]; z[i] = v * il[i]; }
}

// Z = S^-1 Bm for 4 right-hand sides.  false if S is not SPD.
__device__ bool chol4_solve(const double S[4][4], const double Bm[4][4], double Z[4][4]) {
    double L[4][4] = {}, il[4];
    if (!chol4(S, L, il)) return false;
    for (int c = 0; c < 4; ++c) {
        double rhs[4], z[4];
        for (int i = 0; i < 4; ++i) rhs[i] = Bm[i][c];
        chol4_backsolve(L, il, rhs, z);
        for (int i = 0; i < 4; ++i) Z[i][c] = z[i];
    }
    return true;
}

// Eigenvector of the smallest eigenvalue of a symmetric positive semi-definite 4x4 by inverse iteration on D + mu I
// (mu = 1e-12 trace keeps the factorisation defined for noise-free data, where the smallest eigenvalue is 0 up to rounding).
// The eigenvalue gap of this problem is the ratio of the noise to the signal energy (1e-2 .. 1e-6), so kInvIters steps reach
// fp64 precision; a fully serial cyclic Jacobi sweep costs ~10x more dependent fp64 operations.
constexpr int kInvIters = 5;
__device__ bool sym4_min_eigvec(const double D[4][4], double* v) {
    const double tr = D[0][0] + D[1][1] + D[2][2] + D[3][3];
    if (!(tr > 0.0) || isinf(tr)) return false;
    double A[4][4], L[4][4] = {}, il[4];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) A[r][c] = D[r][c] + (r == c ? 1e-12 * tr : 0.0);
    if (!chol4(A, L, il)) return false;
    v[0] = 0.1; v[1] = 0.1; v[2] = 0.1; v[3] = 1.0;
    for (int it = 0; it < kInvIters; ++it) {
        double w[4];
        chol4_backsolve(L, il, v, w);
        const double nrm = rsqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2] + w[3] * w[3]);
        for (int k = 0; k < 4; ++k) v[k] = w[k] * nrm;
    }
    return isfinite(v[0]) && isfinite(v[1]) && isfinite(v[2]) && isfinite(v[3]);
}

struct InitShared {
    double red[kMaxWarps * 48], fin[48];
    double Ki[9], K[9], R[9], t[3], cen[3], scale;
    int ok;
};

__global__ void __launch_bounds__(kInitNT) lc_init_kernel(const lc_init_args d) {
    __shared__ InitShared s;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = d.n_points ? min(max(d.n_points[b], 0), d.N) : d.N;
    const float* X = static_cast<const float*>(d.pts3d.ptr) + b * d.pts3d.stride[0];
    const float* x = static_cast<const float*>(d.pts2d.ptr) + b * d.pts2d.stride[0];
    const float* w = d.weights.ptr ? static_cast<const float*>(d.weights.ptr) + b * d.weights.stride[0] : nullptr;
    const int64_t Xn = d.pts3d.stride[1], Xc = d.pts3d.stride[2], xn = d.pts2d.stride[1], xc = d.pts2d.stride[2];
    const int64_t wn = w ? d.weights.stride[1] : 0, wc = w ? d.weights.stride[2] : 0;
    const double thr = d.reproj_thresh_b.ptr ? static_cast<double>(ldf(d.reproj_thresh_b, b * d.reproj_thresh_b.stride[0])) : static_cast<double>(d.reproj_thresh);
    if (tid < 9) s.K[tid] = ldf(d.K, b * d.K.stride[0] + (tid / 3) * d.K.stride[1] + (tid % 3) * d.K.stride[2]);
    __syncthreads();
    if (tid == 0) {
        const double* R = s.K;
        const double c00 = R[4] * R[8] - R[5] * R[7], c01 = R[5] * R[6] - R[3] * R[8], c02 = R[3] * R[7] - R[4] * R[6];
        const double idet = 1.0 / (R[0] * c00 + R[1] * c01 + R[2] * c02);
        s.Ki[0] = c00 * idet; s.Ki[1] = (R[2] * R[7] - R[1] * R[8]) * idet; s.Ki[2] = (R[1] * R[5] - R[2] * R[4]) * idet;
        s.Ki[3] = c01 * idet; s.Ki[4] = (R[0] * R[8] - R[2] * R[6]) * idet; s.Ki[5] = (R[2] * R[3] - R[0] * R[5]) * idet;
        s.Ki[6] = c02 * idet; s.Ki[7] = (R[1] * R[6] - R[0] * R[7]) * idet; s.Ki[8] = (R[0] * R[4] - R[1] * R[3]) * idet;
        s.ok = n >= 6 ? 1 : 0;
    }
    // A start pose does not need every correspondence: the sums run over every `step`-th point (about kInitMaxPts of them);
    // the inlier mask at the end covers all points.
    const int step = n > kInitMaxPts ? (n + kInitMaxPts - 1) / kInitMaxPts : 1;
    const int n_used = (n + step - 1) / step;
    // ---- Hartley normalisation of the model points ----
    {
        double acc[6] = {0, 0, 0, 0, 0, 0};
        for (int i = tid * step; i < n; i += kInitNT * step) {
            const double a = X[i * Xn], bb = X[i * Xn + Xc], c = X[i * Xn + 2 * Xc];
            acc[0] += a; acc[1] += bb; acc[2] += c; acc[3] += a * a; acc[4] += bb * bb; acc[5] += c * c;
        }
        block_reduce<6, kInitNT>(acc, s.red, s.fin);
        if (tid == 0) {
            const double inv = n_used > 0 ? 1.0 / n_used : 0.0;
            double var = 0.0;
            for (int k = 0; k < 3; ++k) { s.cen[k] = s.fin[k] * inv; var += s.fin[3 + k] * inv - s.cen[k] * s.cen[k]; }
            s.scale = var > 0.0 ? sqrt(var / 3.0) : 1.0;
        }
        __syncthreads();
    }
    const double c0 = s.cen[0], c1 = s.cen[1], c2 = s.cen[2], isc = 1.0 / s.scale, ithr2 = 1.0 / (thr * thr);

    for (int round = 0; round <= d.irls_rounds; ++round) {
        // ---- 40 weighted sums ----
        double acc[40];
#pragma unroll
        for (int k = 0; k < 40; ++k) acc[k] = 0.0;
        const bool robust = round > 0;
        double R[9], t[3];
        if (robust) { for (int k = 0; k < 9; ++k) R[k] = s.R[k]; for (int k = 0; k < 3; ++k) t[k] = s.t[k]; }
        for (int i = tid * step; i < n; i += kInitNT * step) {
            const double X0 = X[i * Xn], X1 = X[i * Xn + Xc], X2 = X[i * Xn + 2 * Xc];
            const double u = x[i * xn], v = x[i * xn + xc];
            double wt = w ? 0.5 * (static_cast<double>(w[i * wn]) + static_cast<double>(w[i * wn + wc])) : 1.0;
            if (!(wt > 0.0) || isinf(wt)) wt = 0.0;
            if (robust) {
                const double p0 = R[0] * X0 + R[1] * X1 + R[2] * X2 + t[0], p1 = R[3] * X0 + R[4] * X1 + R[5] * X2 + t[1],
                             p2 = R[6] * X0 + R[7] * X1 + R[8] * X2 + t[2];
                const double h2 = s.K[6] * p0 + s.K[7] * p1 + s.K[8] * p2, ih2 = fast_rcp(h2);
                const double eu = (s.K[0] * p0 + s.K[1] * p1 + s.K[2] * p2) * ih2 - u, ev = (s.K[3] * p0 + s.K[4] * p1 + s.K[5] * p2) * ih2 - v;
                const double e2 = (eu * eu + ev * ev) * ithr2;
                wt *= (h2 > 0.0 && isfinite(e2)) ? fast_rcp(1.0 + e2) : 0.0;     // Cauchy
            }
            const double izn = fast_rcp(s.Ki[6] * u + s.Ki[7] * v + s.Ki[8]);
            const double xh = (s.Ki[0] * u + s.Ki[1] * v + s.Ki[2]) * izn, yh = (s.Ki[3] * u + s.Ki[4] * v + s.Ki[5]) * izn;
            const double Y[4] = {(X0 - c0) * isc, (X1 - c1) * isc, (X2 - c2) * isc, 1.0};
            const double wq = wt * (xh * xh + yh * yh), wx = wt * xh, wy = wt * yh;
            int k = 0;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = r; c < 4; ++c) {
                    const double pp = Y[r] * Y[c];
                    acc[k] = fma(wt, pp, acc[k]); acc[10 + k] = fma(wx, pp, acc[10 + k]);
                    acc[20 + k] = fma(wy, pp, acc[20 + k]); acc[30 + k] = fma(wq, pp, acc[30 + k]);
                    ++k;
                }
        }
        block_reduce<40, kInitNT>(acc, s.red, s.fin);
        if (tid == 0 && s.ok) {
            double S[4][4], Sx[4][4], Sy[4][4], Sq[4][4];
            int k = 0;
            for (int r = 0; r < 4; ++r)
                for (int c = r; c < 4; ++c) {
                    S[r][c] = S[c][r] = s.fin[k]; Sx[r][c] = Sx[c][r] = s.fin[10 + k];
                    Sy[r][c] = Sy[c][r] = s.fin[20 + k]; Sq[r][c] = Sq[c][r] = s.fin[30 + k];
                    ++k;
                }
            double Zx[4][4], Zy[4][4];
            if (!chol4_solve(S, Sx, Zx) || !chol4_solve(S, Sy, Zy)) {
                s.ok = 0;
            } else {
                double D[4][4];
                for (int r = 0; r < 4; ++r)
                    for (int c = 0; c < 4; ++c) {
                        double v = Sq[r][c];
                        for (int m = 0; m < 4; ++m) v -= Sx[r][m] * Zx[m][c] + Sy[r][m] * Zy[m][c];
                        D[r][c] = v;
                    }
                for (int r = 0; r < 4; ++r)
                    for (int c = r + 1; c < 4; ++c) D[r][c] = D[c][r] = 0.5 * (D[r][c] + D[c][r]);
                double p3[4] = {0, 0, 0, 1}, p1[4], p2[4];
                if (!sym4_min_eigvec(D, p3)) s.ok = 0;
                for (int r = 0; r < 4; ++r) {
                    p1[r] = Zx[r][0] * p3[0] + Zx[r][1] * p3[1] + Zx[r][2] * p3[2] + Zx[r][3] * p3[3];
                    p2[r] = Zy[r][0] * p3[0] + Zy[r][1] * p3[1] + Zy[r][2] * p3[2] + Zy[r][3] * p3[3];
                }
                // undo the normalisation X' = (X - cen) / scale:  Q = P' T
                double Q[3][4];
                const double* P[3] = {p1, p2, p3};
                for (int r = 0; r < 3; ++r) {
                    for (int c = 0; c < 3; ++c) Q[r][c] = P[r][c] * isc;
                    Q[r][3] = P[r][3] - (P[r][0] * c0 + P[r][1] * c1 + P[r][2] * c2) * isc;
                }
                const double depth_c = Q[2][0] * c0 + Q[2][1] * c1 + Q[2][2] * c2 + Q[2][3];
                if (depth_c < 0.0)
                    for (int r = 0; r < 3; ++r)
                        for (int c = 0; c < 4; ++c) Q[r][c] = -Q[r][c];
                const double n1 = sqrt(Q[0][0] * Q[0][0] + Q[0][1] * Q[0][1] + Q[0][2] * Q[0][2]);
                const double n2 = sqrt(Q[1][0] * Q[1][0] + Q[1][1] * Q[1][1] + Q[1][2] * Q[1][2]);
                const double lam = 0.5 * (n1 + n2);
                double r1[3] = {Q[0][0] / n1, Q[0][1] / n1, Q[0][2] / n1};
                const double dp = r1[0] * Q[1][0] + r1[1] * Q[1][1] + r1[2] * Q[1][2];
                double r2[3] = {Q[1][0] - dp * r1[0], Q[1][1] - dp * r1[1], Q[1][2] - dp * r1[2]};
                const double n2o = sqrt(r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2]);
                for (int k2 = 0; k2 < 3; ++k2) r2[k2] /= n2o;
                const double r3[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
                for (int k2 = 0; k2 < 3; ++k2) { s.R[k2] = r1[k2]; s.R[3 + k2] = r2[k2]; s.R[6 + k2] = r3[k2]; s.t[k2] = Q[k2][3] / lam; }
                bool fin = isfinite(lam) && lam > 0.0 && n2o > 0.0;
                for (int k2 = 0; k2 < 9; ++k2) fin = fin && isfinite(s.R[k2]);
                for (int k2 = 0; k2 < 3; ++k2) fin = fin && isfinite(s.t[k2]);
                if (!fin) s.ok = 0;
            }
        }
        __syncthreads();
        if (!s.ok) break;
    }
    const bool ok = s.ok != 0;
    // ---- outputs: state (wxyz + t, w >= 0 like axis_angle_to_quaternion of an rvec), inlier mask ----
    if (tid == 0) {
        double q[4] = {1, 0, 0, 0}, tt[3] = {0, 0, 0};
        if (ok) {
            const double* R = s.R;
            const double tr = R[0] + R[4] + R[8];
            if (tr > 0.0) { const double sq = sqrt(tr + 1.0) * 2.0; q[0] = 0.25 * sq; q[1] = (R[7] - R[5]) / sq; q[2] = (R[2] - R[6]) / sq; q[3] = (R[3] - R[1]) / sq; }
            else if (R[0] > R[4] && R[0] > R[8]) { const double sq = sqrt(1.0 + R[0] - R[4] - R[8]) * 2.0; q[0] = (R[7] - R[5]) / sq; q[1] = 0.25 * sq; q[2] = (R[1] + R[3]) / sq; q[3] = (R[2] + R[6]) / sq; }
            else if (R[4] > R[8]) { const double sq = sqrt(1.0 + R[4] - R[0] - R[8]) * 2.0; q[0] = (R[2] - R[6]) / sq; q[1] = (R[1] + R[3]) / sq; q[2] = 0.25 * sq; q[3] = (R[5] + R[7]) / sq; }
            else { const double sq = sqrt(1.0 + R[8] - R[0] - R[4]) * 2.0; q[0] = (R[3] - R[1]) / sq; q[1] = (R[2] + R[6]) / sq; q[2] = (R[5] + R[7]) / sq; q[3] = 0.25 * sq; }
            if (q[0] < 0.0) for (int k = 0; k < 4; ++k) q[k] = -q[k];
            for (int k = 0; k < 3; ++k) tt[k] = s.t[k];
        }
        for (int k = 0; k < 4; ++k) stf(d.state, b * d.state.stride[0] + k * d.state.stride[1], static_cast<float>(q[k]));
        for (int k = 0; k < 3; ++k) stf(d.state, b * d.state.stride[0] + (4 + k) * d.state.stride[1], static_cast<float>(tt[k]));
        if (d.invalid) d.invalid[b] = ok ? 0 : 1;
    }
    if (d.inlier || d.n_inliers) {
        int cnt = 0;
        for (int i = tid; i < d.N; i += kInitNT) {
            bool in = false;
            if (ok && i < n) {
                const double X0 = X[i * Xn], X1 = X[i * Xn + Xc], X2 = X[i * Xn + 2 * Xc];
                const double p0 = s.R[0] * X0 + s.R[1] * X1 + s.R[2] * X2 + s.t[0], p1 = s.R[3] * X0 + s.R[4] * X1 + s.R[5] * X2 + s.t[1],
                             p2 = s.R[6] * X0 + s.R[7] * X1 + s.R[8] * X2 + s.t[2];
                const double h2 = s.K[6] * p0 + s.K[7] * p1 + s.K[8] * p2;
                const double eu = (s.K[0] * p0 + s.K[1] * p1 + s.K[2] * p2) / h2 - x[i * xn], ev = (s.K[3] * p0 + s.K[4] * p1 + s.K[5] * p2) / h2 - x[i * xn + xc];
                in = h2 > 0.0 && (eu * eu + ev * ev) < thr * thr;
            }
            if (d.inlier) d.inlier[static_cast<int64_t>(b) * d.N + i] = in ? 1 : 0;
            cnt += in ? 1 : 0;
        }
        double c1d[1] = {static_cast<double>(cnt)};
        block_reduce<1, kInitNT>(c1d, s.red, s.fin);
        if (tid == 0 && d.n_inliers) d.n_inliers[b] = static_cast<int>(s.fin[0]);
    }
}

int launch_init(const lc_init_args& d, cudaStream_t st) {
    lc_init_kernel<<<d.B, kInitNT, 0, st>>>(d);
    note_kernel("lc::lc_init_kernel");
    return static_cast<int>(cudaGetLastError());
}

}  // namespace lc
