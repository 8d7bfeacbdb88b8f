// lc_b200 — pose-error metrics and symmetric pose-candidate selection on the device.
//
//   lc_pose_errors_kernel   compute_pose_errors (lib/utils/evaluate.py:333-339): ADD, ADI, rotation and translation error
//                           (lib/utils/error6d.py:87-140), one CTA per pose.  The reference runs these in numpy /
//                           scipy.cKDTree inside a multiprocessing.Pool(6) (evaluate.py:193-210).
//   lc_select_pose_kernel   symmetry.select_pose_2d / select_pose_3d (symmetry.py:8-56): mean error of every pose
//                           candidate (B,K,3,4) against the predicted points, argmin over K, one CTA per sample; the
//                           reference materialises (B,K,N,3) intermediates.
#include "lc_resident.cuh"

namespace lc {

constexpr int kEvalNT = 256;
constexpr int kEvalTile = 2048;   // model points per shared-memory tile of the nearest-neighbour search

__device__ __forceinline__ double block_sum_eval(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < kEvalNT / 32; ++w) s += red[w];
    __syncthreads();
    return s;
}

__global__ void __launch_bounds__(kEvalNT) lc_pose_errors_kernel(const lc_eval_args d) {
    __shared__ float tile[kEvalTile * 3];
    __shared__ double red[kEvalNT / 32];
    __shared__ double Re[9], Rg[9], te[3], tg[3], A[9], c[3];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid < 9) {
        Re[tid] = ld<double>(d.R_est, b * d.R_est.stride[0] + (tid / 3) * d.R_est.stride[1] + (tid % 3) * d.R_est.stride[2]);
        Rg[tid] = ld<double>(d.R_gt, b * d.R_gt.stride[0] + (tid / 3) * d.R_gt.stride[1] + (tid % 3) * d.R_gt.stride[2]);
    } else if (tid < 12) {
        te[tid - 9] = ld<double>(d.t_est, b * d.t_est.stride[0] + (tid - 9) * d.t_est.stride[1]);
        tg[tid - 9] = ld<double>(d.t_gt, b * d.t_gt.stride[0] + (tid - 9) * d.t_gt.stride[1]);
    }
    __syncthreads();
    // ADI search frame: |(Re X_i + te) - (Rg X_j + tg)| = |X_i - (A X_j + c)| with A = Re^T Rg, c = Re^T (tg - te) when Re is a
    // rotation; the search runs over the raw model points, the distance itself is re-evaluated in the camera frame in fp64.
    if (tid < 9) {
        const int r = tid / 3, cc = tid % 3;
        A[tid] = Re[r] * Rg[cc] + Re[3 + r] * Rg[3 + cc] + Re[6 + r] * Rg[6 + cc];
    } else if (tid < 12) {
        const int r = tid - 9;
        c[r] = Re[r] * (tg[0] - te[0]) + Re[3 + r] * (tg[1] - te[1]) + Re[6 + r] * (tg[2] - te[2]);
    }
    __syncthreads();
    const int64_t off = d.pts_offset ? d.pts_offset[b] : 0;
    const int M = d.pts_count ? d.pts_count[b] : d.M;
    const double* P = static_cast<const double*>(d.pts.ptr);
    const int64_t ps = d.pts.stride[0], pc = d.pts.stride[1];
    auto xform = [](const double* R, const double* t, double x, double y, double z, double* o) {
        o[0] = R[0] * x + R[1] * y + R[2] * z + t[0]; o[1] = R[3] * x + R[4] * y + R[5] * z + t[1]; o[2] = R[6] * x + R[7] * y + R[8] * z + t[2];
    };
    double add_acc = 0.0, adi_acc = 0.0;
    for (int j0 = 0; j0 < M; j0 += kEvalNT) {
        const int j = j0 + tid;
        const bool live = j < M;
        double X[3] = {0, 0, 0}, pg[3] = {0, 0, 0};
        float qx = 0.f, qy = 0.f, qz = 0.f;
        if (live) {
            X[0] = P[(off + j) * ps]; X[1] = P[(off + j) * ps + pc]; X[2] = P[(off + j) * ps + 2 * pc];
            double pe[3];
            xform(Re, te, X[0], X[1], X[2], pe);
            xform(Rg, tg, X[0], X[1], X[2], pg);
            const double dx = pe[0] - pg[0], dy = pe[1] - pg[1], dz = pe[2] - pg[2];
            add_acc += sqrt(dx * dx + dy * dy + dz * dz);                      // error6d.py:98-101
            qx = static_cast<float>(A[0] * X[0] + A[1] * X[1] + A[2] * X[2] + c[0]);
            qy = static_cast<float>(A[3] * X[0] + A[4] * X[1] + A[5] * X[2] + c[1]);
            qz = static_cast<float>(A[6] * X[0] + A[7] * X[1] + A[8] * X[2] + c[2]);
        }
        if (d.adi.ptr) {
            // nearest estimated-pose vertex of this ground-truth-pose vertex (error6d.py:115-123), brute force over tiles
            float best = INFINITY;
            int besti = 0;
            for (int i0 = 0; i0 < M; i0 += kEvalTile) {
                const int cnt = min(kEvalTile, M - i0);
                __syncthreads();
                for (int k = tid; k < cnt * 3; k += kEvalNT) {
                    const int i = k / 3, cc = k - i * 3;
                    tile[cc * kEvalTile + i] = static_cast<float>(P[(off + i0 + i) * ps + cc * pc]);
                }
                __syncthreads();
                if (live) {
                    for (int i = 0; i < cnt; ++i) {
                        const float dx = tile[i] - qx, dy = tile[kEvalTile + i] - qy, dz = tile[2 * kEvalTile + i] - qz;
                        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        if (d2 < best) { best = d2; besti = i0 + i; }
                    }
                }
            }
            if (live) {
                double pe[3];
                xform(Re, te, P[(off + besti) * ps], P[(off + besti) * ps + pc], P[(off + besti) * ps + 2 * pc], pe);
                const double dx = pe[0] - pg[0], dy = pe[1] - pg[1], dz = pe[2] - pg[2];
                adi_acc += sqrt(dx * dx + dy * dy + dz * dz);
            }
        }
    }
    const double add_sum = block_sum_eval(add_acc, red), adi_sum = block_sum_eval(adi_acc, red);
    if (tid == 0) {
        const double inv = M > 0 ? 1.0 / M : 0.0;
        if (d.add.ptr) st<double>(d.add, b * d.add.stride[0], add_sum * inv);
        if (d.adi.ptr) st<double>(d.adi, b * d.adi.stride[0], adi_sum * inv);
        if (d.re.ptr) {
            // error6d.py:133-140: acos of 0.5 (trace(R_est R_gt^-1) - 1), degrees
            const double c00 = Rg[4] * Rg[8] - Rg[5] * Rg[7], c01 = Rg[5] * Rg[6] - Rg[3] * Rg[8], c02 = Rg[3] * Rg[7] - Rg[4] * Rg[6];
            const double idet = 1.0 / (Rg[0] * c00 + Rg[1] * c01 + Rg[2] * c02);
            const double Gi[9] = {c00 * idet, (Rg[2] * Rg[7] - Rg[1] * Rg[8]) * idet, (Rg[1] * Rg[5] - Rg[2] * Rg[4]) * idet,
                                  c01 * idet, (Rg[0] * Rg[8] - Rg[2] * Rg[6]) * idet, (Rg[2] * Rg[3] - Rg[0] * Rg[5]) * idet,
                                  c02 * idet, (Rg[1] * Rg[6] - Rg[0] * Rg[7]) * idet, (Rg[0] * Rg[4] - Rg[1] * Rg[3]) * idet};
            double tr = 0.0;
            for (int i = 0; i < 3; ++i)
                for (int k = 0; k < 3; ++k) tr += Re[i * 3 + k] * Gi[k * 3 + i];
            const double cs = fmin(1.0, fmax(-1.0, 0.5 * (tr - 1.0)));
            st<double>(d.re, b * d.re.stride[0], 180.0 * acos(cs) / 3.14159265358979323846);
        }
        if (d.te.ptr) {
            const double dx = tg[0] - te[0], dy = tg[1] - te[1], dz = tg[2] - te[2];
            st<double>(d.te, b * d.te.stride[0], sqrt(dx * dx + dy * dy + dz * dz));
        }
    }
}

// MODE 0: select_pose_2d (symmetry.py:8-31): err_k = mean_n | proj(K (R_k X_n + t_k)) - x_n |
// MODE 1: select_pose_3d (symmetry.py:33-56): err_k = mean_n | Xout_n - R_k^T (K^-1 h_n - t_k) |
template <int MODE>
__global__ void __launch_bounds__(kEvalNT) lc_select_pose_kernel(const lc_candi_args d) {
    __shared__ double red[kEvalNT / 32];
    __shared__ float Kc[9], Ki[9];
    __shared__ float best_err;
    __shared__ int best_k;
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid < 9) Kc[tid] = ldf(d.K, b * d.K.stride[0] + (tid / 3) * d.K.stride[1] + (tid % 3) * d.K.stride[2]);
    __syncthreads();
    if (tid == 0) {
        const float* R = Kc;
        const float c00 = R[4] * R[8] - R[5] * R[7], c01 = R[5] * R[6] - R[3] * R[8], c02 = R[3] * R[7] - R[4] * R[6];
        const float idet = 1.f / (R[0] * c00 + R[1] * c01 + R[2] * c02);
        Ki[0] = c00 * idet; Ki[1] = (R[2] * R[7] - R[1] * R[8]) * idet; Ki[2] = (R[1] * R[5] - R[2] * R[4]) * idet;
        Ki[3] = c01 * idet; Ki[4] = (R[0] * R[8] - R[2] * R[6]) * idet; Ki[5] = (R[2] * R[3] - R[0] * R[5]) * idet;
        Ki[6] = c02 * idet; Ki[7] = (R[1] * R[6] - R[0] * R[7]) * idet; Ki[8] = (R[0] * R[4] - R[1] * R[3]) * idet;
        best_err = INFINITY; best_k = 0;
    }
    __syncthreads();
    const float* A3 = static_cast<const float*>(d.pts_a.ptr) + b * d.pts_a.stride[0];   // pts3d (2d) / pts3d_out (3d)
    const float* B3 = static_cast<const float*>(d.pts_b.ptr) + b * d.pts_b.stride[0];   // pts2d (2d) / homo_z (3d)
    const int64_t an = d.pts_a.stride[1], ac = d.pts_a.stride[2], bn = d.pts_b.stride[1], bc_ = d.pts_b.stride[2];
    for (int k = 0; k < d.Kc; ++k) {
        float T[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) T[e] = ldf(d.candi, b * d.candi.stride[0] + k * d.candi.stride[1] + (e / 4) * d.candi.stride[2] + (e % 4) * d.candi.stride[3]);
        float acc = 0.f;
        for (int n = tid; n < d.N; n += kEvalNT) {
            const float a0 = A3[n * an], a1 = A3[n * an + ac], a2 = A3[n * an + 2 * ac];
            if (MODE == 0) {
                const float p0 = T[0] * a0 + T[1] * a1 + T[2] * a2 + T[3], p1 = T[4] * a0 + T[5] * a1 + T[6] * a2 + T[7],
                            p2 = T[8] * a0 + T[9] * a1 + T[10] * a2 + T[11];
                const float h0 = Kc[0] * p0 + Kc[1] * p1 + Kc[2] * p2, h1 = Kc[3] * p0 + Kc[4] * p1 + Kc[5] * p2, h2 = Kc[6] * p0 + Kc[7] * p1 + Kc[8] * p2;
                const float du = h0 / h2 - B3[n * bn], dv = h1 / h2 - B3[n * bn + bc_];
                acc += sqrtf(du * du + dv * dv);
            } else {
                const float h0 = B3[n * bn], h1 = B3[n * bn + bc_], h2 = B3[n * bn + 2 * bc_];
                const float c0 = Ki[0] * h0 + Ki[1] * h1 + Ki[2] * h2 - T[3], c1 = Ki[3] * h0 + Ki[4] * h1 + Ki[5] * h2 - T[7],
                            c2 = Ki[6] * h0 + Ki[7] * h1 + Ki[8] * h2 - T[11];
                const float r0 = T[0] * c0 + T[4] * c1 + T[8] * c2, r1 = T[1] * c0 + T[5] * c1 + T[9] * c2, r2 = T[2] * c0 + T[6] * c1 + T[10] * c2;
                const float dx = a0 - r0, dy = a1 - r1, dz = a2 - r2;
                acc += sqrtf(dx * dx + dy * dy + dz * dz);
            }
        }
        const double s = block_sum_eval(acc, red);
        if (tid == 0) {
            const float err = static_cast<float>(s / d.N);
            if (d.err.ptr) stf(d.err, b * d.err.stride[0] + k * d.err.stride[1], err);
            if (err < best_err) { best_err = err; best_k = k; }   // torch.argmin: first minimum
        }
    }
    __syncthreads();
    if (tid < 12 && d.best.ptr)
        stf(d.best, b * d.best.stride[0] + (tid / 4) * d.best.stride[1] + (tid % 4) * d.best.stride[2],
            ldf(d.candi, b * d.candi.stride[0] + best_k * d.candi.stride[1] + (tid / 4) * d.candi.stride[2] + (tid % 4) * d.candi.stride[3]));
    if (tid == 0 && d.best_index) d.best_index[b] = best_k;
}

int launch_pose_errors(const lc_eval_args& d, cudaStream_t st) {
    lc_pose_errors_kernel<<<d.B, kEvalNT, 0, st>>>(d);
    note_kernel("lc::lc_pose_errors_kernel");
    return static_cast<int>(cudaGetLastError());
}
int launch_select_pose(const lc_candi_args& d, cudaStream_t st) {
    if (d.mode == 0) lc_select_pose_kernel<0><<<d.B, kEvalNT, 0, st>>>(d);
    else lc_select_pose_kernel<1><<<d.B, kEvalNT, 0, st>>>(d);
    note_kernel("lc::lc_select_pose_kernel<%d>", d.mode);
    return static_cast<int>(cudaGetLastError());
}

}  // namespace lc
