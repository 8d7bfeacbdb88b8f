// Device-side building blocks shared by the LC kernels: strided element access, the
// multi-value warp/CTA reduction, packed symmetric 6x6 helpers and small dense algebra.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/lc_b200.h"

namespace lc {

constexpr int kMaxWarps = 16;   // CTA sizes: 32 .. 512 threads
constexpr int kSym = 21;        // unique entries of a symmetric 6x6
constexpr unsigned kFull = 0xffffffffu;

// packed upper-triangular index of (i,j), i <= j, row-major over the upper triangle
__host__ __device__ constexpr int sym_idx(int i, int j) { return i * 6 - (i * (i - 1)) / 2 + (j - i); }

template <typename T>
__device__ __forceinline__ double ld(const lc_view& v, int64_t off) {
    return static_cast<double>(static_cast<const T*>(v.ptr)[off]);
}
template <typename T>
__device__ __forceinline__ void st(const lc_view& v, int64_t off, double x) {
    static_cast<T*>(v.ptr)[off] = static_cast<T>(x);
}

// torch.nan_to_num semantics for element type T (cer_solver.py:27-29)
template <typename T>
__device__ __forceinline__ double nan_to_num(double x) {
    const double mx = sizeof(T) == 4 ? 3.4028234663852886e38 : 1.7976931348623157e308;
    if (isnan(x)) return 0.0;
    if (isinf(x)) return x > 0 ? mx : -mx;
    return x;
}

// ---------------------------------------------------------------------------------------------
// Multi-value reduction.  Every thread holds V partial sums; the CTA needs the V totals.
// A plain butterfly costs 5*V 64-bit shuffles per warp.  Here each butterfly stage halves the
// number of live values per lane (lane keeps one half, ships the other half to its partner), so
// a warp spends V-1 (+padding) shuffles in total and ends with ceil(V/32) totals per lane.
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int half_up(int c) { return (c + 1) / 2; }

template <int CUR, int OFF>
__device__ __forceinline__ void rs_stage(double* v, unsigned lane) {
    constexpr int H = half_up(CUR);
    const bool up = (lane & OFF) != 0;
#pragma unroll
    for (int j = 0; j < H; ++j) {
        const double lo = v[j];
        const double hi = (j + H < CUR) ? v[j + H] : 0.0;
        const double send = up ? lo : hi;
        const double keep = up ? hi : lo;
        v[j] = keep + __shfl_xor_sync(kFull, send, OFF);
    }
}

// fp32 flavour for the fp32 point passes (one SHFL per exchanged value instead of two, no conversions): the per-thread
// partial sums are fp32 already; the pairwise tree adds five roundings at 1/sqrt(nodes) weight, the totals of the warps are
// then summed in fp64.
template <int CUR, int OFF>
__device__ __forceinline__ void rs_stage_f(float* v, unsigned lane) {
    constexpr int H = half_up(CUR);
    const bool up = (lane & OFF) != 0;
#pragma unroll
    for (int j = 0; j < H; ++j) {
        const float lo = v[j];
        const float hi = (j + H < CUR) ? v[j + H] : 0.f;
        const float send = up ? lo : hi;
        const float keep = up ? hi : lo;
        v[j] = keep + __shfl_xor_sync(kFull, send, OFF);
    }
}

template <int V>
struct ReduceShape {
    static constexpr int c1 = half_up(V), c2 = half_up(c1), c3 = half_up(c2), c4 = half_up(c3), c5 = half_up(c4);
};

// After the call, v[0..c5) of each lane hold totals; slot k of lane `lane` is original index
// orig_index<V>(k, lane) (or -1 for a padding slot).
template <int V>
__device__ __forceinline__ void warp_reduce_scatter(double (&v)[V], unsigned lane) {
    using S = ReduceShape<V>;
    rs_stage<V, 16>(v, lane);
    rs_stage<S::c1, 8>(v, lane);
    rs_stage<S::c2, 4>(v, lane);
    rs_stage<S::c3, 2>(v, lane);
    rs_stage<S::c4, 1>(v, lane);
}

template <int V>
__device__ __forceinline__ void warp_reduce_scatter_f(float (&v)[V], unsigned lane) {
    using S = ReduceShape<V>;
    rs_stage_f<V, 16>(v, lane);
    rs_stage_f<S::c1, 8>(v, lane);
    rs_stage_f<S::c2, 4>(v, lane);
    rs_stage_f<S::c3, 2>(v, lane);
    rs_stage_f<S::c4, 1>(v, lane);
}

template <int V>
__device__ __forceinline__ int orig_index(int k, unsigned lane) {
    using S = ReduceShape<V>;
    int idx = k;
    if (idx >= S::c5) return -1;
    idx += (lane & 1) ? S::c5 : 0;
    if (idx >= S::c4) return -1;
    idx += (lane & 2) ? S::c4 : 0;
    if (idx >= S::c3) return -1;
    idx += (lane & 4) ? S::c3 : 0;
    if (idx >= S::c2) return -1;
    idx += (lane & 8) ? S::c2 : 0;
    if (idx >= S::c1) return -1;
    idx += (lane & 16) ? S::c1 : 0;
    if (idx >= V) return -1;
    return idx;
}

// CTA-wide sum of V values per thread into fin[0..V) (shared).  red is [NT/32][V] shared scratch.
// Contains the barriers that make fin visible to every thread on return.
template <int V, int NT>
__device__ __forceinline__ void block_reduce(double (&v)[V], double* red, double* fin) {
    constexpr int NW = NT / 32;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    warp_reduce_scatter<V>(v, lane);
    double* dst = (NW == 1) ? fin : red + warp * V;
#pragma unroll
    for (int k = 0; k < ReduceShape<V>::c5; ++k) {
        const int idx = orig_index<V>(k, lane);
        if (idx >= 0) dst[idx] = v[k];
    }
    __syncthreads();
    if (NW > 1) {
        for (int j = threadIdx.x; j < V; j += NT) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) s += red[w * V + j];
            fin[j] = s;
        }
        __syncthreads();
    }
}

// Same for fp32 per-thread partial sums: fp32 tree inside the warp, fp64 across the warps.
template <int V, int NT>
__device__ __forceinline__ void block_reduce_f(float (&v)[V], double* red, double* fin) {
    constexpr int NW = NT / 32;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    warp_reduce_scatter_f<V>(v, lane);
    double* dst = (NW == 1) ? fin : red + warp * V;
#pragma unroll
    for (int k = 0; k < ReduceShape<V>::c5; ++k) {
        const int idx = orig_index<V>(k, lane);
        if (idx >= 0) dst[idx] = static_cast<double>(v[k]);
    }
    __syncthreads();
    if (NW > 1) {
        for (int j = threadIdx.x; j < V; j += NT) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) s += red[w * V + j];
            fin[j] = s;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// small dense algebra on shared-memory 6x6 matrices (row-major, 36 doubles)
// ---------------------------------------------------------------------------------------------

// O = A * B, one output entry per thread (threads 0..35); caller provides the barrier afterwards
template <int NT>
__device__ __forceinline__ void mm6_par(const double* A, const double* B, double* O) {
    for (int e = threadIdx.x; e < 36; e += NT) {
        const int r = e / 6, c = e % 6;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) s = fma(A[r * 6 + k], B[k * 6 + c], s);
        O[e] = s;
    }
}

// 1/a to ~1 ulp without the IEEE division sequence: MUFU.RCP64H seed + two Newton steps (5 issue slots
// instead of ~14 DFMA-equivalents measured for a correctly rounded division on B200).
__device__ __forceinline__ double fast_rcp(double a) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a, y, 1.0);
    y = fma(y, e, y);
    e = fma(-a, y, 1.0);
    return fma(y, e, y);
}

// Lower Cholesky + inverse of an SPD 6x6 (single thread, everything in registers: packed triangles, constant
// indices).  Returns 0 on success, else the order of the first non-positive leading minor (LAPACK info).
// A, C: full row-major 6x6.  C = A^-1.
__device__ inline int chol6_inverse(const double* A, double* C) {
    double L[21], il[6];   // L[i][j], i >= j, stored at sym_idx(j, i)
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = A[j * 6 + j];
#pragma unroll
        for (int k = 0; k < j; ++k) d = fma(-L[sym_idx(k, j)], L[sym_idx(k, j)], d);
        if (!(d > 0.0) || isinf(d)) return j + 1;
        il[j] = rsqrt(d);
        L[sym_idx(j, j)] = d * il[j];
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double v = A[i * 6 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) v = fma(-L[sym_idx(k, i)], L[sym_idx(k, j)], v);
            L[sym_idx(j, i)] = v * il[j];
        }
    }
    double Li[21];         // (L^-1)[r][c], r >= c, stored at sym_idx(c, r)
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        Li[sym_idx(c, c)] = il[c];
#pragma unroll
        for (int r = c + 1; r < 6; ++r) {
            double v = 0.0;
#pragma unroll
            for (int k = c; k < r; ++k) v = fma(-L[sym_idx(k, r)], Li[sym_idx(c, k)], v);
            Li[sym_idx(c, r)] = v * il[r];
        }
    }
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = a; b < 6; ++b) {
            double v = 0.0;
#pragma unroll
            for (int k = b; k < 6; ++k) v = fma(Li[sym_idx(a, k)], Li[sym_idx(b, k)], v);
            C[a * 6 + b] = v;
            C[b * 6 + a] = v;
        }
    return 0;
}

// Solve (A + diag(dd)) y = g for SPD A (packed symmetric `Ap`), single thread. Returns false on breakdown.
__device__ inline bool chol6_solve_packed(const double* Ap, const double* dd, const double* g, double* y) {
    double L[21];  // packed lower == packed upper of the transpose; index via sym_idx(min,max)
    double il[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = Ap[sym_idx(j, j)] + dd[j];
#pragma unroll
        for (int k = 0; k < j; ++k) d = fma(-L[sym_idx(k, j)], L[sym_idx(k, j)], d);
        if (!(d > 0.0) || isinf(d)) return false;
        il[j] = rsqrt(d);
        L[sym_idx(j, j)] = d * il[j];
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double v = Ap[sym_idx(j, i)];
#pragma unroll
            for (int k = 0; k < j; ++k) v = fma(-L[sym_idx(k, i)], L[sym_idx(k, j)], v);
            L[sym_idx(j, i)] = v * il[j];  // L[i][j]
        }
    }
    double z[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double v = g[i];
#pragma unroll
        for (int k = 0; k < i; ++k) v = fma(-L[sym_idx(k, i)], z[k], v);
        z[i] = v * il[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double v = z[i];
#pragma unroll
        for (int k = i + 1; k < 6; ++k) v = fma(-L[sym_idx(i, k)], y[k], v);
        y[i] = v * il[i];
    }
    return true;
}

// rotation_conversions.py:39-68 quaternion_to_matrix, including its two_s = 2/|q| scaling
__device__ inline void quat_to_R_ref(const double* q, double* R, double* qnorm) {
    const double r = q[0], i = q[1], j = q[2], k = q[3];
    const double n = sqrt(r * r + i * i + j * j + k * k);
    const double two_s = 2.0 / n;
    R[0] = 1 - two_s * (j * j + k * k); R[1] = two_s * (i * j - k * r); R[2] = two_s * (i * k + j * r);
    R[3] = two_s * (i * j + k * r); R[4] = 1 - two_s * (i * i + k * k); R[5] = two_s * (j * k - i * r);
    R[6] = two_s * (i * k - j * r); R[7] = two_s * (j * k + i * r); R[8] = 1 - two_s * (i * i + j * j);
    *qnorm = n;
}

}  // namespace lc
