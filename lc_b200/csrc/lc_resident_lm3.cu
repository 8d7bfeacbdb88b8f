// lc_b200 — solve-only resident kernel with THREE poses per SM (2048 < N <= ~5.2k).
//
// Why: with both resident CTAs of an SM inside their Jacobian passes the all-fp64 pass runs at ~50 of the 56 DFMA lanes/clk/SM,
// but over a whole launch the fp64 pipe is busy only 44 % of the time (profiles/README.md): whenever one of the two
// CTAs is in a serial section (trust-region step, reduction, staging) the other one alone cannot fill the pipe — its own
// per-point dependency chain bounds it.  A third pose per SM closes most of those gaps.  Shared memory is what limited the SM to
// two poses (20 B/point); the solve reads the image points x only once per pass, in order, exactly like the weights, so here x
// (8 B/point) is streamed from L2 next to the weights (requested one loop trip ahead, both prefetched into L2 by one bulk
// request at CTA start) and only the model points X (12 B/point, re-read through the rotation every pass) are staged:
// 48 KB + 15 KB per pose at N = 4096 -> three CTAs of 128 threads x 168 registers per SM.
//
// Same Ceres-faithful trust-region loop (lm_advance), same per-point arithmetic and accumulation order per thread as
// lm_eval_pass_res; fp32 tensors, diagonal weights, any strides (planar 16-byte aligned pts3d goes through the TMA).
#include <atomic>

#include "lc_resident.cuh"

namespace lc {

constexpr int kLm3NT = 128;

// what the solve needs per pose in shared memory (2.4 KB instead of the 15 KB PoseShared of the loss kernels)
struct LmShared {
    double K[9], pose[7];
    double red[(kLm3NT / 32) * 28];
    double fin[28];
    unsigned long long tma_bar;
    LmState lm;
};

__device__ __forceinline__ float* lm3_points(unsigned char* base) {
    return reinterpret_cast<float*>(base + ((sizeof(LmShared) + 15) & ~size_t(15)));
}
static size_t lm3_smem_bytes(int n) { return ((sizeof(LmShared) + 15) & ~size_t(15)) + sizeof(float) * 3 * static_cast<size_t>(round_up4(n)); }

// One evaluation pass (cost; with JAC also J'^T J' and J'^T r in the left basis): X from shared memory, x and the weights from L2.
template <int NT, bool JAC>
__device__ __forceinline__ void lm3_eval_pass(const lc_args& a, LmShared& s, const float* A0, const float* A1, const float* A2, int b, int n,
                                              bool sanitize) {
    const LmState& L = s.lm;
    double acc[28];
#pragma unroll
    for (int k = 0; k < 28; ++k) acc[k] = 0.0;
    const double k00 = s.K[0], k01 = s.K[1], k10 = s.K[3], k11 = s.K[4], cx = s.K[2], cy = s.K[5];
    const double R0 = L.Rm[0], R1 = L.Rm[1], R2 = L.Rm[2], R3 = L.Rm[3], R4 = L.Rm[4], R5 = L.Rm[5], R6 = L.Rm[6], R7 = L.Rm[7], R8 = L.Rm[8];
    const double t0 = L.te[0], t1 = L.te[1], t2 = L.te[2];
    const float* pw = static_cast<const float*>(a.weights.ptr) + b * a.weights.stride[0];
    const int64_t swn = a.weights.stride[1], swc = a.weights.stride[2];
    const float* p2 = static_cast<const float*>(a.pts2d.ptr) + b * a.pts2d.stride[0];
    const int64_t s2n = a.pts2d.stride[1], s2c = a.pts2d.stride[2];
    const bool icov = a.weight_mode == LC_W_ICOV_DIAG;
    float w0 = 0.f, w1 = 0.f, u0 = 0.f, u1 = 0.f;
    if (static_cast<int>(threadIdx.x) < n) {
        const int i = threadIdx.x;
        w0 = __ldg(pw + i * swn); w1 = __ldg(pw + i * swn + swc); u0 = __ldg(p2 + i * s2n); u1 = __ldg(p2 + i * s2n + s2c);
    }
    for (int i = threadIdx.x; i < n; i += NT) {
        float wa = w0, wb = w1, pxf = u0, pyf = u1;
        const int inext = i + NT;
        if (inext < n) {
            w0 = __ldg(pw + inext * swn); w1 = __ldg(pw + inext * swn + swc);
            u0 = __ldg(p2 + inext * s2n); u1 = __ldg(p2 + inext * s2n + s2c);
        }
        if (sanitize) { wa = nan_to_num_f(wa); wb = nan_to_num_f(wb); pxf = nan_to_num_f(pxf); pyf = nan_to_num_f(pyf); }
        if (icov) { wa = sqrtf(wa); wb = sqrtf(wb); }   // cer_solver.py:37-38; for LC_W_INV_STD sqrt(fl(s*s)) == |s|
        const double la = fabsf(wa), lc_ = fabsf(wb);
        const double X0 = A0[i], X1 = A1[i], X2 = A2[i];
        const double px = pxf, py = pyf;
        const double q0 = fma(R0, X0, fma(R1, X1, R2 * X2));
        const double q1 = fma(R3, X0, fma(R4, X1, R5 * X2));
        const double q2 = fma(R6, X0, fma(R7, X1, R8 * X2));
        const double p0 = q0 + t0, p1 = q1 + t1, pz = q2 + t2;
        const double iz = fast_rcp(pz);
        const double up = fma(p0, k00, p1 * k01) * iz, vp = fma(p0, k10, p1 * k11) * iz;
        const double du = up - (px - cx), dv = vp - (py - cy);
        const double r0 = du * la, r1 = dv * lc_;
        acc[27] = fma(r0, r0, fma(r1, r1, acc[27]));
        if (JAC) {
            const double a0 = la * iz, a1 = lc_ * iz;
            double J0[6], J1[6];
            J0[3] = a0 * k00; J0[4] = a0 * k01; J0[5] = -a0 * up;
            J1[3] = a1 * k10; J1[4] = a1 * k11; J1[5] = -a1 * vp;
            J0[0] = fma(q1, J0[5], -q2 * J0[4]); J0[1] = fma(q2, J0[3], -q0 * J0[5]); J0[2] = fma(q0, J0[4], -q1 * J0[3]);
            J1[0] = fma(q1, J1[5], -q2 * J1[4]); J1[1] = fma(q2, J1[3], -q0 * J1[5]); J1[2] = fma(q0, J1[4], -q1 * J1[3]);
            int k = 0;
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = r; c < 6; ++c) {
                    acc[k] = fma(J0[r], J0[c], fma(J1[r], J1[c], acc[k]));
                    ++k;
                }
#pragma unroll
            for (int c = 0; c < 6; ++c) acc[21 + c] = fma(J0[c], r0, fma(J1[c], r1, acc[21 + c]));
        }
    }
    acc[27] *= 0.5;
    block_reduce<28, NT>(acc, s.red, s.fin);
}

template <int NT, int CTAS>
__global__ void __launch_bounds__(NT, CTAS) lc_lm3_kernel(const lc_args a, int npad, int tma_mask) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LmShared& s = *reinterpret_cast<LmShared*>(smem_raw);
    float* A0 = lm3_points(smem_raw);
    float* A1 = A0 + npad;
    float* A2 = A1 + npad;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = a.n_points ? min(max(a.n_points[b], 0), a.N) : a.N;
    const bool sanitize = (a.flags & LC_FLAG_NAN_TO_NUM) != 0;
    // ---- stage the model points; ask the L2 for the image points and the weights ----
    {
        const float* p3 = static_cast<const float*>(a.pts3d.ptr) + b * a.pts3d.stride[0];
        const int64_t s3n = a.pts3d.stride[1], s3c = a.pts3d.stride[2];
        const bool tma3 = (tma_mask & 1) != 0;
        if (tma_mask) {
            if (tid == 0) mbar_init(&s.tma_bar, 1);
            __syncthreads();
            if (tid == 0) {
                const unsigned slab = static_cast<unsigned>(min(a.N, npad)) * 4u;
                if (tma3) {
                    mbar_expect_tx(&s.tma_bar, slab * 3u);
                    tma_load_1d(A0, p3, slab, &s.tma_bar); tma_load_1d(A1, p3 + s3c, slab, &s.tma_bar); tma_load_1d(A2, p3 + 2 * s3c, slab, &s.tma_bar);
                }
                if (tma_mask & 2) {
                    const float* p2 = static_cast<const float*>(a.pts2d.ptr) + b * a.pts2d.stride[0];
                    l2_prefetch_bulk(p2, slab); l2_prefetch_bulk(p2 + a.pts2d.stride[2], slab);
                }
                if (tma_mask & 4) {
                    const float* pw = static_cast<const float*>(a.weights.ptr) + b * a.weights.stride[0];
                    l2_prefetch_bulk(pw, slab); l2_prefetch_bulk(pw + a.weights.stride[2], slab);
                }
            }
        }
        if (!tma3) {
            for (int i = tid; i < n; i += NT) { cp_async4(A0 + i, p3 + i * s3n); cp_async4(A1 + i, p3 + i * s3n + s3c); cp_async4(A2 + i, p3 + i * s3n + 2 * s3c); }
        }
    }
    if (tid < 9) {
        float v = ldf(a.K, b * a.K.stride[0] + (tid / 3) * a.K.stride[1] + (tid % 3) * a.K.stride[2]);
        s.K[tid] = sanitize ? nan_to_num_f(v) : v;
    } else if (tid < 16) {
        float v = ldf(a.pose, b * a.pose.stride[0] + (tid - 9) * a.pose.stride[1]);
        s.pose[tid - 9] = sanitize ? nan_to_num_f(v) : v;
    }
    cp_async_commit_wait_all();
    if (tma_mask & 1) mbar_wait(&s.tma_bar, 0);
    if (sanitize)
        for (int i = tid; i < n; i += NT) { A0[i] = nan_to_num_f(A0[i]); A1[i] = nan_to_num_f(A1[i]); A2[i] = nan_to_num_f(A2[i]); }
    __syncthreads();

    // ---- LM solve (fp64), exactly the loop of lc_resident_kernel ----
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
            if (kind == CTL_EVAL_COST) lm3_eval_pass<NT, false>(a, s, A0, A1, A2, b, n, sanitize);
            else lm3_eval_pass<NT, true>(a, s, A0, A1, A2, b, n, sanitize);
            if (tid == 0)
                lm_advance(L, s.fin, kind, first, a.max_iter, a.function_tolerance, (a.flags & LC_FLAG_TOL_NEEDS_SUCCESS) != 0, trace);
            first = false;
            __syncthreads();
            if (L.ctl == CTL_STOP) break;
        }
        solved = L.term == TERM_CONVERGENCE;
    }
    if (tid == 0) lm_write_result<float, LmShared>(a, s, b, n, solved);
}

// A (B,N,C) fp32 view whose component slabs are contiguous and 16-byte aligned for every pose
static bool lm3_planar(const lc_view& v, int n) {
    return v.ptr && v.stride[1] == 1 && (n % 4) == 0 && (reinterpret_cast<uintptr_t>(v.ptr) % 16) == 0 && (v.stride[0] % 4) == 0 && (v.stride[2] % 4) == 0;
}

static std::atomic<int> g_lm3_smem[64];   // opt-in shared memory per block of the device, cached per device index

#ifndef LC_LM3_MIN_BATCH
#define LC_LM3_MIN_BATCH 2048   // see the measurements at the top of lm3_supported's callers (profiles/phase_timing_r2.txt)
#endif

// Poses per SM: three (158 registers, no spills).  Four fit in shared memory up to N ~ 4.3k with the 2.4 KB LmShared, but then the
// threads get 128 registers, the fp64 accumulators spill, and with ~200 KB of the SM's 256 KB configured as shared memory the L1
// that would absorb the spills is ~25 KB: measured 259 vs 203 us at B = 1024 and 1621 vs 1236 us at B = 8192.  LC_B200_LM_CTAS = 4
// selects that variant for A/B runs.
static int lm3_ctas(const lc_args& a, int smem_optin) {
    if (const char* e = getenv("LC_B200_LM_CTAS")) {
        if (atoi(e) == 4 && 4 * (lm3_smem_bytes(a.N) + 1024) <= static_cast<size_t>(smem_optin) + 1024) return 4;
    }
    return 3;
}
static int lm3_optin_smem() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    const bool cached = dev >= 0 && dev < 64;
    if (cached && (v = g_lm3_smem[dev].load(std::memory_order_relaxed)) > 0) return v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cached) g_lm3_smem[dev].store(v, std::memory_order_relaxed);
    return v;
}

// LC_B200_LM3 = 1 forces this kernel at any batch size (tests, A/B runs), = 0 disables it; LC_B200_LM_CTAS = 3 | 4 pins the poses per SM.
bool lm3_supported(const lc_args& a) {
    const char* e = getenv("LC_B200_LM3");
    if (e && e[0] == '0') return false;
    if (a.dtype != LC_F32 || a.N <= 2048) return false;   // N <= 2048: four 128-thread CTAs of the 20 B/point kernel fit already
    if (a.weight_mode != LC_W_ICOV_DIAG && a.weight_mode != LC_W_INV_STD) return false;
    const int v = lm3_optin_smem();
    if (v <= 0) return false;
    // three CTAs (+1 KB of system shared memory each) must fit in the SM's shared memory = the per-block opt-in limit + 1 KB
    if (3 * (lm3_smem_bytes(a.N) + 1024) > static_cast<size_t>(v) + 1024) return false;
    if (e && e[0] == '1') return true;
    return a.B >= LC_LM3_MIN_BATCH;
}

template <int CTAS>
static int launch_lm3_t(const lc_args& a, cudaStream_t st, int lim) {
    const size_t smem = lm3_smem_bytes(a.N);
    static std::atomic<bool> configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        const cudaError_t e = cudaFuncSetAttribute(lc_lm3_kernel<kLm3NT, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim > 0 ? lim : static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    const int mask = getenv("LC_B200_NO_TMA") ? 0 : ((lm3_planar(a.pts3d, a.N) ? 1 : 0) | (lm3_planar(a.pts2d, a.N) ? 2 : 0) | (lm3_planar(a.weights, a.N) ? 4 : 0));
    lc_lm3_kernel<kLm3NT, CTAS><<<a.B, kLm3NT, smem, st>>>(a, round_up4(a.N), mask);
    note_kernel("lc::lc_lm3_kernel<%d,LM,X in smem,%d CTAs/SM>", kLm3NT, CTAS);
    return static_cast<int>(cudaGetLastError());
}

int launch_lm3(const lc_args& a, cudaStream_t st) {
    const int lim = lm3_optin_smem();
    return lm3_ctas(a, lim) == 4 ? launch_lm3_t<4>(a, st, lim) : launch_lm3_t<3>(a, st, lim);
}

}  // namespace lc
