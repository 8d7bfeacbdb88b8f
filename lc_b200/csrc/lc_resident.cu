// lc_b200 — shared-memory resident sm_100a kernel for the LC hot path (the headline path).
//
// One CTA per pose.  The pose's correspondences are read from HBM exactly once, converted to a planar
// fp32 layout in shared memory (20 B/point: X[3], x[2]); the weights (8 B/point) are re-read from L2 by the
// passes that need them, which keeps the footprint at N = 4096 small enough for TWO CTAs per SM, so one
// CTA's serial sections (trust-region step, 6x6 algebra) and loads overlap the other's point passes.  The only
// other HBM traffic is the gradient write-back.  DRAM bytes = the algorithmic 48*N + O(1) per pose.
//
//   LM phase  (MODE & 1): fp64.  Each trust-region iteration is one fused pass (cost + J'^T J' + J'^T r in
//             the left basis, 28 accumulators/thread) + one multi-value CTA reduction + a single-thread
//             6x6 step.  The converging iteration evaluates the cost only (lc_pose.cuh: lm_advance).
//   LC phase  (MODE & 2): pass 1 in fp64 computes the camera-frame point P and the clamped error ec once
//             (as q = R X) and stores them as fp32 IN PLACE over X and x; passes 2-4 (robust statistics, H/G/b
//             accumulation, reverse pass) are fp32 per point with per-thread partial sums of <= N/NT terms
//             that are reduced across the CTA in fp64; the 6x6 algebra is fp64 (lc_pose.cuh).
//
// Staging uses cp.async (LDGSTS): every element of the pose is in flight at once while the pose constants are set up.
// Loss-only launches (MODE == 2) stage only X and x (20 B/point) and re-read the weights from L2, so two CTAs fit
// per SM and one CTA's load / 6x6 sections overlap the other's point passes.
//
// Restrictions (anything else takes the streaming kernel): fp32 tensors, diagonal weights, 64 < N <= limit.
#include <atomic>

#include "lc_resident.cuh"

namespace lc {

// TM = true: the model points live in tensor memory (lc_resident.cuh: XAcc), shared memory holds only x -> ec.
#ifdef LC_TIMING
__device__ int g_live_ctas[256];   // CTAs currently resident per SM (tools/phase_timing.py: measured concurrency)
#endif

template <int NT, int MODE, bool TM>
__global__ void __launch_bounds__(NT, NT <= 128 ? 4 : 2) lc_resident_kernel(const lc_args a, int npad, int tma_mask, int n_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PoseShared& s = *reinterpret_cast<PoseShared*>(smem_raw);
    const ResLayout l = TM ? res_layout_tm(smem_raw, npad) : res_layout(smem_raw, npad);
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int n = a.n_points ? min(max(a.n_points[b], 0), a.N) : a.N;
    if (n > n_max) return;   // ragged batch split by n_points: this pose belongs to the streaming launch (lc_abi.cu)
    const bool sanitize = (MODE & MODE_LM) && (a.flags & LC_FLAG_NAN_TO_NUM);
    uint32_t tb = 0;
    if (TM) {
        // kTmemCols columns of tensor memory for this CTA (4 CTAs x 128 = all 512 columns of the SM); a CTA that finds
        // none free waits inside tcgen05.alloc until a resident CTA releases its columns
        if (tid < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "n"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tb = s.tmem_base + ((static_cast<uint32_t>(tid >> 5) & 3u) * 32u << 16);
    }
    auto tmem_release = [&]() {
        if (TM) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s.tmem_base), "n"(kTmemCols) : "memory");
        }
    };
#ifdef LC_TIMING
    if (tid == 0) { for (int k = 0; k < 8; ++k) s.fin_timing[k] = 0; for (int k = 0; k < 6; ++k) s.lm.tm[k] = 0; }
    const long long t_begin = clock64();
    unsigned smid_;
    asm("mov.u32 %0, %%smid;" : "=r"(smid_));
    int live_at_start = 0;
    if (tid == 0) live_at_start = atomicAdd(&g_live_ctas[smid_], 1) + 1;
#endif
    { LC_TIC(tq1);

    // ---- stage the correspondences.  Planar, 16-byte aligned arrays (what the dense call site produces, tma_mask
    //      bit 0 = pts3d, bit 1 = pts2d) go through the TMA: one 1-D bulk copy per component slab, issued by one
    //      thread, completing on an mbarrier.  Anything else: one 4-byte cp.async per element (any strides).  Either
    //      way every byte of the pose is in flight while the pose constants are set up. ----
    {
        const float* p3 = static_cast<const float*>(a.pts3d.ptr) + b * a.pts3d.stride[0];
        const float* p2 = static_cast<const float*>(a.pts2d.ptr) + b * a.pts2d.stride[0];
        const int64_t s3n = a.pts3d.stride[1], s3c = a.pts3d.stride[2], s2n = a.pts2d.stride[1], s2c = a.pts2d.stride[2];
        const bool tma3 = !TM && (tma_mask & 1) != 0, tma2 = (tma_mask & 2) != 0;
        if (tma_mask) {
            if (tid == 0) mbar_init(&s.tma_bar, 1);
            __syncthreads();
            if (tid == 0) {
                const unsigned slab = static_cast<unsigned>(min(a.N, npad)) * 4u;   // npad < N only for split ragged batches
                mbar_expect_tx(&s.tma_bar, slab * ((tma3 ? 3u : 0u) + (tma2 ? 2u : 0u)));
                if (tma3) { tma_load_1d(l.A0, p3, slab, &s.tma_bar); tma_load_1d(l.A1, p3 + s3c, slab, &s.tma_bar); tma_load_1d(l.A2, p3 + 2 * s3c, slab, &s.tma_bar); }
                if (tma2) { tma_load_1d(l.B0, p2, slab, &s.tma_bar); tma_load_1d(l.B1, p2 + s2c, slab, &s.tma_bar); }
                if (tma_mask & 4) {   // planar 16-byte aligned weights: two slabs into L2 ahead of their first use
                    const float* pw = static_cast<const float*>(a.weights.ptr) + b * a.weights.stride[0];
                    l2_prefetch_bulk(pw, slab);
                    l2_prefetch_bulk(pw + a.weights.stride[2], slab);
                }
            }
        }
        if (!(tma3 && tma2)) {
            for (int i = tid; i < npad; i += NT) {
                if (i < n) {
                    if (!TM && !tma3) { cp_async4(l.A0 + i, p3 + i * s3n); cp_async4(l.A1 + i, p3 + i * s3n + s3c); cp_async4(l.A2 + i, p3 + i * s3n + 2 * s3c); }
                    if (!tma2) { cp_async4(l.B0 + i, p2 + i * s2n); cp_async4(l.B1 + i, p2 + i * s2n + s2c); }
                } else {
                    if (!TM && !tma3) { l.A0[i] = 0.f; l.A1[i] = 0.f; l.A2[i] = 0.f; }
                    if (!tma2) { l.B0[i] = 0.f; l.B1[i] = 0.f; }
                }
            }
        }
        if (TM) {
            // model points: global -> registers -> the thread's TMEM lane, sixteen points (48 loads) in flight per thread
            const int wbase = tid & ~31;
            for (int k0 = 0; k0 * NT + wbase < n; k0 += 16) {
                float v[16][3];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int i = tid + (k0 + j) * NT;
                    const bool live = i < n;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float x = live ? p3[i * s3n + c * s3c] : 0.f;
                        v[j][c] = sanitize ? nan_to_num_f(x) : x;
                    }
                }
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if ((k0 + j) * NT + wbase < n) tmem_st4(tb + 4u * (k0 + j), v[j][0], v[j][1], v[j][2], 0.f);   // warp-uniform predicate
            }
            tmem_wait_st();
        }
    }
    // ---- pose constants ----
    if (tid < 9) {
        float v = ldf(a.K, b * a.K.stride[0] + (tid / 3) * a.K.stride[1] + (tid % 3) * a.K.stride[2]);
        s.K[tid] = sanitize ? nan_to_num_f(v) : v;
    } else if (tid < 16) {
        float v = ldf(a.pose, b * a.pose.stride[0] + (tid - 9) * a.pose.stride[1]);
        s.pose[tid - 9] = sanitize ? nan_to_num_f(v) : v;
    }
    if (MODE & MODE_LC) {
        for (int k = tid; k < 24; k += NT)
            s.bbox[k] = ldf(a.bbox, b * a.bbox.stride[0] + (k / 3) * a.bbox.stride[1] + (k % 3) * a.bbox.stride[2]);
    }
    cp_async_commit_wait_all();
    if (tma_mask) mbar_wait(&s.tma_bar, 0);
    if (sanitize) {
        // solver prologue on the thread's own elements: nan_to_num (cer_solver.py:27-29)
        for (int i = tid; i < n; i += NT) {
            if (!TM) { l.A0[i] = nan_to_num_f(l.A0[i]); l.A1[i] = nan_to_num_f(l.A1[i]); l.A2[i] = nan_to_num_f(l.A2[i]); }
            l.B0[i] = nan_to_num_f(l.B0[i]); l.B1[i] = nan_to_num_f(l.B1[i]);
        }
    }
    __syncthreads();
    LC_TOC(tq1, 0); }
    const XAcc<TM> xs{l, tb};

    // =========================== LM solve (fp64) ===========================
    if (MODE & MODE_LM) {
        LmState& L = s.lm;
#ifdef LC_TIMING
        double* trace = nullptr;
#else
        double* trace = a.trace ? a.trace + (int64_t)b * (a.max_iter + 2) * 4 : nullptr;
#endif
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
                { LC_TIC(tq2);
                if (kind == CTL_EVAL_COST) lm_eval_pass_res<NT, false>(a, s, l, b, n, sanitize, xs);
                else lm_eval_pass_res<NT, true>(a, s, l, b, n, sanitize, xs);
                LC_TOC(tq2, 1); }
                LC_TIC(tq3);
                if (tid == 0)
                    lm_advance(L, s.fin, kind, first, a.max_iter, a.function_tolerance, (a.flags & LC_FLAG_TOL_NEEDS_SUCCESS) != 0, trace);
                first = false;
                __syncthreads();
                LC_TOC(tq3, 2);
                if (L.ctl == CTL_STOP) break;
            }
            solved = L.term == TERM_CONVERGENCE;
        }
        if (tid == 0) lm_write_result<float>(a, s, b, n, solved);
        __syncthreads();
    }
#ifdef LC_TIMING
    if (!(MODE & MODE_LC)) {
        if (tid == 0 && a.trace) { double* tr = a.trace + (int64_t)b * (a.max_iter + 2) * 4; for (int k = 0; k < 7; ++k) tr[k] = (double)s.fin_timing[k]; tr[7] = (double)(clock64() - t_begin);
            for (int k = 0; k < 6; ++k) tr[48 + k] = (double)s.lm.tm[k]; tr[60] = live_at_start; }
        if (tid == 0) atomicAdd(&g_live_ctas[smid_], -1);
        tmem_release();
        return;
    }
#else
    if (!(MODE & MODE_LC)) { tmem_release(); return; }
#endif

    // =========================== LC loss ===========================
    const DirectWeights wsrc{static_cast<const float*>(a.weights.ptr) + b * a.weights.stride[0], a.weights.stride[1], a.weights.stride[2]};
    DirectSink sink{a, b};
    lc_phase_res<NT>(a, s, l, b, n, wsrc, sink, xs);
    tmem_release();
#ifdef LC_TIMING
    if (tid == 0 && a.trace) { double* tr = a.trace + (int64_t)b * (a.max_iter + 2) * 4; for (int k = 0; k < 7; ++k) tr[k] = (double)s.fin_timing[k]; tr[7] = (double)(clock64() - t_begin);
        for (int k = 0; k < 40; ++k) tr[8 + k] = (double)(s.marks[k] - s.marks[0]); tr[60] = live_at_start; }
    if (tid == 0) atomicAdd(&g_live_ctas[smid_], -1);
#endif
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// opt-in shared-memory limit of the CURRENT device, cached per device index (the ABI is re-entrant: relaxed atomics,
// a race only repeats the query)
static std::atomic<int> g_max_smem[64];

static int max_optin_smem() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    const bool cached = dev >= 0 && dev < 64;
    if (cached && (v = g_max_smem[dev].load(std::memory_order_relaxed)) > 0) return v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cached) g_max_smem[dev].store(v, std::memory_order_relaxed);
    return v;
}

bool resident_supported(const lc_args& a, int mode) {
    if (a.dtype != LC_F32 || a.N < kResidentMinN) return false;
    if ((mode & MODE_LM) && a.weight_mode != LC_W_ICOV_DIAG && a.weight_mode != LC_W_INV_STD) return false;
    const size_t need = resident_smem_bytes(a.N);
    return need <= static_cast<size_t>(max_optin_smem());
}

// 128 threads x 4 CTAs per SM while four poses fit in shared memory (N <= 2048: 40 KB + 15 KB each), 256 threads x 2
// CTAs above.  Both give 16 warps/SM at 128 registers; the finer granularity hides the per-pose serial sections
// better (measured on B200 at equal total points: N = 1024 P3 511 vs 651 us, N = 2048 389 vs 442 us; N = 2900, where
// only three 128-thread CTAs fit, prefers 256).
static int resident_threads_for(int n, int mode) {
    if (const char* e = getenv("LC_B200_RES_NT")) return atoi(e);   // tuning knob for benchmarks
    if (n <= 2048) return 128;
    // solve-only: 192 threads x 164 registers keep the 28 fp64 accumulators + pose constants of the LM pass out of local
    // memory (the 256-thread build spills 8 of them); 12 warps/SM are enough for the fp64-pipe-bound pass (P2 196 -> 189 us).
    // The loss passes are issue-bound and want the 16 warps (P1 +18 %, P3 +6 % with 192).
    return mode == MODE_LM ? 192 : 256;
}

// A (B,N,C) fp32 view can be staged by 1-D TMA bulk copies when every component slab is contiguous (point stride 1)
// and 16-byte aligned for every pose: base pointer, batch stride, component stride and N all multiples of 16 bytes.
static bool tma_ok(const lc_view& v, int n) {
    return v.stride[1] == 1 && (n % 4) == 0 && (reinterpret_cast<uintptr_t>(v.ptr) % 16) == 0 && (v.stride[0] % 4) == 0 &&
           (v.stride[2] % 4) == 0;
}
static int tma_mask_for(const lc_args& a) {
    if (getenv("LC_B200_NO_TMA")) return 0;
    const int m = (tma_ok(a.pts3d, a.N) ? 1 : 0) | (tma_ok(a.pts2d, a.N) ? 2 : 0);
    const bool w_planar = (a.weight_mode == LC_W_ICOV_DIAG || a.weight_mode == LC_W_INV_STD) && tma_ok(a.weights, a.N);
    return m | ((m && w_planar && !getenv("LC_B200_NO_L2PF")) ? 4 : 0);
}

template <int NT, int MODE, bool TM = false>
static int launch_res_t(const lc_args& a, cudaStream_t st, int cap) {
    const int n_res = cap > 0 ? cap : a.N;   // points held on chip per pose
    const size_t smem = TM ? resident_smem_bytes_tm(n_res) : resident_smem_bytes(n_res);
    static std::atomic<bool> configured[64];   // per instantiation and per device (the opt-in smem limit is a per-device attribute)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        const cudaError_t e = cudaFuncSetAttribute(lc_resident_kernel<NT, MODE, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin_smem());
        if (e != cudaSuccess) return static_cast<int>(e);
        if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    lc_resident_kernel<NT, MODE, TM><<<a.B, NT, smem, st>>>(a, round_up4(n_res), tma_mask_for(a), cap > 0 ? cap : 0x7fffffff);
    note_kernel("lc::lc_resident_kernel<%d,%s,%s>", NT, MODE == MODE_LM ? "LM" : (MODE == MODE_LC ? "LC" : "LM|LC"), TM ? "TMEM" : "smem");
    return static_cast<int>(cudaGetLastError());
}

// Which modes take the tensor-memory variant (measured on B200 at B = 1024 x N = 4096, see profiles/README.md): the loss-only
// kernel gains from four finer-grained CTAs per SM; LC_B200_TMEM = 0 | 1 | lc overrides for A/B runs.
static bool tmem_enabled(int mode) {
    if (const char* e = getenv("LC_B200_TMEM")) return e[0] == '1' || (e[0] == 'l' && mode == MODE_LC);
    return mode == MODE_LC;
}

template <int MODE>
static int launch_res_m(const lc_args& a, cudaStream_t st, int cap) {
    // 256 threads x 2 CTAs per SM (20 B/point of shared memory): one CTA's 6x6 / trust-region sections and loads
    // overlap the other CTA's point passes
    // 2048 < N <= 4096: four 128-thread CTAs per SM with the model points in tensor memory (lc_resident.cuh: XAcc)
    if (cap == 0 && a.N > 2048 && a.N <= kTmemMaxN && tmem_enabled(MODE)) return launch_res_t<128, MODE, true>(a, st, cap);
    const int nt = resident_threads_for(cap > 0 ? cap : a.N, MODE);
    if (nt == 128) return launch_res_t<128, MODE>(a, st, cap);
    if (nt == 192) return launch_res_t<192, MODE>(a, st, cap);
    return launch_res_t<256, MODE>(a, st, cap);
}

// cap > 0: ragged batch whose padded N exceeds the resident limit; only poses with n_points <= cap are processed here
int launch_resident_pose(const lc_args& a, int mode, cudaStream_t st, int cap) {
    switch (mode) {
        case MODE_LM: return launch_res_m<MODE_LM>(a, st, cap);
        case MODE_LC: return launch_res_m<MODE_LC>(a, st, cap);
        default: return launch_res_m<MODE_LM | MODE_LC>(a, st, cap);
    }
}

// Ragged batches (n_points given) whose padded N does not fit: the poses with n_points <= cap still can take the resident
// kernel.  cap = the largest point count that leaves two CTAs per SM; 0 when the split does not apply.
int resident_split_capacity(const lc_args& a, int mode) {
    if (!a.n_points || a.dtype != LC_F32 || a.N < kResidentMinN) return 0;
    if ((mode & MODE_LM) && a.weight_mode != LC_W_ICOV_DIAG && a.weight_mode != LC_W_INV_STD) return 0;
    const size_t per_cta = static_cast<size_t>(max_optin_smem()) / 2;
    const size_t fixed = resident_smem_bytes(0) + 1024;
    if (per_cta <= fixed) return 0;
    const int cap = static_cast<int>((per_cta - fixed) / 20) & ~3;
    return cap >= kResidentMinN && cap < a.N ? cap : 0;
}

}  // namespace lc
