// The shared-memory resident kernel template (see lc_resident.cu for the design notes) and its launcher.  Instantiated in
// two translation units so they compile in parallel: lc_resident.cu (scalar point loops, any strides) and
// lc_resident_vec.cu (vectorised point loops for planar 16-byte aligned slabs, lc_vec.cuh).
#pragma once

#include <atomic>

#include "lc_vec.cuh"

// the planar-weights LM pass (lm_eval_pass_planar) in the vectorised CTA-per-pose kernels: 0 = keep the strided pass
#ifndef LC_RES_PLANAR_LM
#define LC_RES_PLANAR_LM 0
#endif

namespace lc {

// TM = true: the model points live in tensor memory (lc_resident.cuh: XAcc), shared memory holds only x -> ec.
#ifdef LC_TIMING
__device__ int g_live_ctas[256];   // CTAs currently resident per SM (tools/phase_timing.py: measured concurrency)
#endif

template <int CL>
__device__ __forceinline__ const lc_args& pick_args(const lc_args& param, const lc_args& shared) {
    if constexpr (CL == 1) return param;
    else return shared;
}

// CL > 1: one pose per thread-block CLUSTER of CL CTAs (launched with cudaLaunchAttributeClusterDimension); CTA `rank` holds the
// points [rank * npad, rank * npad + npad) and sees them through its own copy of the arguments (ClusterShared::args) whose
// point-indexed pointers are advanced to its share, so every point loop below runs unchanged on the share.  The reductions are
// completed across the cluster (Clu::combine), the serial sections run redundantly in every CTA, rank 0 writes the per-pose outputs.
// Used for the poses of a launch that would otherwise leave most CTA slots of the last wave empty (lc_resident.cu).
// pdl: bit 0 = let the next kernel of the stream start as soon as this grid's CTAs are resident (griddepcontrol.launch_dependents),
// bit 1 = this grid was launched that way behind another one: wait for it before exiting (stream order of completion).
template <int NT, int MODE, bool TM, bool VEC, int CL>
__global__ void __launch_bounds__(NT, NT <= 128 ? 4 : 2) lc_resident_kernel(const lc_args a_in, int npad, int tma_mask, int n_max, int pose_base, int pdl) {
    static_assert(!(TM && VEC), "the vectorised phase keeps the model points in shared memory");
    static_assert(CL == 1 || (VEC && !TM), "cluster-split poses use the vectorised shared-memory kernel");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PoseShared& s = *reinterpret_cast<PoseShared*>(smem_raw);
    const ResLayout l = TM ? res_layout_tm(smem_raw, npad) : res_layout(smem_raw, npad);
    const int b = pose_base + static_cast<int>(blockIdx.x) / CL;
    const int tid = threadIdx.x;
    if (pdl & 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int n_total = a_in.n_points ? min(max(a_in.n_points[b], 0), a_in.N) : a_in.N;
    if (CL == 1 && n_total > n_max) return;   // ragged batch split by n_points: this pose belongs to the streaming launch (lc_abi.cu)
    ClusterShared& cs = *reinterpret_cast<ClusterShared*>(smem_raw + resident_smem_bytes(CL > 1 ? npad : 0));
    Clu<CL> cl;
    cl.n_total = n_total;
    int n = n_total;
    if constexpr (CL > 1) {
        const int rank = static_cast<int>(cluster_ctarank());
        const int off = rank * npad;
        if (tid == 0) {
            cs.args = a_in;
            lc_args& A = cs.args;
            A.N = min(max(a_in.N - off, 0), npad);
            auto adv = [&](lc_view& v) { if (v.ptr) v.ptr = static_cast<float*>(v.ptr) + off * v.stride[1]; };
            adv(A.pts3d); adv(A.pts2d); adv(A.weights); adv(A.valid); adv(A.g_pts3d); adv(A.g_pts2d); adv(A.g_weights);
        }
        __syncthreads();
        n = min(max(n_total - off, 0), npad);
        cl.xch = cs.xch; cl.parity = 0; cl.leader = rank == 0;
    }
    const lc_args& a = pick_args<CL>(a_in, cs.args);
    auto finish = [&]() {
        if (CL > 1) cluster_sync_all();   // no CTA leaves while a peer may still read its exchange buffer
        if (pdl & 2) asm volatile("griddepcontrol.wait;" ::: "memory");
    };
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
    if (VEC && tid < 4) {
        // the vectorised passes touch whole groups of four points: clear the (at most three) staged slots beyond n
        const int i = n + tid;
        if (i < ((n + 3) & ~3)) {
            if (!TM) { l.A0[i] = 0.f; l.A1[i] = 0.f; l.A2[i] = 0.f; }
            l.B0[i] = 0.f; l.B1[i] = 0.f;
        }
    }
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
    // weights that need a per-point transform in the solve (sqrt of inverse variances, nan_to_num)
    const bool wgen = sanitize || a.weight_mode == LC_W_ICOV_DIAG;

    // =========================== LM solve (fp64) ===========================
    if (MODE & MODE_LM) {
        LmState& L = s.lm;
#ifdef LC_TIMING
        double* trace = nullptr;
#else
        double* trace = (a.trace && cl.leader) ? a.trace + (int64_t)b * (a.max_iter + 2) * 4 : nullptr;
#endif
        bool solved = false;
        if (n_total >= 3) {
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
                if (VEC && kind != CTL_EVAL_COST && (a.flags & LC_FLAG_LM_MIXED)) {
                    if (wgen) lm_eval_pass_mixed<NT, true>(a, s, l, b, n, sanitize, cl);
                    else lm_eval_pass_mixed<NT, false>(a, s, l, b, n, sanitize, cl);
                } else if (VEC && LC_RES_PLANAR_LM) {
                    if (kind == CTL_EVAL_COST) {
                        if (wgen) lm_eval_pass_planar<NT, false, true>(a, s, l, b, n, sanitize, cl);
                        else lm_eval_pass_planar<NT, false, false>(a, s, l, b, n, sanitize, cl);
                    } else {
                        if (wgen) lm_eval_pass_planar<NT, true, true>(a, s, l, b, n, sanitize, cl);
                        else lm_eval_pass_planar<NT, true, false>(a, s, l, b, n, sanitize, cl);
                    }
                } else if (kind == CTL_EVAL_COST) lm_eval_pass_res<NT, false>(a, s, l, b, n, sanitize, xs, cl);
                else lm_eval_pass_res<NT, true>(a, s, l, b, n, sanitize, xs, cl);
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
        if (tid == 0) lm_write_result<float>(a, s, b, n_total, solved, cl.leader);
        __syncthreads();
    }
#ifdef LC_TIMING
    if (!(MODE & MODE_LC)) {
        if (tid == 0 && a.trace) { double* tr = a.trace + (int64_t)b * (a.max_iter + 2) * 4; for (int k = 0; k < 7; ++k) tr[k] = (double)s.fin_timing[k]; tr[7] = (double)(clock64() - t_begin);
            for (int k = 0; k < 6; ++k) tr[48 + k] = (double)s.lm.tm[k]; tr[60] = live_at_start; }
        if (tid == 0) atomicAdd(&g_live_ctas[smid_], -1);
        tmem_release();
        finish();
        return;
    }
#else
    if (!(MODE & MODE_LC)) { tmem_release(); finish(); return; }
#endif

    // =========================== LC loss ===========================
    if (VEC) lc_phase_vec<NT>(a, s, l, b, n, cl);
    else {
        const DirectWeights wsrc{static_cast<const float*>(a.weights.ptr) + b * a.weights.stride[0], a.weights.stride[1], a.weights.stride[2]};
        DirectSink sink{a, b};
        lc_phase_res<NT>(a, s, l, b, n, wsrc, sink, xs);
    }
    tmem_release();
    finish();
#ifdef LC_TIMING
    if (tid == 0 && a.trace) { double* tr = a.trace + (int64_t)b * (a.max_iter + 2) * 4; for (int k = 0; k < 7; ++k) tr[k] = (double)s.fin_timing[k]; tr[7] = (double)(clock64() - t_begin);
        for (int k = 0; k < 40; ++k) tr[8 + k] = (double)(s.marks[k] - s.marks[0]); tr[60] = live_at_start; }
    if (tid == 0) atomicAdd(&g_live_ctas[smid_], -1);
#endif
}


// What one launch covers: poses [pose_base, pose_base + count) of the batch, each on a cluster of `cl` CTAs (1 = one CTA per pose).
struct ResLaunch {
    int cap;         // > 0: ragged batch, only poses with n_points <= cap are processed here (cl == 1)
    int max_smem, tma_mask;
    int pose_base, count, cl;
    int pdl;         // bit 0: trigger dependents at CTA start; bit 1: launched programmatically behind the previous kernel
};

template <int NT, int MODE, bool TM, bool VEC, int CL>
static int launch_res_t(const lc_args& a, cudaStream_t st, const ResLaunch& r) {
    // points held on chip per CTA
    const int n_res = CL > 1 ? round_up4((a.N + CL - 1) / CL) : (r.cap > 0 ? r.cap : a.N);
    const size_t smem = (TM ? resident_smem_bytes_tm(n_res) : resident_smem_bytes(n_res)) + (CL > 1 ? sizeof(ClusterShared) : 0);
    static std::atomic<bool> configured[64];   // per instantiation and per device (the opt-in smem limit is a per-device attribute)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        const cudaError_t e = cudaFuncSetAttribute(lc_resident_kernel<NT, MODE, TM, VEC, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, r.max_smem);
        if (e != cudaSuccess) return static_cast<int>(e);
        if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(r.count) * CL);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    unsigned na = 0;
    if (CL > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = CL; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
    }
    if (r.pdl & 2) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    if (getenv("LC_B200_DEBUG_OCC")) {   // diagnostic: resident CTAs per SM / clusters per device for this launch
        int ncl = -1, nb = -1;
        if (CL > 1) cudaOccupancyMaxActiveClusters(&ncl, lc_resident_kernel<NT, MODE, TM, VEC, CL>, &cfg);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lc_resident_kernel<NT, MODE, TM, VEC, CL>, NT, smem);
        fprintf(stderr, "[lc_b200] NT=%d MODE=%d CL=%d smem=%zu grid=%u: max active clusters %d, CTAs/SM %d\n", NT, MODE, CL, smem, cfg.gridDim.x, ncl, nb);
    }
    const cudaError_t e = cudaLaunchKernelEx(&cfg, lc_resident_kernel<NT, MODE, TM, VEC, CL>, a, round_up4(n_res), r.tma_mask,
                                             r.cap > 0 ? r.cap : 0x7fffffff, r.pose_base, r.pdl);
    if (CL > 1) note_kernel("lc::lc_resident_kernel<%d,%s,smem,vec4,cluster%d>", NT, MODE == MODE_LM ? "LM" : (MODE == MODE_LC ? "LC" : "LM|LC"), CL);
    else note_kernel("lc::lc_resident_kernel<%d,%s,%s,%s>", NT, MODE == MODE_LM ? "LM" : (MODE == MODE_LC ? "LC" : "LM|LC"), TM ? "TMEM" : "smem", VEC ? "vec4" : "scalar");
    return static_cast<int>(e != cudaSuccess ? e : cudaGetLastError());
}

// one (threads, mode, storage) dispatcher per translation unit
template <bool VEC, int CL>
static int launch_res_any(const lc_args& a, int mode, int nt, bool tm, cudaStream_t st, const ResLaunch& r) {
#define LC_RES_CASE(NT_, MODE_, TM_) return launch_res_t<NT_, MODE_, TM_, VEC, CL>(a, st, r)
    if (mode == MODE_LM) {
        if (nt == 128) LC_RES_CASE(128, MODE_LM, false);
        if (nt == 192) LC_RES_CASE(192, MODE_LM, false);
        LC_RES_CASE(256, MODE_LM, false);
    }
    if (mode == MODE_LC) {
        if constexpr (!VEC) { if (tm) LC_RES_CASE(128, MODE_LC, true); }
        if (nt == 128) LC_RES_CASE(128, MODE_LC, false);
        LC_RES_CASE(256, MODE_LC, false);
    }
    if (nt == 128) LC_RES_CASE(128, MODE_LM | MODE_LC, false);
    LC_RES_CASE(256, MODE_LM | MODE_LC, false);
#undef LC_RES_CASE
}

int launch_res_scalar(const lc_args& a, int mode, int nt, bool tm, cudaStream_t st, const ResLaunch& r);
int launch_res_vec(const lc_args& a, int mode, int nt, bool tm, cudaStream_t st, const ResLaunch& r);
int launch_res_cluster2(const lc_args& a, int mode, int nt, cudaStream_t st, const ResLaunch& r);   // lc_resident_cluster.cu

}  // namespace lc
