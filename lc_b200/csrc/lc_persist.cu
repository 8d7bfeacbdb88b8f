// lc_b200 — persistent, software-pipelined sm_100a kernel for the LC hot path at large N (the headline configuration).
//
// Why: in the CTA-per-pose kernels (lc_resident_kernel.cuh) every pose has serial sections (the 6x6 forward / reverse algebra,
// the trust-region step) during which one warp works and the rest of the CTA waits at a barrier; a second resident CTA fills
// the SM only partly, the serial warp itself is starved (B200 issues the highest warp id first, so the busy warps of the other
// CTA always win: measured ~70 cycles per dependent instruction), and 1024 poses on 296 CTA slots leave a 0.46-wave tail.
//
// Here ONE CTA per SM lives for the whole launch and walks its share of the poses (pose = blockIdx.x + k gridDim.x: 6.92
// poses per SM at B = 1024, so the tail is one pose of seven on 12 of the 148 SMs).  The CTA keeps TWO poses in flight, each
// with its own shared-memory buffer (PoseShared + the planar fp32 arrays of lc_resident.cuh):
//   * warps 0..kPW-1 are WORKERS: they execute the parallel phases — one point pass per phase — alternating
//     between the two poses;
//   * the last warp is the SERIAL warp: it has the highest warp id of the CTA, hence issue priority, and executes everything that
//     is O(1) per pose: pose setup, the sum of the per-warp partial sums, thresholds, the trust-region step, the
//     register-resident 6x6 sections (lc_six_fast), the TMA loads of the next pose and the TMA stores of the gradients.
// While the serial warp works on pose A the workers run a pass of pose B, and vice versa: the serial sections disappear from
// the critical path as long as they are shorter than a pass.  Hand-over uses named barriers (bar.arrive / bar.sync):
//   barrier 2+c  "parallel phase of context c done"   workers arrive, serial warp syncs
//   barrier 4+c  "serial phase of context c done"     serial warp arrives, workers sync
// Staging is TMA in (cp.async.bulk + mbarrier, issued by the serial warp as soon as a buffer is free, the L2 prefetch of the
// weight slabs with it) and TMA out (the gradients are written in place over q / ec and leave as bulk stores).
//
// Point passes: lc_vec.cuh (four consecutive points per thread, packed fp32 FFMA2, fp64 LM pass).  Restrictions (anything
// else takes the CTA-per-pose kernels): fp32, diagonal weights, planar 16-byte aligned pts3d / pts2d / weights / gradients,
// two pose buffers must fit in shared memory (N <= ~4.6k), N large enough to feed 512 threads (N >= 2048).
#include <atomic>

#include "lc_vec.cuh"

namespace lc {

// Worker warps per CTA.  Registers are allocated per group of four warps, so 16 workers + the serial warp (17 warps -> 20)
// would leave 96 registers per thread; 11 + 1 warps get 168 (no spills in any pass, and 352 threads divide the 1024 groups /
// 4096 points of the headline shape into 3 / 12 rounds at 97 %), 15 + 1 get 128.
#ifndef LC_PERSIST_WARPS
#define LC_PERSIST_WARPS 11
#endif
constexpr int kPW = LC_PERSIST_WARPS;   // worker warps
constexpr int kWT = kPW * 32;           // worker threads
constexpr int kPT = kWT + 32;           // + the serial warp (the highest warp id)

enum { PH_DONE = 0, PH_LM = 1, PH_LC1 = 2, PH_LC2 = 3, PH_LC3 = 4, PH_LC4 = 5, PH_LC4G = 6 };

struct PCtx {
    PoseShared ps;
    int pose;            // pose index of this context, -1 = none
    int n;               // live correspondences
    int phase;           // next parallel phase (PH_*)
    int kind;            // LM evaluation kind (CTL_EVAL_*) of the next PH_LM
    int first;           // the next lm_advance is the first of this pose
    int wait_tma;        // the next parallel phase is the first to touch the staged arrays: wait for the TMA
    int tma_parity;      // mbarrier phase parity of that wait
    int tma_out;         // the gradients of this pose leave as TMA stores (fast pass 4)
    float d0, d1, sq0, sq1;
    double vcnt;
};

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

inline size_t persist_ctx_bytes(int npad) { return ((sizeof(PCtx) + 15) & ~size_t(15)) + sizeof(float) * 5 * static_cast<size_t>(npad); }

__device__ __forceinline__ ResLayout persist_layout(unsigned char* ctx_base, int npad) {
    float* f = reinterpret_cast<float*>(ctx_base + ((sizeof(PCtx) + 15) & ~size_t(15)));
    ResLayout l;
    l.A0 = f; l.A1 = f + npad; l.A2 = f + 2 * npad; l.B0 = f + 3 * npad; l.B1 = f + 4 * npad;
    return l;
}

// every warp: warp-level reduce-scatter of its V partial sums into red[warp][V]
template <int V>
__device__ __forceinline__ void warp_partials(double (&v)[V], double* red) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    warp_reduce_scatter<V>(v, lane);
#pragma unroll
    for (int k = 0; k < ReduceShape<V>::c5; ++k) {
        const int idx = orig_index<V>(k, lane);
        if (idx >= 0) red[warp * V + idx] = v[k];
    }
}
template <int V>
__device__ __forceinline__ void warp_partials_f(float (&v)[V], double* red) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    warp_reduce_scatter_f<V>(v, lane);
#pragma unroll
    for (int k = 0; k < ReduceShape<V>::c5; ++k) {
        const int idx = orig_index<V>(k, lane);
        if (idx >= 0) red[warp * V + idx] = static_cast<double>(v[k]);
    }
}
// serial warp: fin[j] = sum over the worker warps
template <int V>
__device__ __forceinline__ void sum_partials(const double* red, double* fin, int lane) {
    for (int j = lane; j < V; j += 32) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int w = 0; w + 1 < kPW; w += 2) { s0 += red[w * V + j]; s1 += red[(w + 1) * V + j]; }
        if (kPW & 1) s0 += red[(kPW - 1) * V + j];
        fin[j] = s0 + s1;
    }
    __syncwarp();
}

template <int MODE>
__global__ void __launch_bounds__(kPT, 1) lc_persist_kernel(const lc_args a, int npad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31;
    const bool serial = tid >= kWT;
    const size_t ctx_bytes = ((sizeof(PCtx) + 15) & ~size_t(15)) + sizeof(float) * 5 * static_cast<size_t>(npad);
    // context c lives at smem_raw + c * ctx_bytes; c is a RUN-TIME value everywhere below so that the phase code exists once
    // (two copies would not fit the instruction cache together with the serial sections)
    auto ctx_of = [&](int c) -> PCtx& { return *reinterpret_cast<PCtx*>(smem_raw + c * ctx_bytes); };
    auto lay_of = [&](int c) { return persist_layout(smem_raw + c * ctx_bytes, npad); };
    const bool sanitize = (MODE & MODE_LM) && (a.flags & LC_FLAG_NAN_TO_NUM);
    const bool wgen = sanitize || a.weight_mode == LC_W_ICOV_DIAG;
    const bool want_any = a.g_pts3d.ptr || a.g_pts2d.ptr || a.g_weights.ptr;

    if (tid == kWT) {
        mbar_init(&ctx_of(0).ps.tma_bar, 1);
        mbar_init(&ctx_of(1).ps.tma_bar, 1);
    }
    __syncthreads();

    if (!serial) {
        // =============================== workers ===============================
        unsigned active = 3u;   // bit c: context c still has poses
#ifdef LC_TIMING
        long long tw[16] = {0}, t_wait = 0;   // cycles of worker thread 0 per phase kind / waiting for the serial warp
#endif
        while (active) {
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                if (!(active & (1u << c))) continue;
#ifdef LC_TIMING
                const long long tb0 = clock64();
#endif
                bar_sync(4 + c, kPT);                       // the serial phase of context c is done
#ifdef LC_TIMING
                const long long tb1 = clock64();
                t_wait += tb1 - tb0;
#endif
                PCtx& x = ctx_of(c);
                PoseShared& s = x.ps;
                const int ph = x.phase;
                if (ph == PH_DONE) { active &= ~(1u << c); continue; }
                const ResLayout l = lay_of(c);
                const int b = x.pose, n = x.n;
                if (x.wait_tma) {
                    mbar_wait(&s.tma_bar, static_cast<unsigned>(x.tma_parity));
                    if (sanitize) {
                        // solver prologue on the thread's own elements: nan_to_num (cer_solver.py:27-29); the LM pass below uses
                        // the same element ownership (i = tid + k * kWT), later phases are separated by barriers
                        for (int i = tid; i < n; i += kWT) {
                            l.A0[i] = nan_to_num_f(l.A0[i]); l.A1[i] = nan_to_num_f(l.A1[i]); l.A2[i] = nan_to_num_f(l.A2[i]);
                            l.B0[i] = nan_to_num_f(l.B0[i]); l.B1[i] = nan_to_num_f(l.B1[i]);
                        }
                    }
                }
                if ((MODE & MODE_LM) && ph == PH_LM) {
                    double acc[28];
                    if (x.kind == CTL_EVAL_COST) {
                        if (wgen) lm_eval_accum_planar<kWT, false, true>(a, s, l, b, n, sanitize, tid, acc);
                        else lm_eval_accum_planar<kWT, false, false>(a, s, l, b, n, sanitize, tid, acc);
                    } else {
                        if (wgen) lm_eval_accum_planar<kWT, true, true>(a, s, l, b, n, sanitize, tid, acc);
                        else lm_eval_accum_planar<kWT, true, false>(a, s, l, b, n, sanitize, tid, acc);
                    }
                    warp_partials<28>(acc, s.red);
                } else if (MODE & MODE_LC) {
                    const VecIO io = make_vec_io(a, b);
                    if (ph == PH_LC1) {
                        float acc[4];
                        lc_pass1_vec<kWT>(a, s, l, io, n, tid, acc);
                        warp_partials_f<4>(acc, s.red);
                    } else if (ph == PH_LC2) {
                        float acc[2];
                        lc_pass2_vec<kWT>(l, io, n, tid, x.d0, x.d1, acc);
                        warp_partials_f<2>(acc, s.red);
                    } else if (ph == PH_LC3) {
                        const PointConsts pc = make_point_consts(s, x.d0, x.d1, x.sq0, x.sq1);
                        float accd[48];
                        lc_pass3_vec<kWT>(l, io, n, tid, pc, accd);
                        warp_partials_f<48>(accd, s.red);
                    } else if (ph == PH_LC4) {
                        const PointConsts pc = make_point_consts(s, x.d0, x.d1, x.sq0, x.sq1);
                        lc_pass4_vec<kWT>(a, s, l, io, n, tid, pc);
#ifndef LC_PERSIST_LEAN
                    } else if (ph == PH_LC4G) {
                        const XAcc<false> xs{l, 0u};
                        const DirectWeights wsrc{io.w0, 1, a.weights.stride[2]};
                        DirectSink sink{a, b};
                        lc_pass4_scalar<kWT>(a, s, l, n, x.d0, x.d1, x.sq0, x.sq1, wsrc, sink, xs);
#endif
                    }
                }
                __threadfence_block();
#ifdef LC_TIMING
                tw[ph] += clock64() - tb1;
#endif
                bar_arrive(2 + c, kPT);                     // the parallel phase of context c is done
            }
        }
#ifdef LC_TIMING
        if (tid == 0 && a.trace) {
            double* tr = a.trace + static_cast<int64_t>(blockIdx.x) * 64;
            for (int k = 0; k < 8; ++k) tr[k] = static_cast<double>(tw[k]);
            tr[8] = static_cast<double>(t_wait);
        }
#endif
        return;
    }

    // =============================== serial warp ===============================
    int next_k0 = 0, next_k1 = 1;     // context c walks the poses blockIdx.x + (2 j + c) gridDim.x
    unsigned parity = 0u, active = 3u;   // bit c: mbarrier phase parity / context c still has poses

#ifdef LC_TIMING_SIX2
    long long six_t[2] = {0, 0};
#endif
    bool need_lc_setup = false, need_init = false;   // requests of the serial code below, served at ONE place each (code size)
    // assign the next pose of context c: constants, TMA loads, LM start.  Sets x.phase (PH_DONE when there is none) or asks for
    // the LC setup.
    auto init_pose = [&](int c) {
        PCtx& x = ctx_of(c);
        PoseShared& s = x.ps;
        const ResLayout l = lay_of(c);
        for (;;) {
            const int kk = c ? next_k1 : next_k0;
            const long long bl = static_cast<long long>(blockIdx.x) + static_cast<long long>(kk) * gridDim.x;
            if (c) next_k1 += 2; else next_k0 += 2;
            if (bl >= a.B) {
                if (lane == 0) { x.pose = -1; x.phase = PH_DONE; }
                active &= ~(1u << c);
                __syncwarp();
                return;
            }
            const int b = static_cast<int>(bl);
            const int n = a.n_points ? min(max(a.n_points[b], 0), a.N) : a.N;
            if ((MODE == MODE_LM) && n < 3) {
                // fewer than 3 correspondences (ceres.cpp:84-91): invalid, state = start; nothing to stage
                if (lane < 7) s.pose[lane] = ldf(a.pose, b * a.pose.stride[0] + lane * a.pose.stride[1]);
                __syncwarp();
                if (lane == 0) lm_write_result<float>(a, s, b, n, false);
                __syncwarp();
                continue;
            }
            if (lane == 0) {
                // every byte of the pose in flight while the constants are set up
                const float* p3 = static_cast<const float*>(a.pts3d.ptr) + b * a.pts3d.stride[0];
                const float* p2 = static_cast<const float*>(a.pts2d.ptr) + b * a.pts2d.stride[0];
                const unsigned slab = static_cast<unsigned>(a.N) * 4u;
                mbar_expect_tx(&s.tma_bar, slab * 5u);
                tma_load_1d(l.A0, p3, slab, &s.tma_bar); tma_load_1d(l.A1, p3 + a.pts3d.stride[2], slab, &s.tma_bar);
                tma_load_1d(l.A2, p3 + 2 * a.pts3d.stride[2], slab, &s.tma_bar);
                tma_load_1d(l.B0, p2, slab, &s.tma_bar); tma_load_1d(l.B1, p2 + a.pts2d.stride[2], slab, &s.tma_bar);
                const float* pw = static_cast<const float*>(a.weights.ptr) + b * a.weights.stride[0];
                l2_prefetch_bulk(pw, slab);
                l2_prefetch_bulk(pw + a.weights.stride[2], slab);
                x.pose = b; x.n = n; x.wait_tma = 1; x.tma_parity = static_cast<int>((parity >> c) & 1u); x.first = 1; x.tma_out = 0;
            }
            parity ^= 1u << c;
            if (lane < 9) {
                const float v = ldf(a.K, b * a.K.stride[0] + (lane / 3) * a.K.stride[1] + (lane % 3) * a.K.stride[2]);
                s.K[lane] = sanitize ? nan_to_num_f(v) : v;
            } else if (lane < 16) {
                const float v = ldf(a.pose, b * a.pose.stride[0] + (lane - 9) * a.pose.stride[1]);
                s.pose[lane - 9] = sanitize ? nan_to_num_f(v) : v;
            }
            if (MODE & MODE_LC) {
                if (lane < 24) s.bbox[lane] = ldf(a.bbox, b * a.bbox.stride[0] + (lane / 3) * a.bbox.stride[1] + (lane % 3) * a.bbox.stride[2]);
            }
            __syncwarp();
            if ((MODE & MODE_LM) && n >= 3) {
                if (lane == 0) {
                    LmState& L = s.lm;
                    quat_to_angle_axis(s.pose, L.x);
                    L.x[3] = s.pose[4]; L.x[4] = s.pose[5]; L.x[5] = s.pose[6];
                    lm_set_eval_point(L, L.x);
                    L.ctl = CTL_EVAL_FULL;
                    x.kind = CTL_EVAL_FULL; x.phase = PH_LM;
                }
                __syncwarp();
                return;
            }
            if (MODE & MODE_LM) {   // fused mode, fewer than 3 correspondences: the solve is invalid, the loss runs at the start pose
                if (lane == 0) lm_write_result<float>(a, s, b, n, false);
                __syncwarp();
            }
            need_lc_setup = true;
            return;
        }
    };

    // the serial phase that follows the parallel phase `ph` of context c
    auto serial_phase = [&](int c) {
        PCtx& x = ctx_of(c);
        PoseShared& s = x.ps;
        const ResLayout l = lay_of(c);
        const int ph = x.phase, b = x.pose, n = x.n;
        if (lane == 0) x.wait_tma = 0;
        if ((MODE & MODE_LM) && ph == PH_LM) {
            sum_partials<28>(s.red, s.fin, lane);
            LmState& L = s.lm;
            if (lane == 0) {
                double* trace = a.trace ? a.trace + static_cast<int64_t>(b) * (a.max_iter + 2) * 4 : nullptr;
                lm_advance(L, s.fin, x.kind, x.first != 0, a.max_iter, a.function_tolerance, (a.flags & LC_FLAG_TOL_NEEDS_SUCCESS) != 0, trace);
                x.first = 0;
                x.kind = L.ctl;
            }
            __syncwarp();
            if (L.ctl != CTL_STOP) return;                  // next: another PH_LM
            if (lane == 0) lm_write_result<float>(a, s, b, n, L.term == TERM_CONVERGENCE);
            __syncwarp();
            if (MODE & MODE_LC) need_lc_setup = true;       // the loss runs at the solved pose (s.pose, rounded to fp32 like the reference)
            else need_init = true;
            return;
        }
        if (MODE & MODE_LC) {
            if (ph == PH_LC1) {
                sum_partials<4>(s.red, s.fin, lane);
                if (lane == 0) {
                    bool any_clamped;
                    lc_thresholds1(a, s.fin, n, x.vcnt, x.d0, x.d1, any_clamped);
                    const bool fastK = s.K[6] == 0.0 && s.K[7] == 0.0 && s.K[8] == 1.0;
                    x.tma_out = (fastK && !any_clamped) ? 1 : 0;
                    x.phase = PH_LC2;
                }
            } else if (ph == PH_LC2) {
                sum_partials<2>(s.red, s.fin, lane);
                if (lane == 0) { lc_thresholds2(a, s.fin, x.vcnt, x.sq0, x.sq1); x.phase = PH_LC3; }
            } else if (ph == PH_LC3) {
                sum_partials<48>(s.red, s.fin, lane);
#ifdef LC_TIMING_SIX2
                {   // experiment: the same section twice, cold then warm instruction cache (timing build only)
                    double keep[48];
                    for (int j = lane; j < 48; j += 32) keep[j / 32] = s.fin[j];
                    const long long e0 = clock64();
                    lc_six_fast<float, false>(a, s, b, want_any);
                    __syncwarp();
                    const long long e1 = clock64();
                    for (int j = lane; j < 48; j += 32) s.fin[j] = keep[j / 32];
                    if (lane == 0) s.flag = 0;
                    __syncwarp();
                    const long long e2 = clock64();
                    lc_six_fast<float, false>(a, s, b, want_any);
                    six_t[0] += e1 - e0; six_t[1] += clock64() - e2;
                }
#elif defined(LC_PERSIST_LEAN)
                lc_six_fast<float, false>(a, s, b, want_any);
#else
                lc_six_fast<float>(a, s, b, want_any);
#endif
                __syncwarp();
                if (want_any) { if (lane == 0) x.phase = x.tma_out ? PH_LC4 : PH_LC4G; }
                else need_init = true;
            } else {   // PH_LC4 / PH_LC4G: gradients out, next pose in
                if (ph == PH_LC4 && lane == 0) lc_pass4_store(l, make_vec_io(a, b), n);
                __syncwarp();
                need_init = true;
            }
            __syncwarp();
        }
    };

    unsigned started = 0u;
#ifdef LC_TIMING
    long long ts[16] = {0}, ts_wait = 0, ts_init = 0;
    const long long t_begin = clock64();
#endif
    while (active) {
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            if (!(active & (1u << c))) continue;
            need_init = false; need_lc_setup = false;
            if (started & (1u << c)) {
#ifdef LC_TIMING
                const long long q0 = clock64();
#endif
                bar_sync(2 + c, kPT);                       // the parallel phase of context c is done
#ifdef LC_TIMING
                const long long q1 = clock64();
                ts_wait += q1 - q0;
                const int phq = ctx_of(c).phase;
#endif
                serial_phase(c);
#ifdef LC_TIMING
                ts[phq] += clock64() - q1;
#endif
            } else {
                started |= 1u << c;
                need_init = true;
            }
#ifdef LC_TIMING
            const long long q2 = clock64();
#endif
            if (need_init) init_pose(c);                    // (announces PH_DONE when the pose list is exhausted)
#ifdef LC_TIMING
            ts_init += clock64() - q2;
#endif
            if ((MODE & MODE_LC) && need_lc_setup) {
                PCtx& x = ctx_of(c);
                lc_pose_setup_warp(x.ps, true);
                if (lane == 0) x.phase = PH_LC1;
                __syncwarp();
            }
            __threadfence_block();
            bar_arrive(4 + c, kPT);                         // the serial phase of context c is done
        }
    }
#ifdef LC_TIMING
    if (lane == 0 && a.trace) {
        double* tr = a.trace + static_cast<int64_t>(blockIdx.x) * 64;
        for (int k = 0; k < 8; ++k) tr[16 + k] = static_cast<double>(ts[k]);
        tr[24] = static_cast<double>(ts_wait); tr[25] = static_cast<double>(ts_init); tr[26] = static_cast<double>(clock64() - t_begin);
#ifdef LC_TIMING_SIX2
        tr[27] = static_cast<double>(six_t[0]); tr[28] = static_cast<double>(six_t[1]);
#endif
    }
#endif
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool planar_ok(const lc_view& v, int n) {
    return v.stride[1] == 1 && (n % 4) == 0 && (reinterpret_cast<uintptr_t>(v.ptr) % 16) == 0 && (v.stride[0] % 4) == 0 && (v.stride[2] % 4) == 0;
}

struct DevInfo { int max_smem, sms; };
static bool device_info(DevInfo& d) {
    static std::atomic<int> c_smem[64], c_sms[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return false; }
    const bool cached = dev >= 0 && dev < 64;
    if (cached && (d.max_smem = c_smem[dev].load(std::memory_order_relaxed)) > 0) { d.sms = c_sms[dev].load(std::memory_order_relaxed); return true; }
    if (cudaDeviceGetAttribute(&d.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); return false; }
    if (cached) { c_sms[dev].store(d.sms, std::memory_order_relaxed); c_smem[dev].store(d.max_smem, std::memory_order_relaxed); }
    return true;
}

bool persist_supported(const lc_args& a, int mode) {
    // opt-in (LC_B200_PERSIST=1): measured on B200 at B = 1024 x N = 4096 the CTA-per-pose kernels are faster (profiles/README.md)
    const char* e = getenv("LC_B200_PERSIST");
    if (!e || e[0] != '1') return false;
    if (a.dtype != LC_F32 || a.N < 2048) return false;
    if (a.weight_mode != LC_W_ICOV_DIAG && a.weight_mode != LC_W_INV_STD) return false;
    if (!planar_ok(a.pts3d, a.N) || !planar_ok(a.weights, a.N)) return false;
    // pts2d: planar per pose; a batch-broadcast grid (stride 0) is fine
    if (!(a.pts2d.stride[1] == 1 && (reinterpret_cast<uintptr_t>(a.pts2d.ptr) % 16) == 0 && (a.pts2d.stride[0] % 4) == 0 && (a.pts2d.stride[2] % 4) == 0)) return false;
    if (mode & MODE_LC) {
        if (a.g_pts3d.ptr && !planar_ok(a.g_pts3d, a.N)) return false;
        if (a.g_pts2d.ptr && !planar_ok(a.g_pts2d, a.N)) return false;
        if (a.g_weights.ptr && !planar_ok(a.g_weights, a.N)) return false;
        if (a.valid.ptr && !(a.valid.stride[1] == 1 && (reinterpret_cast<uintptr_t>(a.valid.ptr) % 16) == 0 && (a.valid.stride[0] % 4) == 0)) return false;
    }
    DevInfo d;
    if (!device_info(d)) return false;
    // one pose per SM at a time: worth it once there are at least as many poses as SMs ... and the two buffers must fit
    if (a.B < d.sms) return false;
    return 2 * persist_ctx_bytes(a.N) <= static_cast<size_t>(d.max_smem);
}

template <int MODE>
static int launch_persist_t(const lc_args& a, cudaStream_t st, const DevInfo& d) {
    static std::atomic<bool> configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        const cudaError_t e = cudaFuncSetAttribute(lc_persist_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, d.max_smem);
        if (e != cudaSuccess) return static_cast<int>(e);
        if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    const int grid = a.B < d.sms ? a.B : d.sms;
    lc_persist_kernel<MODE><<<grid, kPT, 2 * persist_ctx_bytes(a.N), st>>>(a, a.N);
    note_kernel("lc::lc_persist_kernel<%s>", MODE == MODE_LM ? "LM" : (MODE == MODE_LC ? "LC" : "LM|LC"));
    return static_cast<int>(cudaGetLastError());
}

int launch_persist_pose(const lc_args& a, int mode, cudaStream_t st) {
    DevInfo d;
    if (!device_info(d)) return static_cast<int>(cudaErrorUnknown);
    switch (mode) {
        case MODE_LM: return launch_persist_t<MODE_LM>(a, st, d);
        case MODE_LC: return launch_persist_t<MODE_LC>(a, st, d);
        default: return launch_persist_t<MODE_LM | MODE_LC>(a, st, d);
    }
}

}  // namespace lc
