// Building blocks of the shared-memory resident kernels (lc_resident.cu, lc_dense.cu): shared-memory layout,
// asynchronous staging helpers (cp.async / TMA bulk + mbarrier), the fp64 LM evaluation pass and the LC phase.
#pragma once

#include <cstdio>
#include <cstdlib>

#include "lc_pose.cuh"

namespace lc {

// Phase timing for tools/phase_timing.py (build with -DLC_TIMING): thread 0 accumulates clock64() deltas per phase and
// the resident kernel writes them to lc_args.trace[b*stride + 0..7].  Compiled out of the product build.
#ifdef LC_TIMING
#define LC_TIC(v) const long long v = clock64()
#define LC_TOC(v, slot) do { if (threadIdx.x == 0) s.fin_timing[slot] += clock64() - v; } while (0)
#else
#define LC_TIC(v) do {} while (0)
#define LC_TOC(v, slot) do {} while (0)
#endif

constexpr int kResidentMinN = 65;

__host__ __device__ inline int round_up4(int n) { return (n + 3) & ~3; }

// planar fp32 arrays behind the PoseShared block (20 B/point)
struct ResLayout {
    float *A0, *A1, *A2;  // X, later q = R X
    float *B0, *B1;       // x, later the clamped error ec
};

__device__ __forceinline__ ResLayout res_layout(unsigned char* base, int npad) {
    float* f = reinterpret_cast<float*>(base + ((sizeof(PoseShared) + 15) & ~size_t(15)));
    ResLayout l;
    l.A0 = f; l.A1 = f + npad; l.A2 = f + 2 * npad; l.B0 = f + 3 * npad; l.B1 = f + 4 * npad;
    return l;
}

// TM kernels: only x -> ec lives in shared memory (8 B/point)
__device__ __forceinline__ ResLayout res_layout_tm(unsigned char* base, int npad) {
    float* f = reinterpret_cast<float*>(base + ((sizeof(PoseShared) + 15) & ~size_t(15)));
    ResLayout l;
    l.A0 = nullptr; l.A1 = nullptr; l.A2 = nullptr; l.B0 = f; l.B1 = f + npad;
    return l;
}
inline size_t resident_smem_bytes_tm(int n) {
    return ((sizeof(PoseShared) + 15) & ~size_t(15)) + sizeof(float) * 2 * static_cast<size_t>(round_up4(n));
}

__host__ __device__ inline size_t resident_smem_bytes(int n) {
    return ((sizeof(PoseShared) + 15) & ~size_t(15)) + sizeof(float) * 5 * static_cast<size_t>(round_up4(n));
}

__device__ __forceinline__ float ldf(const lc_view& v, int64_t off) { return static_cast<const float*>(v.ptr)[off]; }
__device__ __forceinline__ void stf(const lc_view& v, int64_t off, float x) { static_cast<float*>(v.ptr)[off] = x; }
__device__ __forceinline__ float nan_to_num_f(float x) {
    if (isnan(x)) return 0.f;
    if (isinf(x)) return x > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
    return x;
}

// 4-byte asynchronous global -> shared copy (LDGSTS): no register staging, every copy of a pose is in flight at once
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---- TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: one instruction per contiguous slab ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// Ask the L2 to fetch a contiguous slab (cp.async.bulk.prefetch.L2): the weights are not staged in shared memory, this
// makes their first pass an L2 hit instead of an HBM round trip.  bytes must be a multiple of 16, the address 16-byte aligned.
__device__ __forceinline__ void l2_prefetch_bulk(const void* gsrc, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    }
}

// ---- tensor memory (TMEM) as per-thread scratch ----
// The 256 KB of tensor memory per SM are otherwise unused by this path.  Every thread only ever touches its own points
// (i = tid + k NT), so the per-point 3-vector (X, later q = R X) can live in the thread's own TMEM lane, 4 columns per point
// (tcgen05.st / tcgen05.ld 32x32b.x4), while only x -> ec (8 B/point) stays in shared memory: 47 KB per CTA at N = 4096, i.e.
// four 128-thread CTAs per SM where shared memory alone allows two 256-thread ones.  tools/tmem_probe.cu checks the idiom.
__device__ __forceinline__ void tmem_st4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)),
                 "r"(__float_as_uint(c)), "r"(__float_as_uint(d)) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t addr, float& a, float& b, float& c) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    a = __uint_as_float(r0); b = __uint_as_float(r1); c = __uint_as_float(r2);
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
constexpr int kTmemCols = 128;           // columns per CTA: 32 points per thread x 4
constexpr int kTmemMaxN = 128 * 32;      // with 128 threads

// Where the per-point 3-vector lives.  TM = false: planar shared-memory arrays (any N that fits).  TM = true: the thread's
// TMEM lane; loads / stores are warp-collective (.sync.aligned), so the point loops of TM kernels run to a warp-uniform
// bound with a `live` predicate (see LC_POINT_LOOP).
template <bool TM>
struct XAcc {
    ResLayout l;
    uint32_t tb;   // TMEM address of this warp's lane quarter (TM only)
    __device__ __forceinline__ void ld(int i, int k, bool live, float& a, float& b, float& c) const {
        if (TM) tmem_ld4(tb + 4u * k, a, b, c);
        else if (live) { a = l.A0[i]; b = l.A1[i]; c = l.A2[i]; }
    }
    __device__ __forceinline__ void st(int i, int k, bool live, float a, float b, float c) const {
        if (TM) tmem_st4(tb + 4u * k, a, b, c, 0.f);
        else if (live) { l.A0[i] = a; l.A1[i] = b; l.A2[i] = c; }
    }
};
// for (point i of this thread, k-th of them): TM kernels iterate to a warp-uniform bound, `live` masks the tail
#define LC_POINT_LOOP(TM, NT, n)                                                                                   \
    for (int i = threadIdx.x, k = 0; ((TM) ? (k * (NT) + static_cast<int>(threadIdx.x & ~31u)) : i) < (n); i += (NT), ++k)

// ---- one pose split over a thread-block cluster (lc_resident_kernel.cuh, CL > 1) ----
// The CTAs of a cluster each hold a contiguous share of the pose's points and run the SAME code on it; after every CTA-wide
// reduction the per-CTA totals are exchanged through distributed shared memory and summed in rank order, so every CTA of the
// cluster continues with bit-identical totals and takes identical decisions in the (redundantly executed) serial sections.
// One cluster barrier per reduction: the exchange buffer alternates, a CTA cannot be two reductions ahead of a peer.
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_dsmem_f64(const double* local_smem, unsigned rank) {
    unsigned remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_smem)), "r"(rank));
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote) : "memory");
    return v;
}
struct ClusterShared {
    lc_args args;        // this CTA's view of the call: point-indexed pointers advanced to its share, N = its share
    double xch[2][48];   // reduction exchange, double-buffered
};
template <int CL>
struct Clu {
    double (*xch)[48];
    int parity, n_total;   // n_total: points of the whole pose (statistics that divide by the count)
    bool leader;           // rank 0 writes the per-pose outputs
    template <int V>
    __device__ __forceinline__ void combine(double* fin) {   // after block_reduce: fin[0..V) := sum over the cluster's CTAs
        const int tid = threadIdx.x;
        double* mine = xch[parity];
        if (tid < V) mine[tid] = fin[tid];
        cluster_sync_all();
        if (tid < V) {
            double t = 0.0;
#pragma unroll
            for (unsigned r = 0; r < CL; ++r) t += ld_dsmem_f64(mine + tid, r);
            fin[tid] = t;
        }
        __syncthreads();
        parity ^= 1;
    }
};
template <>
struct Clu<1> {
    int n_total;
    static constexpr bool leader = true;
    template <int V>
    __device__ __forceinline__ void combine(double*) {}
};

// One evaluation pass of the reprojection cost (ceres.cpp:30-55) at the point held in L.Rm/L.te, from the
// staged fp32 arrays: cost, and when JAC also J'^T J' and J'^T r in the left basis.
template <int NT, bool JAC, bool TM, class CLU>
__device__ __forceinline__ void lm_eval_pass_res(const lc_args& a, PoseShared& s, const ResLayout& l, int b, int n, bool sanitize, const XAcc<TM>& xs, CLU& cl) {
    const LmState& L = s.lm;
    double acc[28];
#pragma unroll
    for (int k = 0; k < 28; ++k) acc[k] = 0.0;
    const double k00 = s.K[0], k01 = s.K[1], k10 = s.K[3], k11 = s.K[4], cx = s.K[2], cy = s.K[5];
    const double R0 = L.Rm[0], R1 = L.Rm[1], R2 = L.Rm[2], R3 = L.Rm[3], R4 = L.Rm[4], R5 = L.Rm[5], R6 = L.Rm[6], R7 = L.Rm[7], R8 = L.Rm[8];
    const double t0 = L.te[0], t1 = L.te[1], t2 = L.te[2];
    // the sqrt-information weights are re-read from L2 every pass (8 B/point) instead of living in shared memory:
    // that is what lets two CTAs share an SM at N = 4096.  The next point's weights are fetched before the
    // current point is processed.
    const float* pw = static_cast<const float*>(a.weights.ptr) + b * a.weights.stride[0];
    const int64_t swn = a.weights.stride[1], swc = a.weights.stride[2];
    const bool icov = a.weight_mode == LC_W_ICOV_DIAG;
    float w0 = 0.f, w1 = 0.f;
    if (static_cast<int>(threadIdx.x) < n) { w0 = pw[threadIdx.x * swn]; w1 = pw[threadIdx.x * swn + swc]; }
    LC_POINT_LOOP(TM, NT, n) {
        const bool live = !TM || i < n;
        float Xf0 = 0.f, Xf1 = 0.f, Xf2 = 0.f;
        xs.ld(i, k, live, Xf0, Xf1, Xf2);
        if (!live) continue;
        float wa = w0, wb = w1;
        const int inext = i + NT;
        if (inext < n) { w0 = pw[inext * swn]; w1 = pw[inext * swn + swc]; }
        if (sanitize) { wa = nan_to_num_f(wa); wb = nan_to_num_f(wb); }
        // cer_solver.py:37-38: L = diag(sqrt(icov)) in fp32.  For LC_W_INV_STD the reference's sqrt(fl(s*s)) is exactly
        // |s| (barring overflow / underflow of s*s).
        if (icov) { wa = sqrtf(wa); wb = sqrtf(wb); }
        const double la = fabsf(wa), lc_ = fabsf(wb);
        const double X0 = Xf0, X1 = Xf1, X2 = Xf2;
        const double px = l.B0[i], py = l.B1[i];
        const double q0 = fma(R0, X0, fma(R1, X1, R2 * X2));
        const double q1 = fma(R3, X0, fma(R4, X1, R5 * X2));
        const double q2 = fma(R6, X0, fma(R7, X1, R8 * X2));
        const double p0 = q0 + t0, p1 = q1 + t1, p2 = q2 + t2;
        const double iz = fast_rcp(p2);
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
    cl.template combine<28>(s.fin);
}

// ---------------------------------------------------------------------------------------------
// LC phase on the staged arrays (A = X -> q, B = x -> ec), shared by the resident kernel and the dense-producer kernel
// (lc_dense.cu).  WSrc supplies the inverse-std weights of point i, Sink receives the per-point gradients.
// ---------------------------------------------------------------------------------------------
struct DirectWeights {   // weights streamed from a strided (N,2) view (L2 resident after the first touch)
    const float* p;
    int64_t sn, sc;
    __device__ __forceinline__ void get(int i, float& s0, float& s1) const { s0 = p[i * sn]; s1 = p[i * sn + sc]; }
};
struct DirectSink {      // gradients written to the strided views of lc_args
    const lc_args& a;
    int b;
    __device__ __forceinline__ bool want_any() const { return a.g_pts3d.ptr || a.g_pts2d.ptr || a.g_weights.ptr; }
    __device__ __forceinline__ bool want_pts3d() const { return a.g_pts3d.ptr != nullptr; }
    __device__ __forceinline__ void weight_grad(int i, int c, float g, float) const {
        if (a.g_weights.ptr) stf(a.g_weights, b * a.g_weights.stride[0] + i * a.g_weights.stride[1] + c * a.g_weights.stride[2], g);
    }
    __device__ __forceinline__ void pts2d_grad(int i, int c, float g) const {
        if (a.g_pts2d.ptr) stf(a.g_pts2d, b * a.g_pts2d.stride[0] + i * a.g_pts2d.stride[1] + c * a.g_pts2d.stride[2], g);
    }
    __device__ __forceinline__ void pts3d_grad(int i, float g0, float g1, float g2) const {
        const int64_t o = b * a.g_pts3d.stride[0] + i * a.g_pts3d.stride[1];
        stf(a.g_pts3d, o, g0); stf(a.g_pts3d, o + a.g_pts3d.stride[2], g1); stf(a.g_pts3d, o + 2 * a.g_pts3d.stride[2], g2);
    }
};

// Per-point fp32 geometry shared by passes 3 and 4: left-basis Jacobian rows (depth-decoupled accumulation basis, see
// lc_phase_res) and the robust weights of robust_weights_cov (cov_mixed.py:27-39).
struct PointConsts {
    float t0, t1, t2, k00, k01, k10, k11, uc, vc, d0, d1, sq0, sq1;
};
__device__ __forceinline__ PointConsts make_point_consts(const PoseShared& s, float d0, float d1, float sq0, float sq1) {
    PointConsts c;
    c.t0 = static_cast<float>(s.t[0]); c.t1 = static_cast<float>(s.t[1]); c.t2 = static_cast<float>(s.t[2]);
    c.k00 = static_cast<float>(s.K[0]); c.k01 = static_cast<float>(s.K[1]); c.k10 = static_cast<float>(s.K[3]); c.k11 = static_cast<float>(s.K[4]);
    c.uc = static_cast<float>(s.t[0] / s.t[2]); c.vc = static_cast<float>(s.t[1] / s.t[2]);
    c.d0 = d0; c.d1 = d1; c.sq0 = sq0; c.sq1 = sq1;
    return c;
}
__device__ __forceinline__ void point_terms_f(const PointConsts& pc, float q0, float q1, float q2, const float (&ec)[2], const float (&sk)[2],
                                              float (&J)[2][6], float (&sg)[2], float (&del)[2], float (&w)[2]) {
    const float P0 = q0 + pc.t0, P1 = q1 + pc.t1, P2 = q2 + pc.t2;
    const float iz = __fdividef(1.f, P2);
    const float u0 = P0 * iz, v0 = P1 * iz;
    const float du0 = fmaf(pc.uc, q2, -q0) * iz, dv0 = fmaf(pc.vc, q2, -q1) * iz;   // uvc - uv0
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const float ka = c ? pc.k10 : pc.k00, kb = c ? pc.k11 : pc.k01;
        const float e0 = ka * iz, e1 = kb * iz, e2 = -fmaf(ka, u0, kb * v0) * iz;
        J[c][0] = fmaf(q1, e2, -q2 * e1);
        J[c][1] = fmaf(q2, e0, -q0 * e2);
        J[c][2] = fmaf(q0, e1, -q1 * e0);
        J[c][3] = e0; J[c][4] = e1; J[c][5] = fmaf(ka, du0, kb * dv0) * iz;
        const float dc = c ? pc.d1 : pc.d0, sq = c ? pc.sq1 : pc.sq0;
        const float av = fabsf(ec[c]);
        sg[c] = av > dc ? dc * (2.f * av - dc) : av * av;
        del[c] = sq * rsqrtf(sg[c] + 1e-6f);
        w[c] = sk[c] > del[c] ? del[c] * (2.f * sk[c] - del[c]) : sk[c] * sk[c];
    }
}

// pass 4 (fp32), one point per thread and iteration, any strides: per-coordinate adjoints and the three input
// gradients.  Also the fallback of the vectorised phase (lc_vec.cuh) for poses with a general K row 2 or a clamped depth.
template <int NT, class WSrc, class Sink, bool TM>
__device__ __forceinline__ void lc_pass4_scalar(const lc_args& a, PoseShared& s, const ResLayout& l, int n, float d0, float d1, float sq0, float sq1,
                                                const WSrc& wsrc, Sink& sink, const XAcc<TM>& xs) {
    const PointConsts pc = make_point_consts(s, d0, d1, sq0, sq1);
    const float t0 = pc.t0, t1 = pc.t1, t2 = pc.t2;
    auto point_terms = [&](int i, float q0, float q1, float q2, float (&J)[2][6], float (&ec)[2], float (&sk)[2], float (&sg)[2], float (&del)[2],
                           float (&w)[2]) {
        ec[0] = l.B0[i]; ec[1] = l.B1[i];
        wsrc.get(i, sk[0], sk[1]);
        point_terms_f(pc, q0, q1, q2, ec, sk, J, sg, del, w);
    };
    // gradient slots of the padding beyond n (ragged batches): defined, zero
    for (int i = n + static_cast<int>(threadIdx.x); i < a.N; i += NT) {
        for (int c = 0; c < 2; ++c) { sink.weight_grad(i, c, 0.f, 0.f); sink.pts2d_grad(i, c, 0.f); }
        if (sink.want_pts3d()) sink.pts3d_grad(i, 0.f, 0.f, 0.f);
    }
    {
        float cH[kSym], cG[kSym], bL[6];
#pragma unroll
        for (int k = 0; k < kSym; ++k) { cH[k] = static_cast<float>(s.cHL[k]); cG[k] = static_cast<float>(s.cGL[k]); }
#pragma unroll
        for (int k = 0; k < 6; ++k) bL[k] = static_cast<float>(s.bL[k]);
        const float K0 = static_cast<float>(s.K[0]), K1 = static_cast<float>(s.K[1]), K2 = static_cast<float>(s.K[2]);
        const float K3 = static_cast<float>(s.K[3]), K4 = static_cast<float>(s.K[4]), K5 = static_cast<float>(s.K[5]);
        const float K6 = static_cast<float>(s.K[6]), K7 = static_cast<float>(s.K[7]), K8 = static_cast<float>(s.K[8]);
        float Rf[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) Rf[k] = static_cast<float>(s.R[k]);
        LC_POINT_LOOP(TM, NT, n) {
            const bool live = !TM || i < n;
            float q0 = 0.f, q1 = 0.f, q2 = 0.f;
            xs.ld(i, k, live, q0, q1, q2);
            if (!live) continue;
            float J[2][6], ec[2], sk[2], sg[2], del[2], w[2];
            point_terms(i, q0, q1, q2, J, ec, sk, sg, del, w);
            float ecb[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float qh = 0.f, qg = 0.f, lb = 0.f;
                int k = 0;
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    lb = fmaf(J[c][r], bL[r], lb);
#pragma unroll
                    for (int cc = r; cc < 6; ++cc) {
                        const float pp = J[c][r] * J[c][cc];
                        qh = fmaf(cH[k], pp, qh);
                        qg = fmaf(cG[k], pp, qg);
                        ++k;
                    }
                }
                const float Wbar = qh + 2.f * w[c] * sg[c] * qg + ec[c] * lb;
                const float sigbar = w[c] * w[c] * qg;
                sink.weight_grad(i, c, Wbar * (sk[c] > del[c] ? 2.f * del[c] : 2.f * sk[c]), sk[c]);
                const float dc = c ? d1 : d0;
                const float av = fabsf(ec[c]);
                const float sgn = (ec[c] > 0.f) ? 1.f : ((ec[c] < 0.f) ? -1.f : 0.f);
                ecb[c] = sigbar * (av > dc ? 2.f * dc : 2.f * av) * sgn;
                sink.pts2d_grad(i, c, ecb[c]);
            }
            if (sink.want_pts3d()) {
                // gX = -R^T (dproj/dP)^T ecbar,  dproj/dP = (K[:2,:] - proj (x) K[2,:] [z >= 0.1]) / max(z, 0.1)
                const float P0 = q0 + t0, P1 = q1 + t1, P2 = q2 + t2;
                const float KP0 = fmaf(K0, P0, fmaf(K1, P1, K2 * P2));
                const float KP1 = fmaf(K3, P0, fmaf(K4, P1, K5 * P2));
                const float KP2 = fmaf(K6, P0, fmaf(K7, P1, K8 * P2));
                const bool act = KP2 >= 0.1f;
                const float iz = __fdividef(1.f, act ? KP2 : 0.1f);
                const float pr0 = act ? KP0 * iz : 0.f, pr1 = act ? KP1 * iz : 0.f;   // proj * [z >= 0.1]
                const float gP0 = (fmaf(-pr0, K6, K0) * ecb[0] + fmaf(-pr1, K6, K3) * ecb[1]) * iz;
                const float gP1 = (fmaf(-pr0, K7, K1) * ecb[0] + fmaf(-pr1, K7, K4) * ecb[1]) * iz;
                const float gP2 = (fmaf(-pr0, K8, K2) * ecb[0] + fmaf(-pr1, K8, K5) * ecb[1]) * iz;
                sink.pts3d_grad(i, -(Rf[0] * gP0 + Rf[3] * gP1 + Rf[6] * gP2), -(Rf[1] * gP0 + Rf[4] * gP1 + Rf[7] * gP2),
                                -(Rf[2] * gP0 + Rf[5] * gP1 + Rf[8] * gP2));
            }
        }
    }
}

template <int NT, class WSrc, class Sink, bool TM>
__device__ __forceinline__ void lc_phase_res(const lc_args& a, PoseShared& s, const ResLayout& l, int b, int n, const WSrc& wsrc, Sink& sink,
                                             const XAcc<TM>& xs) {
    const int tid = threadIdx.x;
    { LC_TIC(tq1);
    if (tid < 32) lc_pose_setup_warp(s, true);
    __syncthreads();
    LC_TOC(tq1, 3); }
    LC_TIC(tq2);

    const int64_t ovb = a.valid.ptr ? b * a.valid.stride[0] : 0;
    // pass 1 (fp64): P = R X + t, project_apply + clamp_error once per point; P, ec kept as fp32 in place
    {
        const double Lmax = a.max_err_len;
        const double lim = Lmax - 1e-6, lim2 = lim > 0.0 ? lim * lim : -1.0;   // |e|+1e-6 > Lmax  <=>  |e|^2 > lim2
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
        LC_POINT_LOOP(TM, NT, n) {
            const bool live = !TM || i < n;
            float Xf0 = 0.f, Xf1 = 0.f, Xf2 = 0.f;
            xs.ld(i, k, live, Xf0, Xf1, Xf2);
            float qf0 = 0.f, qf1 = 0.f, qf2 = 0.f;
            if (live) {
            const double X0 = Xf0, X1 = Xf1, X2 = Xf2, x0 = l.B0[i], x1 = l.B1[i];
            const double q0 = fma(s.R[0], X0, fma(s.R[1], X1, s.R[2] * X2));
            const double q1 = fma(s.R[3], X0, fma(s.R[4], X1, s.R[5] * X2));
            const double q2 = fma(s.R[6], X0, fma(s.R[7], X1, s.R[8] * X2));
            const double P0 = q0 + s.t[0], P1 = q1 + s.t[1], P2 = q2 + s.t[2];
            const double KP0 = fma(s.K[0], P0, fma(s.K[1], P1, s.K[2] * P2));
            const double KP1 = fma(s.K[3], P0, fma(s.K[4], P1, s.K[5] * P2));
            const double KP2 = fma(s.K[6], P0, fma(s.K[7], P1, s.K[8] * P2));
            const double iz = fast_rcp(KP2 > 0.1 ? KP2 : 0.1);
            double e0 = fma(-KP0, iz, x0), e1 = fma(-KP1, iz, x1);
            const double l2 = fma(e0, e0, e1 * e1);
            if (l2 > lim2) {
                const double len = sqrt(l2) + 1e-6;
                const double f = (len - Lmax) / len;
                e0 = fma(-f, e0, e0);
                e1 = fma(-f, e1, e1);
            }
            const float ec0 = static_cast<float>(e0), ec1 = static_cast<float>(e1);
            // q = R X is what is cached (not P = q + t): q x D then keeps fp32 relative precision even when |X| << |t|
            qf0 = static_cast<float>(q0); qf1 = static_cast<float>(q1); qf2 = static_cast<float>(q2);
            l.B0[i] = ec0; l.B1[i] = ec1;
            const float v = a.valid.ptr ? ldf(a.valid, ovb + i * a.valid.stride[1]) : 1.f;
            acc0 = fmaf(v, fabsf(ec0), acc0); acc1 = fmaf(v, fabsf(ec1), acc1); acc2 += v;
            }
            xs.st(i, k, live, qf0, qf1, qf2);
        }
        if (TM) tmem_wait_st();
        double acc[3] = {acc0, acc1, acc2};
        block_reduce<3, NT>(acc, s.red, s.fin);
    }
    const double vcnt = a.valid.ptr ? s.fin[2] : static_cast<double>(n);
    const float d0 = static_cast<float>(a.rel_thresh * (s.fin[0] / vcnt)), d1 = static_cast<float>(a.rel_thresh * (s.fin[1] / vcnt));
    __syncthreads();
    // pass 2 (fp32): q_a = mean valid s^2 sigma
    {
        float acc0 = 0.f, acc1 = 0.f;
        for (int i = tid; i < n; i += NT) {
            float s0, s1;
            wsrc.get(i, s0, s1);
            const float v = a.valid.ptr ? ldf(a.valid, ovb + i * a.valid.stride[1]) : 1.f;
            const float a0 = fabsf(l.B0[i]), a1 = fabsf(l.B1[i]);
            const float sg0 = a0 > d0 ? d0 * (2.f * a0 - d0) : a0 * a0;
            const float sg1 = a1 > d1 ? d1 * (2.f * a1 - d1) : a1 * a1;
            acc0 = fmaf(v * (s0 * s0), sg0, acc0);
            acc1 = fmaf(v * (s1 * s1), sg1, acc1);
        }
        double acc[2] = {acc0, acc1};
        block_reduce<2, NT>(acc, s.red, s.fin);
    }
    // delta_k = sqrt(we * q_a / (sigma_k + 1e-6)) = sq_a * rsqrt(sigma_k + 1e-6)
    const float sq0 = static_cast<float>(sqrt((s.fin[0] / vcnt) * a.w_e_thresh)), sq1 = static_cast<float>(sqrt((s.fin[1] / vcnt) * a.w_e_thresh));
    __syncthreads();

    const float t0 = static_cast<float>(s.t[0]), t1 = static_cast<float>(s.t[1]), t2 = static_cast<float>(s.t[2]);
    const float k00 = static_cast<float>(s.K[0]), k01 = static_cast<float>(s.K[1]), k10 = static_cast<float>(s.K[3]), k11 = static_cast<float>(s.K[4]);
    // Depth-decoupled accumulation basis.  For an object that is small compared to its depth every point has nearly
    // the same normalised image position uv0, so the t_z Jacobian column D_z = -K uv0 / z is nearly a fixed combination
    // of the t_x, t_y columns and H is ill conditioned (cond ~ (z/size)^2: the depth ambiguity).  The sums are therefore
    // taken with the column t_z' = t_z + uc t_x + vc t_y, (uc, vc) = t_xy / t_z, whose entries
    //   D_z + uc D_x + vc D_y = K (uvc - uv0) / z = K (uvc q_z - q_xy) / z^2
    // are formed from the small vector q directly (no cancellation), and mapped back in fp64 (PoseShared::Tm).
    const float uc = static_cast<float>(s.t[0] / s.t[2]), vc = static_cast<float>(s.t[1] / s.t[2]);

    // per-point fp32 geometry shared by passes 3 and 4: left-basis Jacobian rows and robust weights
    auto point_terms = [&](int i, float q0, float q1, float q2, float (&J)[2][6], float (&ec)[2], float (&sk)[2], float (&sg)[2], float (&del)[2],
                           float (&w)[2]) {
        const float P0 = q0 + t0, P1 = q1 + t1, P2 = q2 + t2;
        ec[0] = l.B0[i]; ec[1] = l.B1[i];
        wsrc.get(i, sk[0], sk[1]);
        const float iz = __fdividef(1.f, P2);
        const float u0 = P0 * iz, v0 = P1 * iz;
        const float du0 = fmaf(uc, q2, -q0) * iz, dv0 = fmaf(vc, q2, -q1) * iz;   // uvc - uv0
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float ka = c ? k10 : k00, kb = c ? k11 : k01;
            const float e0 = ka * iz, e1 = kb * iz, e2 = -fmaf(ka, u0, kb * v0) * iz;
            J[c][0] = fmaf(q1, e2, -q2 * e1);
            J[c][1] = fmaf(q2, e0, -q0 * e2);
            J[c][2] = fmaf(q0, e1, -q1 * e0);
            J[c][3] = e0; J[c][4] = e1; J[c][5] = fmaf(ka, du0, kb * dv0) * iz;
            const float dc = c ? d1 : d0, sq = c ? sq1 : sq0;
            const float av = fabsf(ec[c]);
            sg[c] = av > dc ? dc * (2.f * av - dc) : av * av;
            del[c] = sq * rsqrtf(sg[c] + 1e-6f);
            w[c] = sk[c] > del[c] ? del[c] * (2.f * sk[c] - del[c]) : sk[c] * sk[c];
        }
    };

    // pass 3 (fp32 partial sums, fp64 CTA reduction): H' = sum W J'J'^T, G' = sum W^2 sigma J'J'^T, b' = sum W ec J'
    {
        float acc[48];
#pragma unroll
        for (int k = 0; k < 48; ++k) acc[k] = 0.f;
        LC_POINT_LOOP(TM, NT, n) {
            const bool live = !TM || i < n;
            float q0 = 0.f, q1 = 0.f, q2 = 0.f;
            xs.ld(i, k, live, q0, q1, q2);
            if (!live) continue;
            float J[2][6], ec[2], sk[2], sg[2], del[2], w[2];
            point_terms(i, q0, q1, q2, J, ec, sk, sg, del, w);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                acc_outer<0>(acc, w[c], J[c]);
                acc_outer<21>(acc, w[c] * w[c] * sg[c], J[c]);
                const float wb = w[c] * ec[c];
#pragma unroll
                for (int r = 0; r < 6; ++r) acc[42 + r] = fmaf(wb, J[c][r], acc[42 + r]);
            }
        }
        double accd[48];
#pragma unroll
        for (int k = 0; k < 48; ++k) accd[k] = acc[k];
        block_reduce<48, NT>(accd, s.red, s.fin);
    }
    LC_TOC(tq2, 4);
    { LC_TIC(tq3);
    if (tid < 32) lc_six_fast<float>(a, s, b, sink.want_any());
    __syncthreads();
    if (!sink.want_any()) return;
    LC_TOC(tq3, 5); }
    LC_TIC(tq4);

    lc_pass4_scalar<NT>(a, s, l, n, d0, d1, sq0, sq1, wsrc, sink, xs);
#ifdef LC_TIMING
    __syncthreads();
#endif
    LC_TOC(tq4, 6);
}

}  // namespace lc
