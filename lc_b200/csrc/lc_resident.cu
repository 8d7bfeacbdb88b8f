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
#include <cstdio>
#include <cstdlib>

#include "lc_pose.cuh"

namespace lc {

constexpr int kResidentMinN = 65;

__host__ __device__ inline int round_up4(int n) { return (n + 3) & ~3; }

struct ResLayout {
    float *A0, *A1, *A2;  // X -> P
    float *B0, *B1;       // x -> ec
    float *S0, *S1;       // weights (raw staging only)
};

__device__ __forceinline__ ResLayout res_layout(unsigned char* base, int npad, bool raw) {
    float* f = reinterpret_cast<float*>(base + ((sizeof(PoseShared) + 15) & ~size_t(15)));
    ResLayout l;
    l.A0 = f; l.A1 = f + npad; l.A2 = f + 2 * npad; l.B0 = f + 3 * npad; l.B1 = f + 4 * npad;
    l.S0 = raw ? f + 5 * npad : nullptr;
    l.S1 = raw ? f + 6 * npad : nullptr;
    return l;
}

static size_t resident_smem_bytes(int n, bool raw) {
    return ((sizeof(PoseShared) + 15) & ~size_t(15)) + sizeof(float) * (raw ? 7 : 5) * static_cast<size_t>(round_up4(n));
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
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    }
}

// One evaluation pass of the reprojection cost (ceres.cpp:30-55) at the point held in L.Rm/L.te, from the
// staged fp32 arrays: cost, and when JAC also J'^T J' and J'^T r in the left basis.
template <int NT, bool JAC>
__device__ __forceinline__ void lm_eval_pass_res(const lc_args& a, PoseShared& s, const ResLayout& l, int b, int n, bool sanitize) {
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
    int i = threadIdx.x;
    float w0 = 0.f, w1 = 0.f;
    if (i < n) { w0 = pw[i * swn]; w1 = pw[i * swn + swc]; }
    for (; i < n; i += NT) {
        float wa = w0, wb = w1;
        const int inext = i + NT;
        if (inext < n) { w0 = pw[inext * swn]; w1 = pw[inext * swn + swc]; }
        if (sanitize) { wa = nan_to_num_f(wa); wb = nan_to_num_f(wb); }
        // cer_solver.py:37-38: L = diag(sqrt(icov)) in fp32.  For LC_W_INV_STD the reference's sqrt(fl(s*s)) is exactly
        // |s| (barring overflow / underflow of s*s).
        if (icov) { wa = sqrtf(wa); wb = sqrtf(wb); }
        const double la = fabsf(wa), lc_ = fabsf(wb);
        const double X0 = l.A0[i], X1 = l.A1[i], X2 = l.A2[i];
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
        if (!JAC) continue;
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
    acc[27] *= 0.5;
    block_reduce<28, NT>(acc, s.red, s.fin);
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

template <int NT, class WSrc, class Sink>
__device__ __forceinline__ void lc_phase_res(const lc_args& a, PoseShared& s, const ResLayout& l, int b, int n, const WSrc& wsrc, Sink& sink) {
    const int tid = threadIdx.x;
    if (tid == 0) lc_pose_setup(s, true);
    __syncthreads();

    const int64_t ovb = a.valid.ptr ? b * a.valid.stride[0] : 0;
    // pass 1 (fp64): P = R X + t, project_apply + clamp_error once per point; P, ec kept as fp32 in place
    {
        const double Lmax = a.max_err_len;
        const double lim = Lmax - 1e-6, lim2 = lim > 0.0 ? lim * lim : -1.0;   // |e|+1e-6 > Lmax  <=>  |e|^2 > lim2
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
        for (int i = tid; i < n; i += NT) {
            const double X0 = l.A0[i], X1 = l.A1[i], X2 = l.A2[i], x0 = l.B0[i], x1 = l.B1[i];
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
            l.A0[i] = static_cast<float>(q0); l.A1[i] = static_cast<float>(q1); l.A2[i] = static_cast<float>(q2);
            l.B0[i] = ec0; l.B1[i] = ec1;
            const float v = a.valid.ptr ? ldf(a.valid, ovb + i * a.valid.stride[1]) : 1.f;
            acc0 = fmaf(v, fabsf(ec0), acc0); acc1 = fmaf(v, fabsf(ec1), acc1); acc2 += v;
        }
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
    auto point_terms = [&](int i, float (&J)[2][6], float (&ec)[2], float (&sk)[2], float (&sg)[2], float (&del)[2], float (&w)[2]) {
        const float q0 = l.A0[i], q1 = l.A1[i], q2 = l.A2[i];
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
        for (int i = tid; i < n; i += NT) {
            float J[2][6], ec[2], sk[2], sg[2], del[2], w[2];
            point_terms(i, J, ec, sk, sg, del, w);
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
    lc_six_forward<float, NT>(a, s, b);
    if (!sink.want_any()) return;
    lc_six_backward<NT>(s);

    // pass 4 (fp32): per-coordinate adjoints and the three input gradients
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
        for (int i = tid; i < n; i += NT) {
            float J[2][6], ec[2], sk[2], sg[2], del[2], w[2];
            point_terms(i, J, ec, sk, sg, del, w);
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
                const float P0 = l.A0[i] + t0, P1 = l.A1[i] + t1, P2 = l.A2[i] + t2;
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

template <int NT, int MODE>
__global__ void __launch_bounds__(NT, 2) lc_resident_kernel(const lc_args a, int npad, int tma_mask) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool RAW = false;   // weights are never staged: they are streamed from L2 (see lm_eval_pass_res)
    PoseShared& s = *reinterpret_cast<PoseShared*>(smem_raw);
    const ResLayout l = res_layout(smem_raw, npad, RAW);
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int n = a.n_points ? min(max(a.n_points[b], 0), a.N) : a.N;
    const bool sanitize = (MODE & MODE_LM) && (a.flags & LC_FLAG_NAN_TO_NUM);

    // ---- stage the correspondences.  Planar, 16-byte aligned arrays (what the dense call site produces, tma_mask
    //      bit 0 = pts3d, bit 1 = pts2d) go through the TMA: one 1-D bulk copy per component slab, issued by one
    //      thread, completing on an mbarrier.  Anything else: one 4-byte cp.async per element (any strides).  Either
    //      way every byte of the pose is in flight while the pose constants are set up. ----
    {
        const float* p3 = static_cast<const float*>(a.pts3d.ptr) + b * a.pts3d.stride[0];
        const float* p2 = static_cast<const float*>(a.pts2d.ptr) + b * a.pts2d.stride[0];
        const int64_t s3n = a.pts3d.stride[1], s3c = a.pts3d.stride[2], s2n = a.pts2d.stride[1], s2c = a.pts2d.stride[2];
        const bool tma3 = (tma_mask & 1) != 0, tma2 = (tma_mask & 2) != 0;
        if (tma_mask) {
            if (tid == 0) mbar_init(&s.tma_bar, 1);
            __syncthreads();
            if (tid == 0) {
                const unsigned slab = static_cast<unsigned>(a.N) * 4u;
                mbar_expect_tx(&s.tma_bar, slab * ((tma3 ? 3u : 0u) + (tma2 ? 2u : 0u)));
                if (tma3) { tma_load_1d(l.A0, p3, slab, &s.tma_bar); tma_load_1d(l.A1, p3 + s3c, slab, &s.tma_bar); tma_load_1d(l.A2, p3 + 2 * s3c, slab, &s.tma_bar); }
                if (tma2) { tma_load_1d(l.B0, p2, slab, &s.tma_bar); tma_load_1d(l.B1, p2 + s2c, slab, &s.tma_bar); }
            }
        }
        if (!(tma3 && tma2)) {
            for (int i = tid; i < npad; i += NT) {
                if (i < n) {
                    if (!tma3) { cp_async4(l.A0 + i, p3 + i * s3n); cp_async4(l.A1 + i, p3 + i * s3n + s3c); cp_async4(l.A2 + i, p3 + i * s3n + 2 * s3c); }
                    if (!tma2) { cp_async4(l.B0 + i, p2 + i * s2n); cp_async4(l.B1 + i, p2 + i * s2n + s2c); }
                } else {
                    if (!tma3) { l.A0[i] = 0.f; l.A1[i] = 0.f; l.A2[i] = 0.f; }
                    if (!tma2) { l.B0[i] = 0.f; l.B1[i] = 0.f; }
                }
            }
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
            l.A0[i] = nan_to_num_f(l.A0[i]); l.A1[i] = nan_to_num_f(l.A1[i]); l.A2[i] = nan_to_num_f(l.A2[i]);
            l.B0[i] = nan_to_num_f(l.B0[i]); l.B1[i] = nan_to_num_f(l.B1[i]);
        }
    }
    __syncthreads();

    // =========================== LM solve (fp64) ===========================
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
                if (kind == CTL_EVAL_COST) lm_eval_pass_res<NT, false>(a, s, l, b, n, sanitize);
                else lm_eval_pass_res<NT, true>(a, s, l, b, n, sanitize);
                if (tid == 0)
                    lm_advance(L, s.fin, kind, first, a.max_iter, a.function_tolerance, (a.flags & LC_FLAG_TOL_NEEDS_SUCCESS) != 0, trace);
                first = false;
                __syncthreads();
                if (L.ctl == CTL_STOP) break;
            }
            solved = L.term == TERM_CONVERGENCE;
        }
        if (tid == 0) lm_write_result<float>(a, s, b, n, solved);
        __syncthreads();
    }
    if (!(MODE & MODE_LC)) return;

    // =========================== LC loss ===========================
    const DirectWeights wsrc{static_cast<const float*>(a.weights.ptr) + b * a.weights.stride[0], a.weights.stride[1], a.weights.stride[2]};
    DirectSink sink{a, b};
    lc_phase_res<NT>(a, s, l, b, n, wsrc, sink);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int g_max_smem = -1;

static int max_optin_smem() {
    if (g_max_smem < 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) {
            cudaGetLastError();
            v = 0;
        }
        g_max_smem = v;
    }
    return g_max_smem;
}

bool resident_supported(const lc_args& a, int mode) {
    if (a.dtype != LC_F32 || a.N < kResidentMinN) return false;
    if ((mode & MODE_LM) && a.weight_mode != LC_W_ICOV_DIAG && a.weight_mode != LC_W_INV_STD) return false;
    const size_t need = resident_smem_bytes(a.N, false);
    return need <= static_cast<size_t>(max_optin_smem());
}

static int resident_threads_for(int n) {
    if (const char* e = getenv("LC_B200_RES_NT")) return atoi(e);   // tuning knob for benchmarks
    return n <= 512 ? 128 : 256;
}

// A (B,N,C) fp32 view can be staged by 1-D TMA bulk copies when every component slab is contiguous (point stride 1)
// and 16-byte aligned for every pose: base pointer, batch stride, component stride and N all multiples of 16 bytes.
static bool tma_ok(const lc_view& v, int n) {
    return v.stride[1] == 1 && (n % 4) == 0 && (reinterpret_cast<uintptr_t>(v.ptr) % 16) == 0 && (v.stride[0] % 4) == 0 &&
           (v.stride[2] % 4) == 0;
}
static int tma_mask_for(const lc_args& a) {
    if (getenv("LC_B200_NO_TMA")) return 0;
    return (tma_ok(a.pts3d, a.N) ? 1 : 0) | (tma_ok(a.pts2d, a.N) ? 2 : 0);
}

template <int NT, int MODE>
static int launch_res_t(const lc_args& a, cudaStream_t st) {
    const size_t smem = resident_smem_bytes(a.N, false);
    static size_t configured = 0;   // per instantiation
    if (smem > configured) {
        const cudaError_t e = cudaFuncSetAttribute(lc_resident_kernel<NT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(max_optin_smem()));
        if (e != cudaSuccess) return static_cast<int>(e);
        configured = static_cast<size_t>(max_optin_smem());
    }
    lc_resident_kernel<NT, MODE><<<a.B, NT, smem, st>>>(a, round_up4(a.N), tma_mask_for(a));
    return static_cast<int>(cudaGetLastError());
}

template <int MODE>
static int launch_res_m(const lc_args& a, cudaStream_t st) {
    // 256 threads x 2 CTAs per SM (20 B/point of shared memory): one CTA's 6x6 / trust-region sections and loads
    // overlap the other CTA's point passes
    if (resident_threads_for(a.N) == 128) return launch_res_t<128, MODE>(a, st);
    return launch_res_t<256, MODE>(a, st);
}

int launch_resident_pose(const lc_args& a, int mode, cudaStream_t st) {
    switch (mode) {
        case MODE_LM: return launch_res_m<MODE_LM>(a, st);
        case MODE_LC: return launch_res_m<MODE_LC>(a, st);
        default: return launch_res_m<MODE_LM | MODE_LC>(a, st);
    }
}

}  // namespace lc
