// Fast-path point passes of the resident kernels (lc_resident_kernel.cuh) for the layout the dense call site produces
// (planar fp32 slabs, point stride 1, 16-byte aligned: losses.py:142-161) — what bench.py and the training step feed.
//
// LC phase (lc_phase_vec): every thread owns GROUPS of four consecutive points (group g = tid + k NT, points 4g .. 4g+3):
//   * the on-chip arrays (q = R X, clamped error ec) are read and written as float4 (LDS.128 / STS.128),
//   * the weights re-read from L2 each pass are two LDG.128 per group instead of eight scalar loads + index arithmetic,
//   * the four points of a group give the scheduler four independent geometry chains (rcp, rsqrt, selects),
//   * the gradients are written IN PLACE over q (d/d pts3d) and ec (d/d inv_std) in shared memory and leave the SM as
//     1-D TMA bulk stores (cp.async.bulk.global.shared::cta, one per component slab, issued by one thread): TMA in, TMA out,
//     no store instruction in the point loop.
//   * the 6x6 accumulations use packed fp32 (fma.rn.f32x2, SASS FFMA2): H'[k] and G'[k] share the multiplicand J_j, so one
//     FFMA2 updates both (pass 3); the quadratic forms J^T Hbar J and J^T Gbar J are evaluated row-wise as pairs
//     (27 FFMA2 per coordinate instead of 21 FMUL + 42 FFMA, pass 4).
// The phase assumes the usual camera matrix (K row 2 = [0, 0, 1]) and no point with camera depth < 0.1 (the clamp of
// project_apply, transforms.py:62): then d proj / d P equals the translation block of the Jacobian rows already at hand.
// Poses that violate either (checked on the device, per pose) take the general scalar code (lc_resident.cuh).
// Points >= n (ragged batches, padding) get weight 0 / valid 0: they add nothing to any sum and receive zero gradients.
//
// LM pass (lm_eval_pass_planar): one point per thread and iteration (the fp64 pass is register-bound: 28 accumulators), unit
// strides known at compile time, weight transform hoisted into a template flag, Jacobian rows consumed one at a time.
//
// Math and guard semantics are exactly those of lc_phase_res / lm_eval_pass_res (lc_resident.cuh); cov_mixed.py:100-150.
#pragma once

#include "lc_resident.cuh"


namespace lc {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float f4get(const float4& v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }
__device__ __forceinline__ void f4set(float4& v, int j, float x) { if (j == 0) v.x = x; else if (j == 1) v.y = x; else if (j == 2) v.z = x; else v.w = x; }
__device__ __forceinline__ float2 dup2(float x) { return make_float2(x, x); }

// 1-D TMA bulk store shared -> global (SASS UBLKCP / STAS): one instruction per contiguous slab
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (TMA) that reads them next
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct VecIO {
    const float* w0;   // planar weight slabs of this pose (component 0 / 1), 16-byte aligned
    const float* w1;
    const float* valid;   // (N) of this pose, contiguous and aligned, or nullptr
    float *g3x, *g3y, *g3z;   // planar gradient slabs of this pose, or nullptr
    float *g2u, *g2v;
    float *gwu, *gwv;
};

__device__ __forceinline__ VecIO make_vec_io(const lc_args& a, int b) {
    VecIO io;
    const float* pw = static_cast<const float*>(a.weights.ptr) + b * a.weights.stride[0];
    io.w0 = pw; io.w1 = pw + a.weights.stride[2];
    io.valid = a.valid.ptr ? static_cast<const float*>(a.valid.ptr) + b * a.valid.stride[0] : nullptr;
    float* p = a.g_pts3d.ptr ? static_cast<float*>(a.g_pts3d.ptr) + b * a.g_pts3d.stride[0] : nullptr;
    io.g3x = p; io.g3y = p ? p + a.g_pts3d.stride[2] : nullptr; io.g3z = p ? p + 2 * a.g_pts3d.stride[2] : nullptr;
    p = a.g_pts2d.ptr ? static_cast<float*>(a.g_pts2d.ptr) + b * a.g_pts2d.stride[0] : nullptr;
    io.g2u = p; io.g2v = p ? p + a.g_pts2d.stride[2] : nullptr;
    p = a.g_weights.ptr ? static_cast<float*>(a.g_weights.ptr) + b * a.g_weights.stride[0] : nullptr;
    io.gwu = p; io.gwv = p ? p + a.g_weights.stride[2] : nullptr;
    return io;
}

__device__ __forceinline__ void mask_tail4(float4& v, int nl) {   // zero the components >= nl (nl = live points of the group)
    if (nl < 4) {
        v.w = 0.f;
        if (nl < 3) v.z = 0.f;
        if (nl < 2) v.y = 0.f;
        if (nl < 1) v.x = 0.f;
    }
}
// weights of group g with the points >= n zeroed
__device__ __forceinline__ void load_w4(const VecIO& io, int g, int n, float4& s0, float4& s1) {
    s0 = ldg4(io.w0 + 4 * g); s1 = ldg4(io.w1 + 4 * g);
    mask_tail4(s0, n - 4 * g); mask_tail4(s1, n - 4 * g);
}
__device__ __forceinline__ float4 load_valid4(const VecIO& io, int g, int n) {
    float4 v = io.valid ? ldg4(io.valid + 4 * g) : make_float4(1.f, 1.f, 1.f, 1.f);
    mask_tail4(v, n - 4 * g);
    return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// The four point passes as loop bodies over the groups g = tid, tid + NT, ... of one pose.  They only accumulate / write
// per-thread results; the callers combine them across the CTA (block_reduce in the CTA-per-pose kernel, per-warp partials
// handed to the serial warp in the persistent kernel, lc_persist.cuh).
// ---------------------------------------------------------------------------------------------------------------------

// pass 1 (fp64): P = R X + t, project_apply + clamp_error once per point (any K); q = R X and ec kept as fp32 in place.
// acc = [sum valid |ec_u|, sum valid |ec_v|, sum valid, #points at camera depth < 0.1]
template <int NT>
__device__ __forceinline__ void lc_pass1_vec(const lc_args& a, const PoseShared& s, const ResLayout& l, const VecIO& io, int n, int tid, float (&acc)[4]) {
    const int ng = (n + 3) >> 2;
    const double Lmax = a.max_err_len;
    const double lim = Lmax - 1e-6, lim2 = lim > 0.0 ? lim * lim : -1.0;   // |e|+1e-6 > Lmax  <=>  |e|^2 > lim2
    const double R0 = s.R[0], R1 = s.R[1], R2 = s.R[2], R3 = s.R[3], R4 = s.R[4], R5 = s.R[5], R6 = s.R[6], R7 = s.R[7], R8 = s.R[8];
    const double T0 = s.t[0], T1 = s.t[1], T2 = s.t[2];
    const double K0 = s.K[0], K1 = s.K[1], K2 = s.K[2], K3 = s.K[3], K4 = s.K[4], K5 = s.K[5], K6 = s.K[6], K7 = s.K[7], K8 = s.K[8];
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, nclamp = 0.f;
    for (int g = tid; g < ng; g += NT) {
        const int nl = n - 4 * g;
        float4 X0 = ld4(l.A0 + 4 * g), X1 = ld4(l.A1 + 4 * g), X2 = ld4(l.A2 + 4 * g);
        float4 xu = ld4(l.B0 + 4 * g), xv = ld4(l.B1 + 4 * g);
        // slots beyond n may hold anything (TMA copies the whole padded slab): neutralise them
        mask_tail4(X0, nl); mask_tail4(X1, nl); mask_tail4(X2, nl); mask_tail4(xu, nl); mask_tail4(xv, nl);
        const float4 v = load_valid4(io, g, n);
        float4 eu, ev, q0v, q1v, q2v;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double x0 = f4get(X0, j), x1 = f4get(X1, j), x2 = f4get(X2, j), m0 = f4get(xu, j), m1 = f4get(xv, j);
            const double q0 = fma(R0, x0, fma(R1, x1, R2 * x2));
            const double q1 = fma(R3, x0, fma(R4, x1, R5 * x2));
            const double q2 = fma(R6, x0, fma(R7, x1, R8 * x2));
            const double P0 = q0 + T0, P1 = q1 + T1, P2 = q2 + T2;
            const double KP0 = fma(K0, P0, fma(K1, P1, K2 * P2));
            const double KP1 = fma(K3, P0, fma(K4, P1, K5 * P2));
            const double KP2 = fma(K6, P0, fma(K7, P1, K8 * P2));
            const bool clamped = !(KP2 > 0.1);
            const double iz = fast_rcp(clamped ? 0.1 : KP2);
            double e0 = fma(-KP0, iz, m0), e1 = fma(-KP1, iz, m1);
            const double l2 = fma(e0, e0, e1 * e1);
            if (l2 > lim2) {
                const double len = sqrt(l2) + 1e-6;
                const double f = (len - Lmax) / len;
                e0 = fma(-f, e0, e0);
                e1 = fma(-f, e1, e1);
            }
            const float ec0 = static_cast<float>(e0), ec1 = static_cast<float>(e1);
            f4set(eu, j, ec0); f4set(ev, j, ec1);
            // q = R X is what is cached (not P = q + t): q x D then keeps fp32 relative precision even when |X| << |t|
            f4set(q0v, j, static_cast<float>(q0)); f4set(q1v, j, static_cast<float>(q1)); f4set(q2v, j, static_cast<float>(q2));
            const float vj = f4get(v, j);
            acc0 = fmaf(vj, fabsf(ec0), acc0); acc1 = fmaf(vj, fabsf(ec1), acc1); acc2 += vj;
            if (clamped && j < nl) nclamp += 1.f;
        }
        st4(l.B0 + 4 * g, eu); st4(l.B1 + 4 * g, ev);
        st4(l.A0 + 4 * g, q0v); st4(l.A1 + 4 * g, q1v); st4(l.A2 + 4 * g, q2v);
    }
    acc[0] = acc0; acc[1] = acc1; acc[2] = acc2; acc[3] = nclamp;
}

// pass 2 (fp32): acc = [sum valid s_u^2 sigma_u, sum valid s_v^2 sigma_v]
template <int NT>
__device__ __forceinline__ void lc_pass2_vec(const ResLayout& l, const VecIO& io, int n, int tid, float d0, float d1, float (&acc)[2]) {
    const int ng = (n + 3) >> 2;
    float acc0 = 0.f, acc1 = 0.f;
    // the weights come from L2: the next group's are requested before the current group is processed
    float4 sn0 = make_float4(0.f, 0.f, 0.f, 0.f), sn1 = sn0;
    if (tid < ng) load_w4(io, tid, n, sn0, sn1);
    for (int g = tid; g < ng; g += NT) {
        const float4 s0 = sn0, s1 = sn1;
        if (g + NT < ng) load_w4(io, g + NT, n, sn0, sn1);
        const float4 v = load_valid4(io, g, n);
        const float4 eu = ld4(l.B0 + 4 * g), ev = ld4(l.B1 + 4 * g);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a0 = fabsf(f4get(eu, j)), a1 = fabsf(f4get(ev, j)), sj0 = f4get(s0, j), sj1 = f4get(s1, j), vj = f4get(v, j);
            const float sg0 = a0 > d0 ? d0 * (2.f * a0 - d0) : a0 * a0;
            const float sg1 = a1 > d1 ? d1 * (2.f * a1 - d1) : a1 * a1;
            acc0 = fmaf(vj * (sj0 * sj0), sg0, acc0);
            acc1 = fmaf(vj * (sj1 * sj1), sg1, acc1);
        }
    }
    acc[0] = acc0; acc[1] = acc1;
}

// pass 3 (packed fp32 partial sums): accd = [H' = sum W J'J'^T (21), G' = sum W^2 sigma J'J'^T (21), b' = sum W ec J' (6)]
template <int NT>
__device__ __forceinline__ void lc_pass3_vec(const ResLayout& l, const VecIO& io, int n, int tid, const PointConsts& pc, float (&accd)[48]) {
    const int ng = (n + 3) >> 2;
    float2 hg[kSym];     // (H'[k], G'[k])
    float2 bb[3];        // b' as pairs
#pragma unroll
    for (int k2 = 0; k2 < kSym; ++k2) hg[k2] = make_float2(0.f, 0.f);
#pragma unroll
    for (int k2 = 0; k2 < 3; ++k2) bb[k2] = make_float2(0.f, 0.f);
    float4 sn0 = make_float4(0.f, 0.f, 0.f, 0.f), sn1 = sn0;
    if (tid < ng) load_w4(io, tid, n, sn0, sn1);
    for (int g = tid; g < ng; g += NT) {
        const float4 Q0 = ld4(l.A0 + 4 * g), Q1 = ld4(l.A1 + 4 * g), Q2 = ld4(l.A2 + 4 * g);
        const float4 s0 = sn0, s1 = sn1;
        if (g + NT < ng) load_w4(io, g + NT, n, sn0, sn1);
        const float4 eu = ld4(l.B0 + 4 * g), ev = ld4(l.B1 + 4 * g);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float J[2][6], sg[2], del[2], w[2];
            const float ecj[2] = {f4get(eu, j), f4get(ev, j)}, skj[2] = {f4get(s0, j), f4get(s1, j)};
            point_terms_f(pc, f4get(Q0, j), f4get(Q1, j), f4get(Q2, j), ecj, skj, J, sg, del, w);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float2 wg = make_float2(w[c], w[c] * w[c] * sg[c]);
                float2 Jd[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) Jd[r] = dup2(J[c][r]);
                int k2 = 0;
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    const float2 A = __fmul2_rn(wg, Jd[r]);   // (w J_r, g J_r)
#pragma unroll
                    for (int cc = r; cc < 6; ++cc) { hg[k2] = __ffma2_rn(A, Jd[cc], hg[k2]); ++k2; }
                }
                const float2 wb = dup2(w[c] * ecj[c]);
#pragma unroll
                for (int m = 0; m < 3; ++m) bb[m] = __ffma2_rn(wb, make_float2(J[c][2 * m], J[c][2 * m + 1]), bb[m]);
            }
        }
    }
#pragma unroll
    for (int k2 = 0; k2 < kSym; ++k2) { accd[k2] = hg[k2].x; accd[kSym + k2] = hg[k2].y; }
#pragma unroll
    for (int m = 0; m < 3; ++m) { accd[42 + 2 * m] = bb[m].x; accd[43 + 2 * m] = bb[m].y; }
}

// pass 4 (packed fp32), usual camera matrix and no clamped depth: per-coordinate adjoints; d/d pts3d overwrites q,
// d/d inv_std (or d/d pts2d) overwrites ec in shared memory (the caller stores the slabs with the TMA); padding slots beyond
// the last group are zeroed in global memory directly.  Ends with the proxy fence the TMA store needs.
template <int NT>
__device__ __forceinline__ void lc_pass4_vec(const lc_args& a, const PoseShared& s, const ResLayout& l, const VecIO& io, int n, int tid, const PointConsts& pc) {
    const int ng = (n + 3) >> 2;
    float2 chg[kSym];   // (cH[k], cG[k])
    float bL[6];
#pragma unroll
    for (int k2 = 0; k2 < kSym; ++k2) chg[k2] = make_float2(static_cast<float>(s.cHL[k2]), static_cast<float>(s.cGL[k2]));
#pragma unroll
    for (int k2 = 0; k2 < 6; ++k2) bL[k2] = static_cast<float>(s.bL[k2]);
    // gX = -R^T gP: columns of R as float constants
    const float R0 = static_cast<float>(s.R[0]), R1 = static_cast<float>(s.R[1]), R2 = static_cast<float>(s.R[2]);
    const float R3 = static_cast<float>(s.R[3]), R4 = static_cast<float>(s.R[4]), R5 = static_cast<float>(s.R[5]);
    const float R6 = static_cast<float>(s.R[6]), R7 = static_cast<float>(s.R[7]), R8 = static_cast<float>(s.R[8]);
    const bool want3 = io.g3x != nullptr, wantw = io.gwu != nullptr, want2 = io.g2u != nullptr;
    float4 sn0 = make_float4(0.f, 0.f, 0.f, 0.f), sn1 = sn0;
    if (tid < ng) load_w4(io, tid, n, sn0, sn1);
    for (int g = tid; g < ng; g += NT) {
        float4 Q0 = ld4(l.A0 + 4 * g), Q1 = ld4(l.A1 + 4 * g), Q2 = ld4(l.A2 + 4 * g);
        const float4 s0 = sn0, s1 = sn1;
        if (g + NT < ng) load_w4(io, g + NT, n, sn0, sn1);
        float4 eu = ld4(l.B0 + 4 * g), ev = ld4(l.B1 + 4 * g);
        float4 gw0, gw1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float q0 = f4get(Q0, j), q1 = f4get(Q1, j), q2 = f4get(Q2, j);
            const float ecj[2] = {f4get(eu, j), f4get(ev, j)}, skj[2] = {f4get(s0, j), f4get(s1, j)};
            // geometry as point_terms_f, one coordinate at a time (keeps 6 instead of 12 Jacobian entries live)
            const float P0 = q0 + pc.t0, P1 = q1 + pc.t1, P2 = q2 + pc.t2;
            const float iz = __fdividef(1.f, P2);
            const float u0 = P0 * iz, v0 = P1 * iz;
            const float du0 = fmaf(pc.uc, q2, -q0) * iz, dv0 = fmaf(pc.vc, q2, -q1) * iz;
            float gP0 = 0.f, gP1 = 0.f, gP2 = 0.f;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float ka = c ? pc.k10 : pc.k00, kb = c ? pc.k11 : pc.k01;
                const float e0 = ka * iz, e1 = kb * iz, e2 = -fmaf(ka, u0, kb * v0) * iz;
                float J[6];
                J[0] = fmaf(q1, e2, -q2 * e1);
                J[1] = fmaf(q2, e0, -q0 * e2);
                J[2] = fmaf(q0, e1, -q1 * e0);
                J[3] = e0; J[4] = e1; J[5] = fmaf(ka, du0, kb * dv0) * iz;
                const float dc = c ? pc.d1 : pc.d0, sq = c ? pc.sq1 : pc.sq0;
                const float av = fabsf(ecj[c]);
                const float sg = av > dc ? dc * (2.f * av - dc) : av * av;
                const float del = sq * rsqrtf(sg + 1e-6f);
                const float w = skj[c] > del ? del * (2.f * skj[c] - del) : skj[c] * skj[c];
                float2 Jd[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) Jd[r] = dup2(J[r]);
                // (qh, qg) = sum_r J_r sum_{cc >= r} (cH, cG)_{r,cc} J_cc   (off-diagonal coefficients are doubled)
                float2 qq = make_float2(0.f, 0.f);
                float lb = 0.f;
                int k2 = 0;
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    float2 row = __fmul2_rn(chg[k2], Jd[r]);
                    ++k2;
#pragma unroll
                    for (int cc = r + 1; cc < 6; ++cc) { row = __ffma2_rn(chg[k2], Jd[cc], row); ++k2; }
                    qq = __ffma2_rn(Jd[r], row, qq);
                    lb = fmaf(J[r], bL[r], lb);
                }
                const float qh = qq.x, qg = qq.y;
                const float Wbar = qh + 2.f * w * sg * qg + ecj[c] * lb;
                const float sigbar = w * w * qg;
                const float gwv = Wbar * (skj[c] > del ? 2.f * del : 2.f * skj[c]);
                if (c == 0) f4set(gw0, j, gwv); else f4set(gw1, j, gwv);
                const float sgn = (ecj[c] > 0.f) ? 1.f : ((ecj[c] < 0.f) ? -1.f : 0.f);
                const float ecb = sigbar * (av > dc ? 2.f * dc : 2.f * av) * sgn;
                if (c == 0) f4set(eu, j, ecb); else f4set(ev, j, ecb);   // d/d pts2d replaces ec
                // d proj_c / d P = (e0, e1, e2) for K row 2 = e_z and an inactive depth clamp
                gP0 = fmaf(ecb, e0, gP0); gP1 = fmaf(ecb, e1, gP1); gP2 = fmaf(ecb, e2, gP2);
            }
            // gX = -R^T gP replaces q
            f4set(Q0, j, -(R0 * gP0 + R3 * gP1 + R6 * gP2));
            f4set(Q1, j, -(R1 * gP0 + R4 * gP1 + R7 * gP2));
            f4set(Q2, j, -(R2 * gP0 + R5 * gP1 + R8 * gP2));
        }
        if (want3) { st4(l.A0 + 4 * g, Q0); st4(l.A1 + 4 * g, Q1); st4(l.A2 + 4 * g, Q2); }
        if (wantw) { st4(l.B0 + 4 * g, gw0); st4(l.B1 + 4 * g, gw1); }
        if (want2) {
            if (wantw) { st4(io.g2u + 4 * g, eu); st4(io.g2v + 4 * g, ev); }   // both wanted: d/d pts2d goes out directly
            else { st4(l.B0 + 4 * g, eu); st4(l.B1 + 4 * g, ev); }
        }
    }
    // gradient slots of the padding beyond the last group (ragged batches): defined, zero
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int g = ng + tid; g < (a.N >> 2); g += NT) {
        if (wantw) { st4(io.gwu + 4 * g, z4); st4(io.gwv + 4 * g, z4); }
        if (want2) { st4(io.g2u + 4 * g, z4); st4(io.g2v + 4 * g, z4); }
        if (want3) { st4(io.g3x + 4 * g, z4); st4(io.g3y + 4 * g, z4); st4(io.g3z + 4 * g, z4); }
    }
    fence_async_smem();
}
// one thread, after a barrier that follows lc_pass4_vec: shared memory -> global (slots in [n, 4 ng) are zero by construction)
__device__ __forceinline__ void lc_pass4_store(const ResLayout& l, const VecIO& io, int n) {
    const unsigned bytes = static_cast<unsigned>((n + 3) >> 2) * 16u;
    if (bytes == 0) return;
    const bool want3 = io.g3x != nullptr, wantw = io.gwu != nullptr, want2 = io.g2u != nullptr;
    if (want3) { tma_store_1d(io.g3x, l.A0, bytes); tma_store_1d(io.g3y, l.A1, bytes); tma_store_1d(io.g3z, l.A2, bytes); }
    if (wantw) { tma_store_1d(io.gwu, l.B0, bytes); tma_store_1d(io.gwv, l.B1, bytes); }
    else if (want2) { tma_store_1d(io.g2u, l.B0, bytes); tma_store_1d(io.g2v, l.B1, bytes); }
    tma_store_commit_wait();
}

// robust thresholds from the pass-1 / pass-2 sums (cov_mixed.py:28-37)
__device__ __forceinline__ void lc_thresholds1(const lc_args& a, const double* fin, int n, double& vcnt, float& d0, float& d1, bool& any_clamped) {
    vcnt = a.valid.ptr ? fin[2] : static_cast<double>(n);
    any_clamped = fin[3] > 0.0;   // some point sits at camera depth < 0.1: pass 4 takes the general code
    const double iv = fast_rcp(vcnt);
    d0 = static_cast<float>(a.rel_thresh * (fin[0] * iv)); d1 = static_cast<float>(a.rel_thresh * (fin[1] * iv));
}
// delta_k = sqrt(we * q_a / (sigma_k + 1e-6)) = sq_a * rsqrt(sigma_k + 1e-6)
__device__ __forceinline__ void lc_thresholds2(const lc_args& a, const double* fin, double vcnt, float& sq0, float& sq1) {
    const double iv = fast_rcp(vcnt);
    sq0 = static_cast<float>(sqrt((fin[0] * iv) * a.w_e_thresh)); sq1 = static_cast<float>(sqrt((fin[1] * iv) * a.w_e_thresh));
}

// LC phase on the staged arrays (A = X -> q -> d/d pts3d, B = x -> ec -> d/d inv_std), one CTA per pose.  Same contract as
// lc_phase_res.
template <int NT, class CLU>
__device__ __forceinline__ void lc_phase_vec(const lc_args& a, PoseShared& s, const ResLayout& l, int b, int n, CLU& cl) {
    const int tid = threadIdx.x;
    const VecIO io = make_vec_io(a, b);
    const bool want_any = a.g_pts3d.ptr || a.g_pts2d.ptr || a.g_weights.ptr;
    { LC_TIC(tq1);
    if (tid < 32) lc_pose_setup_warp(s, true);
    __syncthreads();
    LC_TOC(tq1, 3); }
    LC_TIC(tq2);
    // the usual camera matrix: K row 2 = [0, 0, 1] (uniform per pose)
    const bool fastK = s.K[6] == 0.0 && s.K[7] == 0.0 && s.K[8] == 1.0;
    double vcnt;
    float d0, d1, sq0, sq1;
    bool any_clamped;
    {
        float acc[4];
        lc_pass1_vec<NT>(a, s, l, io, n, tid, acc);
        block_reduce_f<4, NT>(acc, s.red, s.fin);
        cl.template combine<4>(s.fin);
        lc_thresholds1(a, s.fin, cl.n_total, vcnt, d0, d1, any_clamped);
        __syncthreads();
    }
    {
        float acc[2];
        lc_pass2_vec<NT>(l, io, n, tid, d0, d1, acc);
        block_reduce_f<2, NT>(acc, s.red, s.fin);
        cl.template combine<2>(s.fin);
        lc_thresholds2(a, s.fin, vcnt, sq0, sq1);
        __syncthreads();
    }
    const PointConsts pc = make_point_consts(s, d0, d1, sq0, sq1);
    {
        float accd[48];
        lc_pass3_vec<NT>(l, io, n, tid, pc, accd);
        block_reduce_f<48, NT>(accd, s.red, s.fin);
        cl.template combine<48>(s.fin);
    }
    LC_TOC(tq2, 4);
    { LC_TIC(tq3);
    if (tid < 32) lc_six_fast<float>(a, s, b, want_any, cl.leader);
    __syncthreads();
    if (!want_any) return;
    LC_TOC(tq3, 5); }
    LC_TIC(tq4);
    if (!fastK || any_clamped) {   // general d proj / d P: scalar pass 4 writing straight to global memory
        const XAcc<false> xs{l, 0u};
        const DirectWeights wsrc{io.w0, 1, a.weights.stride[2]};
        DirectSink sink{a, b};
        lc_pass4_scalar<NT>(a, s, l, n, d0, d1, sq0, sq1, wsrc, sink, xs);
        return;
    }
    lc_pass4_vec<NT>(a, s, l, io, n, tid, pc);
    __syncthreads();
    if (tid == 0) lc_pass4_store(l, io, n);
#ifdef LC_TIMING
    __syncthreads();
#endif
    LC_TOC(tq4, 6);
}

// Accumulation loop of one evaluation pass of the reprojection cost (ceres.cpp:30-55) at the point held in L.Rm/L.te, for
// planar weights (unit point stride).  acc = [J'^T J' packed (21), J'^T r (6), cost] in the left basis; same contract as
// lm_eval_pass_res.  WGEN = false: the weights are inverse std / sqrt-information factors (la = |w|); WGEN = true: any
// diagonal mode incl. nan_to_num and sqrt(icov).
template <int NT, bool JAC, bool WGEN>
__device__ __forceinline__ void lm_eval_accum_planar(const lc_args& a, const PoseShared& s, const ResLayout& l, int b, int n, bool sanitize, int tid,
                                                     double (&acc)[28]) {
    const LmState& L = s.lm;
#pragma unroll
    for (int k = 0; k < 28; ++k) acc[k] = 0.0;
    const double k00 = s.K[0], k01 = s.K[1], k10 = s.K[3], k11 = s.K[4];
    const double R0 = L.Rm[0], R1 = L.Rm[1], R2 = L.Rm[2], R3 = L.Rm[3], R4 = L.Rm[4], R5 = L.Rm[5], R6 = L.Rm[6], R7 = L.Rm[7], R8 = L.Rm[8];
    const double t0 = L.te[0], t1 = L.te[1], t2 = L.te[2];
    const volatile double* Kv = s.K;   // cx, cy are re-read from shared memory where used (keeps 4 registers free)
    const float* w0p = static_cast<const float*>(a.weights.ptr) + b * a.weights.stride[0];
    const float* w1p = w0p + a.weights.stride[2];
    const bool icov = WGEN && a.weight_mode == LC_W_ICOV_DIAG;
    // the sqrt-information weights are re-read from L2 every pass; the next point's weights are fetched before the
    // current point is processed
    float wn0 = 0.f, wn1 = 0.f;
    if (tid < n) { wn0 = __ldg(w0p + tid); wn1 = __ldg(w1p + tid); }
    for (int i = tid; i < n; i += NT) {
        float wa = wn0, wb = wn1;
        const int inext = i + NT;
        if (inext < n) { wn0 = __ldg(w0p + inext); wn1 = __ldg(w1p + inext); }
        if (WGEN) {
            if (sanitize) { wa = nan_to_num_f(wa); wb = nan_to_num_f(wb); }
            // cer_solver.py:37-38: L = diag(sqrt(icov)) in fp32.  For LC_W_INV_STD the reference's sqrt(fl(s*s)) is exactly |s|
            if (icov) { wa = sqrtf(wa); wb = sqrtf(wb); }
        }
        const double la = fabsf(wa), lc_ = fabsf(wb);
        const double X0 = l.A0[i], X1 = l.A1[i], X2 = l.A2[i];
        const double px = l.B0[i], py = l.B1[i];
        const double q0 = fma(R0, X0, fma(R1, X1, R2 * X2));
        const double q1 = fma(R3, X0, fma(R4, X1, R5 * X2));
        const double q2 = fma(R6, X0, fma(R7, X1, R8 * X2));
        const double p0 = q0 + t0, p1 = q1 + t1, p2 = q2 + t2;
        const double iz = fast_rcp(p2);
        const double up = fma(p0, k00, p1 * k01) * iz, vp = fma(p0, k10, p1 * k11) * iz;
        const double du = up - (px - Kv[2]), dv = vp - (py - Kv[5]);
        const double r0 = du * la, r1 = dv * lc_;
        acc[27] = fma(r0, r0, fma(r1, r1, acc[27]));
        if (JAC) {
            // the two Jacobian rows one after the other: six entries live at a time
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const double ac = (c ? lc_ : la) * iz, rc = c ? r1 : r0;
                double J[6];
                J[3] = ac * (c ? k10 : k00); J[4] = ac * (c ? k11 : k01); J[5] = -ac * (c ? vp : up);
                J[0] = fma(q1, J[5], -q2 * J[4]); J[1] = fma(q2, J[3], -q0 * J[5]); J[2] = fma(q0, J[4], -q1 * J[3]);
                int k2 = 0;
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int cc = r; cc < 6; ++cc) { acc[k2] = fma(J[r], J[cc], acc[k2]); ++k2; }
#pragma unroll
                for (int cc = 0; cc < 6; ++cc) acc[21 + cc] = fma(J[cc], rc, acc[21 + cc]);
            }
        }
    }
    acc[27] *= 0.5;
}

// Mixed-precision Jacobian pass (LC_FLAG_LM_MIXED).  fp64: q = R X, p, 1/z, the projection, the residual and the cost, exactly
// as in lm_eval_accum_planar (every trust-region decision keeps its fp64 inputs).  fp32: the two Jacobian rows, stored as
// pairs (J0_i, J1_i), and their sums — J'^T J' as 21 FFMA2 whose two lanes are the two rows (added in fp64 at the end), J'^T r
// as 6 FFMA2.  Per point: 29 fp64 operations instead of 103; the per-thread fp32 partial sums (<= N/NT points) are reduced
// across the CTA in fp64.
template <int NT, bool WGEN>
__device__ __forceinline__ void lm_eval_accum_mixed(const lc_args& a, const PoseShared& s, const ResLayout& l, int b, int n, bool sanitize, int tid,
                                                    double (&acc)[28]) {
    const LmState& L = s.lm;
    float2 A2[kSym], g2[6];
#pragma unroll
    for (int k = 0; k < kSym; ++k) A2[k] = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 6; ++k) g2[k] = make_float2(0.f, 0.f);
    double cost = 0.0;
    const double k00 = s.K[0], k01 = s.K[1], k10 = s.K[3], k11 = s.K[4];
    const float kf00 = static_cast<float>(k00), kf01 = static_cast<float>(k01), kf10 = static_cast<float>(k10), kf11 = static_cast<float>(k11);
    const double R0 = L.Rm[0], R1 = L.Rm[1], R2 = L.Rm[2], R3 = L.Rm[3], R4 = L.Rm[4], R5 = L.Rm[5], R6 = L.Rm[6], R7 = L.Rm[7], R8 = L.Rm[8];
    const double t0 = L.te[0], t1 = L.te[1], t2 = L.te[2];
    const volatile double* Kv = s.K;   // cx, cy are re-read from shared memory where used
    const float* w0p = static_cast<const float*>(a.weights.ptr) + b * a.weights.stride[0];
    const float* w1p = w0p + a.weights.stride[2];
    const bool icov = WGEN && a.weight_mode == LC_W_ICOV_DIAG;
    float wn0 = 0.f, wn1 = 0.f;
    if (tid < n) { wn0 = __ldg(w0p + tid); wn1 = __ldg(w1p + tid); }
    for (int i = tid; i < n; i += NT) {
        float wa = wn0, wb = wn1;
        const int inext = i + NT;
        if (inext < n) { wn0 = __ldg(w0p + inext); wn1 = __ldg(w1p + inext); }
        if (WGEN) {
            if (sanitize) { wa = nan_to_num_f(wa); wb = nan_to_num_f(wb); }
            if (icov) { wa = sqrtf(wa); wb = sqrtf(wb); }
        }
        const float laf = fabsf(wa), lcf = fabsf(wb);
        const double X0 = l.A0[i], X1 = l.A1[i], X2 = l.A2[i];
        const double px = l.B0[i], py = l.B1[i];
        const double q0 = fma(R0, X0, fma(R1, X1, R2 * X2));
        const double q1 = fma(R3, X0, fma(R4, X1, R5 * X2));
        const double q2 = fma(R6, X0, fma(R7, X1, R8 * X2));
        const double p0 = q0 + t0, p1 = q1 + t1, p2 = q2 + t2;
        const double iz = fast_rcp(p2);
        const double up = fma(p0, k00, p1 * k01) * iz, vp = fma(p0, k10, p1 * k11) * iz;
        const double du = up - (px - Kv[2]), dv = vp - (py - Kv[5]);
        const double r0 = du * static_cast<double>(laf), r1 = dv * static_cast<double>(lcf);
        cost = fma(r0, r0, fma(r1, r1, cost));
        // ---- fp32 from here: Jacobian rows in the left basis, J'_c = [q x D_c | D_c] ----
        const float qf0 = static_cast<float>(q0), qf1 = static_cast<float>(q1), qf2 = static_cast<float>(q2);
        const float izf = static_cast<float>(iz), upf = static_cast<float>(up), vpf = static_cast<float>(vp);
        const float2 rr = make_float2(static_cast<float>(r0), static_cast<float>(r1));
        const float a0 = laf * izf, a1 = lcf * izf;
        float2 P[6];   // (row 0, row 1) of column i
        P[3] = make_float2(a0 * kf00, a1 * kf10);
        P[4] = make_float2(a0 * kf01, a1 * kf11);
        P[5] = make_float2(-a0 * upf, -a1 * vpf);
        const float2 Q0 = make_float2(qf0, qf0), Q1 = make_float2(qf1, qf1), Q2 = make_float2(qf2, qf2);
        const float2 nQ0 = make_float2(-qf0, -qf0), nQ1 = make_float2(-qf1, -qf1), nQ2 = make_float2(-qf2, -qf2);
        P[0] = __ffma2_rn(Q1, P[5], __fmul2_rn(nQ2, P[4]));
        P[1] = __ffma2_rn(Q2, P[3], __fmul2_rn(nQ0, P[5]));
        P[2] = __ffma2_rn(Q0, P[4], __fmul2_rn(nQ1, P[3]));
        int k2 = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int cc = r; cc < 6; ++cc) { A2[k2] = __ffma2_rn(P[r], P[cc], A2[k2]); ++k2; }
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) g2[cc] = __ffma2_rn(P[cc], rr, g2[cc]);
    }
#pragma unroll
    for (int k = 0; k < kSym; ++k) acc[k] = static_cast<double>(A2[k].x) + static_cast<double>(A2[k].y);
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[21 + k] = static_cast<double>(g2[k].x) + static_cast<double>(g2[k].y);
    acc[27] = 0.5 * cost;
}

template <int NT, bool WGEN, class CLU>
__device__ __forceinline__ void lm_eval_pass_mixed(const lc_args& a, PoseShared& s, const ResLayout& l, int b, int n, bool sanitize, CLU& cl) {
    double acc[28];
    lm_eval_accum_mixed<NT, WGEN>(a, s, l, b, n, sanitize, threadIdx.x, acc);
    block_reduce<28, NT>(acc, s.red, s.fin);
    cl.template combine<28>(s.fin);
}

template <int NT, bool JAC, bool WGEN, class CLU>
__device__ __forceinline__ void lm_eval_pass_planar(const lc_args& a, PoseShared& s, const ResLayout& l, int b, int n, bool sanitize, CLU& cl) {
    double acc[28];
    lm_eval_accum_planar<NT, JAC, WGEN>(a, s, l, b, n, sanitize, threadIdx.x, acc);
    block_reduce<28, NT>(acc, s.red, s.fin);
    cl.template combine<28>(s.fin);
}

}  // namespace lc
