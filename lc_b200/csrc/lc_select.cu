// lc_b200 — test-time point selection on the device (include/lc_b200.h: lc_select_args).
//
// Replaces the selection half of test.solve_pnp_dense (test.py:67-119): joint (or per-channel) softmax of the weight logits
// times the scale (:84-88), the stride-`sample` sub-sampling of the pixel grid / weights / xyz / segmentation mask with
// top_left = (0,0) (losses.py:142-161), inv_cov = inv_std^2 (:95), the 'mask' / 'quantile' / 'quantile_in_mask' rules
// (:97-104, quantile_msk :36-45 = torch.quantile with linear interpolation) and `nonzero()` + ragged python lists (:106-119)
// -> ONE launch, one CTA per sample, that writes zero-padded (B,Nmax,.) correspondences in selection order plus n_points:
// exactly what cer_solver._batch_tensors (cer_solver.py:67-87) would build, without the device->host sync of nonzero().
//
// The per-sample quantile is an exact order statistic: a radix select over the fp32 bit patterns held in
// shared memory (values are >= 0, so the bit patterns are ordered; round 2: three passes of 11 + 11 + 10 bits), followed by torch's fp32 lerp.
//
// Round 2: the same one-CTA-per-sample structure with its serial pieces removed — warp-aggregated histogram atomics (the keys of a
// sample share their leading byte, plain atomics serialised 32-way), warp-scan bin search, one scan over all (round, warp) cells
// for the compaction instead of three barriers per 1024 pixels, 16-byte loads for the softmax statistics and 16-byte stores for
// the zero padding (the padding is most of the written bytes).  profiles/README.md has the before / after.
#include <atomic>

#include "lc_resident.cuh"

namespace lc {

// 512 threads x 64 registers: two CTAs per SM (1024 threads allowed one, and B = 256 samples then ran as two waves on 148 SMs)
constexpr int kSelNT = 512;

__device__ __forceinline__ float sel_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// value of the rank-k (0-based) smallest key among u[0..n) and of its successor in sorted order; all threads call.
// Three radix passes over the 32 bits of the pattern (11 + 11 + 10).  The keys of one sample share their leading
// bits (similar magnitudes), so in the FIRST pass a plain shared-memory atomicAdd per key would serialise on a handful of bins:
// the lanes of a warp that hit the same bin are aggregated with match.any (one atomic per distinct bin and warp); the trailing
// bits are spread evenly and take plain atomics.  The bin search is a warp scan (64 bins per lane), not one thread walking them.
constexpr int kSelBins = 2048;
__device__ void radix_select_pair(const unsigned* u, int n, int k, unsigned* hist, unsigned* bc, unsigned& kth, unsigned& next) {
    const int tid = threadIdx.x, lane = tid & 31;
    unsigned prefix = 0, mask = 0;
    int kk = k;
    for (int j = tid; j < kSelBins; j += kSelNT) hist[j] = 0;
    __syncthreads();
    const int n_up = (n + 31) & ~31;   // warp-uniform trip count (match.any is warp-collective)
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        const int sh = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
        const unsigned bm = pass == 2 ? 1023u : 2047u;
        if (pass == 0) {
            for (int i = tid; i < n_up; i += kSelNT) {
                const bool act = i < n;
                const unsigned key = act ? ((u[i] >> sh) & bm) : 0xFFFFu;
                const unsigned peers = __match_any_sync(kFull, key);
                if (act && lane == __ffs(peers) - 1) atomicAdd(&hist[key], static_cast<unsigned>(__popc(peers)));
            }
        } else {
            for (int i = tid; i < n; i += kSelNT) {
                const unsigned v = u[i];
                if ((v & mask) == prefix) atomicAdd(&hist[(v >> sh) & bm], 1u);
            }
        }
        __syncthreads();
        if (tid < 32) {
            constexpr int PER = kSelBins / 32;
            unsigned sum = 0;
#pragma unroll 8
            for (int j = 0; j < PER; ++j) sum += hist[PER * lane + j];
            unsigned incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += t; }
            const unsigned excl = incl - sum, want = static_cast<unsigned>(kk);
            if (excl <= want && want < incl) {   // exactly one lane: the one whose bins contain rank kk
                unsigned cum = excl;
                int bin = PER * lane;
                for (int j = 0; j < PER; ++j) {
                    const unsigned c = hist[PER * lane + j];
                    if (cum + c > want) break;
                    cum += c; ++bin;
                }
                bc[0] = static_cast<unsigned>(bin);
                bc[1] = cum;
            }
        }
        __syncthreads();
        prefix |= bc[0] << sh;
        mask |= bm << sh;
        kk -= static_cast<int>(bc[1]);
        if (pass < 2) for (int j = tid; j < kSelBins; j += kSelNT) hist[j] = 0;   // ready for the next pass
        __syncthreads();
    }
    kth = prefix;
    // successor: the same value if rank k+1 still falls on it, else the smallest larger key
    if (tid == 0) { bc[0] = 0; bc[1] = 0xFFFFFFFFu; }
    __syncthreads();
    unsigned cle = 0, mn = 0xFFFFFFFFu;
    for (int i = tid; i < n; i += kSelNT) {
        const unsigned v = u[i];
        if (v <= kth) ++cle; else mn = min(mn, v);
    }
    cle = __reduce_add_sync(kFull, cle);
    mn = __reduce_min_sync(kFull, mn);
    if (lane == 0) { atomicAdd(&bc[0], cle); atomicMin(&bc[1], mn); }
    __syncthreads();
    next = (bc[0] >= static_cast<unsigned>(k) + 2u || bc[1] == 0xFFFFFFFFu) ? kth : bc[1];
    __syncthreads();
}

// zero the floats [beg, end) of a contiguous array: 16-byte stores in the aligned middle
__device__ __forceinline__ void zero_flat(float* p, int64_t beg, int64_t end, int tid) {
    if (beg >= end) return;
    const int64_t mis = (reinterpret_cast<uintptr_t>(p + beg) >> 2) & 3;
    int64_t a0 = beg + ((4 - mis) & 3);
    if (a0 > end) a0 = end;
    const int64_t a1 = a0 + ((end - a0) & ~int64_t(3));
    if (tid < a0 - beg) p[beg + tid] = 0.f;
    if (tid < end - a1) p[a1 + tid] = 0.f;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* q = reinterpret_cast<float4*>(p + a0);
    for (int64_t j = tid; j < ((a1 - a0) >> 2); j += kSelNT) q[j] = z;
}

constexpr int kSelMaxRounds = 96;   // N <= 96 * 512 sampled pixels per sample (shared memory holds 5 B each: <= ~45 k anyway)

__global__ void __launch_bounds__(kSelNT, 2) lc_select_kernel(const lc_select_args d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Hn = (d.H + d.sample - 1) / d.sample, Wn = (d.W + d.sample - 1) / d.sample, N = Hn * Wn, HW = d.H * d.W;
    float* vals = reinterpret_cast<float*>(smem_raw);                    // [N] quantile operand
    unsigned char* mflag = reinterpret_cast<unsigned char*>(vals + N);   // [N] segmentation mask of the sampled pixel
    __shared__ unsigned hist[kSelBins];
    __shared__ unsigned bc[4];
    __shared__ float redf[kSelNT / 32 * 2];
    __shared__ int wsum[kSelNT / 32];
    __shared__ float sm_stat[4];
    constexpr int NW = kSelNT / 32;
    __shared__ int cnts[kSelMaxRounds * NW + 32];   // per (round, warp) selected counts -> exclusive offsets; [R * NW] = total
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- segmentation threshold in logit space.  test.py:70 tests sigmoid(logit) > seg_thresh per pixel; the sigmoid evaluated here
    //      (sel_sigmoid, which reproduces torch's mask bit for bit on the reference fixtures) is monotone, so the smallest float x*
    //      with sel_sigmoid(x*) > seg_thresh is found once per sample by bisection over the ordered float bit patterns (one thread,
    //      32 steps) and every pixel then costs one comparison instead of an exponential and a division. ----
    __shared__ float s_xstar;
    if (tid == 0) {
        auto key2f = [](unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k); };   // increasing key -> increasing float
        const unsigned kmin = ~0xFF800000u /* -inf */, kmax = 0x7F800000u ^ 0x80000000u /* +inf */;
        float xs;
        if (!(sel_sigmoid(key2f(kmax)) > d.seg_thresh)) xs = NAN;              // nothing passes (comparisons with NaN are false)
        else if (sel_sigmoid(key2f(kmin)) > d.seg_thresh) xs = -INFINITY;       // everything (finite or infinite) passes
        else {
            unsigned lo = kmin, hi = kmax;                                      // f(lo) fails, f(hi) passes
            while (hi - lo > 1u) {
                const unsigned mid = lo + ((hi - lo) >> 1);
                if (sel_sigmoid(key2f(mid)) > d.seg_thresh) hi = mid; else lo = mid;
            }
            xs = key2f(hi);
        }
        s_xstar = xs;
    }
    // ---- softmax statistics (test.py:84-88): joint over 2*H*W for a (B,1,1,1) scale, per channel for (B,2,1,1) ----
    const bool fused = d.weights.ptr == nullptr;
    const float* lg = fused ? static_cast<const float*>(d.logits.ptr) + b * d.logits.stride[0] : nullptr;
    const int64_t lgc = fused ? d.logits.stride[1] : 0;
    const float* ml = static_cast<const float*>(d.msk_logits.ptr) + b * d.msk_logits.stride[0];
    // both logit planes are contiguous (lc_abi.cu checks): 16-byte loads when the planes are 16-byte aligned
    const bool v4 = fused && (HW & 3) == 0 && (reinterpret_cast<uintptr_t>(lg) & 15) == 0 && (lgc & 3) == 0;
    const bool ml_lin = d.sample == 1 && d.msk_logits.stride[2] == 1 && d.msk_logits.stride[1] == d.W;
    // every pixel sampled, joint softmax, everything contiguous: the sum pass also produces the per-pixel operands (below)
    const bool merged = v4 && ml_lin && d.scale_dim == 1 && (reinterpret_cast<uintptr_t>(ml) & 15) == 0;
    float m0 = 0.f, m1 = 0.f, k0 = 1.f, k1 = 1.f;
    int cnt = 0;
    if (fused) {
        float mx0 = -INFINITY, mx1 = -INFINITY;
        if (v4) {
            const float4* p0 = reinterpret_cast<const float4*>(lg);
            const float4* p1 = reinterpret_cast<const float4*>(lg + lgc);
            for (int j = tid; j < (HW >> 2); j += kSelNT) {
                const float4 a = __ldg(p0 + j), c = __ldg(p1 + j);
                mx0 = fmaxf(mx0, fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)));
                mx1 = fmaxf(mx1, fmaxf(fmaxf(c.x, c.y), fmaxf(c.z, c.w)));
            }
        } else {
            for (int j = tid; j < HW; j += kSelNT) { mx0 = fmaxf(mx0, lg[j]); mx1 = fmaxf(mx1, lg[lgc + j]); }
        }
        for (int o = 16; o > 0; o >>= 1) { mx0 = fmaxf(mx0, __shfl_xor_sync(kFull, mx0, o)); mx1 = fmaxf(mx1, __shfl_xor_sync(kFull, mx1, o)); }
        if (lane == 0) { redf[warp * 2] = mx0; redf[warp * 2 + 1] = mx1; }
        __syncthreads();
        if (tid < 32) {
            float a0 = lane < kSelNT / 32 ? redf[lane * 2] : -INFINITY, a1 = lane < kSelNT / 32 ? redf[lane * 2 + 1] : -INFINITY;
            for (int o = 16; o > 0; o >>= 1) { a0 = fmaxf(a0, __shfl_xor_sync(kFull, a0, o)); a1 = fmaxf(a1, __shfl_xor_sync(kFull, a1, o)); }
            if (d.scale_dim == 1) a0 = a1 = fmaxf(a0, a1);
            if (lane == 0) { sm_stat[0] = a0; sm_stat[1] = a1; }
        }
        __syncthreads();
        m0 = sm_stat[0]; m1 = sm_stat[1];
        float s0 = 0.f, s1 = 0.f;
        if (merged) {
            // one pass: exponentials once per pixel -> their sums AND the quantile operand e0 + e1 (UNNORMALISED: the order statistic,
            // the interpolated threshold and every comparison are invariant to the common positive factor scale / Z up to the
            // rounding the fused mode tolerates anyway) and the mask flag, four pixels per thread and trip
            const float xstar = s_xstar;
            const float4* p0 = reinterpret_cast<const float4*>(lg);
            const float4* p1 = reinterpret_cast<const float4*>(lg + lgc);
            const float4* pm = reinterpret_cast<const float4*>(ml);
            const bool qim = d.mode == LC_SEL_QUANTILE_IN_MASK;
            for (int j = tid; j < (HW >> 2); j += kSelNT) {
                const float4 a = __ldg(p0 + j), c = __ldg(p1 + j), q = __ldg(pm + j);
                const float e0x = expf(a.x - m0), e0y = expf(a.y - m0), e0z = expf(a.z - m0), e0w = expf(a.w - m0);
                const float e1x = expf(c.x - m1), e1y = expf(c.y - m1), e1z = expf(c.z - m1), e1w = expf(c.w - m1);
                s0 += (e0x + e0y) + (e0z + e0w);
                s1 += (e1x + e1y) + (e1z + e1w);
                const bool mx = q.x >= xstar, my = q.y >= xstar, mz = q.z >= xstar, mw = q.w >= xstar;
                float4 v;
                v.x = (!qim || mx) ? __fadd_rn(e0x, e1x) : 0.f; v.y = (!qim || my) ? __fadd_rn(e0y, e1y) : 0.f;
                v.z = (!qim || mz) ? __fadd_rn(e0z, e1z) : 0.f; v.w = (!qim || mw) ? __fadd_rn(e0w, e1w) : 0.f;
                *reinterpret_cast<float4*>(vals + 4 * j) = v;
                *reinterpret_cast<unsigned*>(mflag + 4 * j) = (mx ? 1u : 0u) | (my ? 0x100u : 0u) | (mz ? 0x10000u : 0u) | (mw ? 0x1000000u : 0u);
                cnt += (mx ? 1 : 0) + (my ? 1 : 0) + (mz ? 1 : 0) + (mw ? 1 : 0);
            }
        } else if (v4) {
            const float4* p0 = reinterpret_cast<const float4*>(lg);
            const float4* p1 = reinterpret_cast<const float4*>(lg + lgc);
            for (int j = tid; j < (HW >> 2); j += kSelNT) {
                const float4 a = __ldg(p0 + j), c = __ldg(p1 + j);
                s0 += (expf(a.x - m0) + expf(a.y - m0)) + (expf(a.z - m0) + expf(a.w - m0));
                s1 += (expf(c.x - m1) + expf(c.y - m1)) + (expf(c.z - m1) + expf(c.w - m1));
            }
        } else {
            for (int j = tid; j < HW; j += kSelNT) { s0 += expf(lg[j] - m0); s1 += expf(lg[lgc + j] - m1); }
        }
        for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(kFull, s0, o); s1 += __shfl_xor_sync(kFull, s1, o); }
        __syncthreads();
        if (lane == 0) { redf[warp * 2] = s0; redf[warp * 2 + 1] = s1; }
        __syncthreads();
        if (tid == 0) {
            double z0 = 0.0, z1 = 0.0;
            for (int w = 0; w < kSelNT / 32; ++w) { z0 += redf[w * 2]; z1 += redf[w * 2 + 1]; }
            if (d.scale_dim == 1) z0 = z1 = z0 + z1;
            const float sc0 = ldf(d.weights_scale, b * d.weights_scale.stride[0]);
            const float sc1 = d.scale_dim == 1 ? sc0 : ldf(d.weights_scale, b * d.weights_scale.stride[0] + d.weights_scale.stride[1]);
            sm_stat[2] = static_cast<float>(sc0 / z0); sm_stat[3] = static_cast<float>(sc1 / z1);
        }
        __syncthreads();
        k0 = sm_stat[2]; k1 = sm_stat[3];
    } else {
        __syncthreads();   // s_xstar
    }
    const float xstar = s_xstar;
    const float* wp = fused ? nullptr : static_cast<const float*>(d.weights.ptr) + b * d.weights.stride[0];
    const int64_t wpc = fused ? 0 : d.weights.stride[1], wpy = fused ? 0 : d.weights.stride[2], wpx = fused ? 0 : d.weights.stride[3];
    auto inv_std_at = [&](int y, int x, float& w0, float& w1) {
        if (fused) { const int p = y * d.W + x; w0 = expf(lg[p] - m0) * k0; w1 = expf(lg[lgc + p] - m1) * k1; }
        else { const int64_t o = y * wpy + x * wpx; w0 = wp[o]; w1 = wp[o + wpc]; }
    };

    // ---- per sampled pixel: segmentation flag and the quantile operand ----
    if (!merged) {
#pragma unroll 4
        for (int i = tid; i < N; i += kSelNT) {
            const int yq = i / Wn, xq = i - yq * Wn, y = yq * d.sample, x = xq * d.sample;
            const bool m = (ml_lin ? ml[i] : ml[y * d.msk_logits.stride[1] + x * d.msk_logits.stride[2]]) >= xstar;   // test.py:70
            float w0, w1;
            inv_std_at(y, x, w0, w1);
            const float mf = m ? 1.f : 0.f;
            // quantile_msk: weights = den_inv_std2d.sum(-1)  (of inv_std * mask for 'quantile_in_mask', test.py:104)
            vals[i] = d.mode == LC_SEL_QUANTILE_IN_MASK ? __fadd_rn(__fmul_rn(w0, mf), __fmul_rn(w1, mf)) : __fadd_rn(w0, w1);
            mflag[i] = m ? 1 : 0;
            cnt += m ? 1 : 0;
        }
    }
    cnt = __reduce_add_sync(kFull, cnt);
    if (lane == 0) wsum[warp] = cnt;
    __syncthreads();
    int n_mask = 0;
    for (int w = 0; w < kSelNT / 32; ++w) n_mask += wsum[w];
    __syncthreads();

    // ---- threshold = torch.quantile(vals, q) in fp32 (rank = q (N-1), lerp between the two neighbouring order statistics) ----
    float thr = 0.f;
    if (d.mode != LC_SEL_MASK) {
        float q = d.quantile;
        if (d.mode == LC_SEL_QUANTILE_IN_MASK) {
            const float vis_ratio = __fdiv_rn(static_cast<float>(n_mask), static_cast<float>(N));   // seg_valid_mask.float().mean(-1)
            q = __fsub_rn(1.f, __fmul_rn(d.one_minus_quantile, vis_ratio));                         // test.py:103
        }
        const float rank = __fmul_rn(q, static_cast<float>(N - 1));
        const float lo = floorf(rank), w = __fsub_rn(rank, lo);
        int klo = static_cast<int>(lo);
        klo = max(0, min(klo, N - 1));
        unsigned ua, ub;
        radix_select_pair(reinterpret_cast<const unsigned*>(vals), N, klo, hist, bc, ua, ub);
        const float a = __uint_as_float(ua), bb = (ceilf(rank) == lo) ? a : __uint_as_float(ub);
        const float df = __fsub_rn(bb, a);
        thr = w < 0.5f ? __fadd_rn(a, __fmul_rn(w, df)) : __fsub_rn(bb, __fmul_rn(df, __fsub_rn(1.f, w)));   // at::lerp
    }

    // ---- ordered compaction (v.nonzero(), test.py:106) into zero-padded outputs (cer_solver.py:67-87) ----
    const float* xz = static_cast<const float*>(d.xyz.ptr) + b * d.xyz.stride[0];
    float n3[3] = {1.f, 1.f, 1.f};
    if (d.noc_scale.ptr)
        for (int k = 0; k < 3; ++k) n3[k] = ldf(d.noc_scale, b * d.noc_scale.stride[0] + k * d.noc_scale.stride[1]);
    const bool take_all = N <= d.min_points;   // select_valid: `t[...] if len(t) > min_cnt else t`
    auto keep = [&](int i) {
        if (i >= N) return false;
        if (take_all) return true;
        if (d.mode == LC_SEL_MASK) return mflag[i] != 0;
        if (d.mode == LC_SEL_QUANTILE) return vals[i] >= thr;
        return (vals[i] >= thr) && mflag[i] != 0;
    };
    auto emit = [&](int slot, int i) {
        const int yq = i / Wn, xq = i - yq * Wn, y = yq * d.sample, x = xq * d.sample;
        float w0, w1;
        inv_std_at(y, x, w0, w1);
        const int64_t ox = y * d.xyz.stride[1] + x * d.xyz.stride[2];
        const int64_t o3 = b * d.pts3d.stride[0] + slot * d.pts3d.stride[1], o2 = b * d.pts2d.stride[0] + slot * d.pts2d.stride[1],
                      oc = b * d.inv_cov.stride[0] + slot * d.inv_cov.stride[1];
        for (int k = 0; k < 3; ++k) stf(d.pts3d, o3 + k * d.pts3d.stride[2], xz[ox + k * d.xyz.stride[3]] * n3[k]);
        stf(d.pts2d, o2, static_cast<float>(x)); stf(d.pts2d, o2 + d.pts2d.stride[2], static_cast<float>(y));
        stf(d.inv_cov, oc, w0 * w0); stf(d.inv_cov, oc + d.inv_cov.stride[2], w1 * w1);
        if (d.index) d.index[static_cast<int64_t>(b) * d.Nmax + slot] = i;
    };
    // Compaction in four steps, three barriers: (1) every thread evaluates its pixels once, keeps the verdicts as a bit mask in
    // registers and the counts of all (round, warp) cells go to shared memory; (2) one warp scans the cells; (3) the indices of
    // the selected pixels are written, in order, into the operand array (nobody reads the operands any more); (4) a DENSE loop
    // over the output slots gathers and writes: every lane busy, consecutive lanes -> consecutive output rows (the selection keeps
    // about a quarter of the pixels, so emitting from the pixel loop ran with a quarter of the lanes and a load latency per round).
    const int R = (N + kSelNT - 1) / kSelNT;
    unsigned kept[(kSelMaxRounds + 31) / 32];
#pragma unroll
    for (int w = 0; w < (kSelMaxRounds + 31) / 32; ++w) kept[w] = 0u;
#pragma unroll
    for (int w = 0; w < (kSelMaxRounds + 31) / 32; ++w) {
        for (int r = 32 * w; r < min(R, 32 * w + 32); ++r) {
            const bool v = keep(r * kSelNT + tid);
            const unsigned bal = __ballot_sync(kFull, v);
            if (v) kept[w] |= 1u << (r & 31);
            if (lane == 0) cnts[r * NW + warp] = __popc(bal);
        }
    }
    __syncthreads();
    if (tid < 32) {
        int carry = 0;
        const int cells = R * NW;
        for (int c0 = 0; c0 < cells; c0 += 32) {
            const int v = c0 + lane < cells ? cnts[c0 + lane] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += t; }
            if (c0 + lane < cells) cnts[c0 + lane] = carry + incl - v;
            carry += __shfl_sync(kFull, incl, 31);
        }
        if (lane == 0) cnts[cells] = carry;
    }
    __syncthreads();
    int* sel = reinterpret_cast<int*>(vals);   // slot -> pixel index (slot <= index: the list fits where the operands were)
#pragma unroll
    for (int w = 0; w < (kSelMaxRounds + 31) / 32; ++w) {
        for (int r = 32 * w; r < min(R, 32 * w + 32); ++r) {
            const bool v = (kept[w] >> (r & 31)) & 1u;
            const unsigned bal = __ballot_sync(kFull, v);
            if (v) sel[cnts[r * NW + warp] + __popc(bal & ((1u << lane) - 1u))] = r * kSelNT + tid;
        }
    }
    __syncthreads();
    {
        const int nsel = cnts[R * NW];
#pragma unroll 2
        for (int slot = tid; slot < nsel; slot += kSelNT) emit(slot, sel[slot]);
    }
    int total = cnts[R * NW];
    // fewer than min_points selected: pad with indices drawn from all N points (test.py:108-113 uses np.random.choice; here a
    // per-sample LCG — same distribution, different stream)
    if (!take_all && total < d.min_points) {
        if (tid < d.min_points - total) {
            unsigned h = 1664525u * static_cast<unsigned>(b * 31 + tid + 1) + 1013904223u;
            h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
            emit(total + tid, static_cast<int>(h % static_cast<unsigned>(N)));
        }
        total = d.min_points;
    }
    if (tid == 0) d.n_points[b] = total;
    // zero padding up to Nmax: 16-byte stores when an output is contiguous in (slot, component) — the usual (B,Nmax,C) tensors —
    // else element by element through the strides
    auto zero_out = [&](const lc_view& v, int C) {
        float* base = static_cast<float*>(v.ptr) + b * v.stride[0];
        if (v.stride[2] == 1 && v.stride[1] == C) { zero_flat(base, static_cast<int64_t>(total) * C, static_cast<int64_t>(d.Nmax) * C, tid); return; }
        for (int slot = total + tid; slot < d.Nmax; slot += kSelNT)
            for (int k = 0; k < C; ++k) base[slot * v.stride[1] + k * v.stride[2]] = 0.f;
    };
    zero_out(d.pts3d, 3); zero_out(d.pts2d, 2); zero_out(d.inv_cov, 2);
    if (d.index)
        for (int slot = total + tid; slot < d.Nmax; slot += kSelNT) d.index[static_cast<int64_t>(b) * d.Nmax + slot] = -1;
}

// cudaError_t as int, or -1 when the sampled point count does not fit in shared memory
int launch_select(const lc_select_args& d, cudaStream_t st) {
    const int Hn = (d.H + d.sample - 1) / d.sample, Wn = (d.W + d.sample - 1) / d.sample, N = Hn * Wn;
    const size_t smem = static_cast<size_t>(N) * 5 + 16;
    int dev = 0, max_smem = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
        return static_cast<int>(cudaGetLastError());
    if (smem + 16384 > static_cast<size_t>(max_smem) || N > kSelMaxRounds * kSelNT) return -1;
    static std::atomic<bool> configured[64];
    if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        const cudaError_t e = cudaFuncSetAttribute(lc_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 16384);
        if (e != cudaSuccess) return static_cast<int>(e);
        if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    lc_select_kernel<<<d.B, kSelNT, smem, st>>>(d);
    note_kernel("lc::lc_select_kernel");
    return static_cast<int>(cudaGetLastError());
}

}  // namespace lc
