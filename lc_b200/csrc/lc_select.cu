// lc_b200 — test-time point selection on the device (include/lc_b200.h: lc_select_args).
//
// Replaces the selection half of test.solve_pnp_dense (test.py:67-119): joint (or per-channel) softmax of the weight logits
// times the scale (:84-88), the stride-`sample` sub-sampling of the pixel grid / weights / xyz / segmentation mask with
// top_left = (0,0) (losses.py:142-161), inv_cov = inv_std^2 (:95), the 'mask' / 'quantile' / 'quantile_in_mask' rules
// (:97-104, quantile_msk :36-45 = torch.quantile with linear interpolation) and `nonzero()` + ragged python lists (:106-119)
// -> ONE launch, one CTA per sample, that writes zero-padded (B,Nmax,.) correspondences in selection order plus n_points:
// exactly what cer_solver._batch_tensors (cer_solver.py:67-87) would build, without the device->host sync of nonzero().
//
// The per-sample quantile is an exact order statistic: 4-pass 8-bit radix select over the fp32 bit patterns held in
// shared memory (values are >= 0, so the bit patterns are ordered), followed by torch's fp32 lerp.
#include <atomic>

#include "lc_resident.cuh"

namespace lc {

constexpr int kSelNT = 1024;

__device__ __forceinline__ float sel_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// value of the rank-k (0-based) smallest key among u[0..n) and of its successor in sorted order; all threads call
__device__ void radix_select_pair(const unsigned* u, int n, int k, unsigned* hist, unsigned* bc, unsigned& kth, unsigned& next) {
    const int tid = threadIdx.x;
    unsigned prefix = 0, mask = 0;
    int kk = k;
    for (int pass = 3; pass >= 0; --pass) {
        for (int j = tid; j < 256; j += kSelNT) hist[j] = 0;
        __syncthreads();
        const int sh = 8 * pass;
        for (int i = tid; i < n; i += kSelNT)
            if ((u[i] & mask) == prefix) atomicAdd(&hist[(u[i] >> sh) & 255u], 1u);
        __syncthreads();
        if (tid == 0) {
            unsigned cum = 0;
            int bin = 0;
            for (; bin < 256; ++bin) {
                if (cum + hist[bin] > static_cast<unsigned>(kk)) break;
                cum += hist[bin];
            }
            bc[0] = static_cast<unsigned>(bin);
            bc[1] = cum;
        }
        __syncthreads();
        prefix |= bc[0] << sh;
        mask |= 0xFFu << sh;
        kk -= static_cast<int>(bc[1]);
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
    atomicAdd(&bc[0], cle);
    atomicMin(&bc[1], mn);
    __syncthreads();
    next = (bc[0] >= static_cast<unsigned>(k) + 2u || bc[1] == 0xFFFFFFFFu) ? kth : bc[1];
    __syncthreads();
}

__global__ void __launch_bounds__(kSelNT) lc_select_kernel(const lc_select_args d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Hn = (d.H + d.sample - 1) / d.sample, Wn = (d.W + d.sample - 1) / d.sample, N = Hn * Wn, HW = d.H * d.W;
    float* vals = reinterpret_cast<float*>(smem_raw);                    // [N] quantile operand
    unsigned char* mflag = reinterpret_cast<unsigned char*>(vals + N);   // [N] segmentation mask of the sampled pixel
    __shared__ unsigned hist[256];
    __shared__ unsigned bc[4];
    __shared__ float redf[kSelNT / 32 * 2];
    __shared__ int wsum[kSelNT / 32];
    __shared__ float sm_stat[4];
    __shared__ int s_base;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- softmax statistics (test.py:84-88): joint over 2*H*W for a (B,1,1,1) scale, per channel for (B,2,1,1) ----
    const bool fused = d.weights.ptr == nullptr;
    const float* lg = fused ? static_cast<const float*>(d.logits.ptr) + b * d.logits.stride[0] : nullptr;
    const int64_t lgc = fused ? d.logits.stride[1] : 0;
    float m0 = 0.f, m1 = 0.f, k0 = 1.f, k1 = 1.f;
    if (fused) {
        float mx0 = -INFINITY, mx1 = -INFINITY;
        for (int j = tid; j < HW; j += kSelNT) { mx0 = fmaxf(mx0, lg[j]); mx1 = fmaxf(mx1, lg[lgc + j]); }
        for (int o = 16; o > 0; o >>= 1) { mx0 = fmaxf(mx0, __shfl_xor_sync(kFull, mx0, o)); mx1 = fmaxf(mx1, __shfl_xor_sync(kFull, mx1, o)); }
        if (lane == 0) { redf[warp * 2] = mx0; redf[warp * 2 + 1] = mx1; }
        __syncthreads();
        if (tid == 0) {
            float a0 = redf[0], a1 = redf[1];
            for (int w = 1; w < kSelNT / 32; ++w) { a0 = fmaxf(a0, redf[w * 2]); a1 = fmaxf(a1, redf[w * 2 + 1]); }
            if (d.scale_dim == 1) a0 = a1 = fmaxf(a0, a1);
            sm_stat[0] = a0; sm_stat[1] = a1;
        }
        __syncthreads();
        m0 = sm_stat[0]; m1 = sm_stat[1];
        float s0 = 0.f, s1 = 0.f;
        for (int j = tid; j < HW; j += kSelNT) { s0 += expf(lg[j] - m0); s1 += expf(lg[lgc + j] - m1); }
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
    }
    const float* wp = fused ? nullptr : static_cast<const float*>(d.weights.ptr) + b * d.weights.stride[0];
    const int64_t wpc = fused ? 0 : d.weights.stride[1], wpy = fused ? 0 : d.weights.stride[2], wpx = fused ? 0 : d.weights.stride[3];
    auto inv_std_at = [&](int y, int x, float& w0, float& w1) {
        if (fused) { const int p = y * d.W + x; w0 = expf(lg[p] - m0) * k0; w1 = expf(lg[lgc + p] - m1) * k1; }
        else { const int64_t o = y * wpy + x * wpx; w0 = wp[o]; w1 = wp[o + wpc]; }
    };
    const float* ml = static_cast<const float*>(d.msk_logits.ptr) + b * d.msk_logits.stride[0];

    // ---- per sampled pixel: segmentation flag and the quantile operand ----
    int cnt = 0;
    for (int i = tid; i < N; i += kSelNT) {
        const int yq = i / Wn, xq = i - yq * Wn, y = yq * d.sample, x = xq * d.sample;
        const bool m = sel_sigmoid(ml[y * d.msk_logits.stride[1] + x * d.msk_logits.stride[2]]) > d.seg_thresh;   // test.py:70
        float w0, w1;
        inv_std_at(y, x, w0, w1);
        const float mf = m ? 1.f : 0.f;
        // quantile_msk: weights = den_inv_std2d.sum(-1)  (of inv_std * mask for 'quantile_in_mask', test.py:104)
        vals[i] = d.mode == LC_SEL_QUANTILE_IN_MASK ? __fadd_rn(__fmul_rn(w0, mf), __fmul_rn(w1, mf)) : __fadd_rn(w0, w1);
        mflag[i] = m ? 1 : 0;
        cnt += m ? 1 : 0;
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
    if (tid == 0) s_base = 0;
    __syncthreads();
    const bool take_all = N <= d.min_points;   // select_valid: `t[...] if len(t) > min_cnt else t`
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
    for (int i0 = 0; i0 < N; i0 += kSelNT) {
        const int i = i0 + tid;
        bool v = false;
        if (i < N) {
            if (take_all) v = true;
            else if (d.mode == LC_SEL_MASK) v = mflag[i] != 0;
            else if (d.mode == LC_SEL_QUANTILE) v = vals[i] >= thr;
            else v = (vals[i] >= thr) && mflag[i] != 0;
        }
        const unsigned bal = __ballot_sync(kFull, v);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += wsum[w];
        if (v) emit(off + __popc(bal & ((1u << lane) - 1u)), i);
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < kSelNT / 32; ++w) t += wsum[w]; s_base += t; }
        __syncthreads();
    }
    int total = s_base;
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
    // zero padding up to Nmax
    for (int slot = total + tid; slot < d.Nmax; slot += kSelNT) {
        const int64_t o3 = b * d.pts3d.stride[0] + slot * d.pts3d.stride[1], o2 = b * d.pts2d.stride[0] + slot * d.pts2d.stride[1],
                      oc = b * d.inv_cov.stride[0] + slot * d.inv_cov.stride[1];
        for (int k = 0; k < 3; ++k) stf(d.pts3d, o3 + k * d.pts3d.stride[2], 0.f);
        for (int k = 0; k < 2; ++k) { stf(d.pts2d, o2 + k * d.pts2d.stride[2], 0.f); stf(d.inv_cov, oc + k * d.inv_cov.stride[2], 0.f); }
        if (d.index) d.index[static_cast<int64_t>(b) * d.Nmax + slot] = -1;
    }
}

// cudaError_t as int, or -1 when the sampled point count does not fit in shared memory
int launch_select(const lc_select_args& d, cudaStream_t st) {
    const int Hn = (d.H + d.sample - 1) / d.sample, Wn = (d.W + d.sample - 1) / d.sample, N = Hn * Wn;
    const size_t smem = static_cast<size_t>(N) * 5 + 16;
    int dev = 0, max_smem = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
        return static_cast<int>(cudaGetLastError());
    if (smem + 4096 > static_cast<size_t>(max_smem)) return -1;
    static std::atomic<bool> configured[64];
    if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        const cudaError_t e = cudaFuncSetAttribute(lc_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 4096);
        if (e != cudaSuccess) return static_cast<int>(e);
        if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    lc_select_kernel<<<d.B, kSelNT, smem, st>>>(d);
    note_kernel("lc::lc_select_kernel");
    return static_cast<int>(cudaGetLastError());
}

}  // namespace lc
