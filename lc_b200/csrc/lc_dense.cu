// lc_b200 — dense producer fused with the LC loss (include/lc_b200.h: lc_dense_args).
//
// One CTA per sample.  Instead of materialising softmax weights, the scaled/strided pts3d, the pixel grid and their
// gradients in HBM (what Loss_fn.dense_pose_loss does around Loss_cov_mixed, losses.py:336-386), the kernel
//   1. reduces the 2*H*W weight logits of the sample to (max, sum exp)                      [joint softmax, :355]
//   2. gathers the sub-sampled pixels straight from the NCHW network outputs into shared memory
//      (X = xyz_noc * noc_scale, x = pixel coordinates; the weights are recomputed from the logits on use)  [:142-161]
//   3. runs the shared-memory resident LC phase (lc_resident.cuh)                             [cov_mixed.py:100-150]
//   4. turns the per-point gradients into d/d xyz_noc, d/d logits (softmax backward) and d/d weights_scale and
//      writes every pixel of those tensors once, coalesced.
//
// ZEBRA = true is the zebrapose branch (losses.py:163-184): step 2 decodes pts3d from the Gray-coded
// bit logits with the MSB-error soft decoding of floatbits.py:99-160 (+ noc_scale and the model transform of
// losses.py:16-45), step 4 routes d/d pts3d to the one bit channel per axis that carries a gradient.
// lc_decode_kernel is the test-time decode (floatbits.py:33-47, 197-224).
#include <atomic>

#include "lc_resident.cuh"

namespace lc {

struct DenseGeom {
    int H, W, sample, top, left, Hn, Wn;
    float inv_wn;
    // sampled point i -> pixel offset inside an (H,W) plane
    __device__ __forceinline__ int pix(int i) const {
        const int yq = __float2int_rz((static_cast<float>(i) + 0.5f) * inv_wn);   // i / Wn, exact for i < 2^20
        const int xq = i - yq * Wn;
        return (top + sample * yq) * W + left + sample * xq;
    }
};

// inv_std of sampled point i = softmax(logits)[pix] * scale, recomputed from the logits (L1/L2 resident)
// exp() of a non-positive logit difference: ex2.approx via __expf (2 instructions, ~1e-6 relative) — the same function
// is used for the normaliser, the weights and the epilogue, so the softmax stays exactly normalised to itself.
__device__ __forceinline__ float sm_exp(float x) { return __expf(x); }

// ---- ZebraPose binary code (floatbits.py) ----
struct ZebraSrc {
    const float* lg;            // bit logits of this sample, (C,H,W) with contiguous planes
    int64_t lgc;                // channel stride
    const unsigned char* raw;   // GT raw bits of this sample
    int64_t rc, ry, rx;         // channel / row / column strides
    const unsigned char* msk;   // msk_noc of this sample
    int64_t my, mx;
    int W;
    int n[3];
    float bf;                   // -1 under a black background (floatbits.py:131)
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// One axis of mod_logits2float_with_gt_bb_scripted (floatbits.py:134-160) at pixel (y, x).  Returns the decoded value;
// isel = the bit whose logit carries the gradient, dval = d val / d logit[c0 + isel] (0 outside the mask).
__device__ __forceinline__ float zebra_axis(const ZebraSrc& z, int c0, int N, int y, int x, int& isel, float& dval) {
    const int p = y * z.W + x;
    const unsigned char* rp = z.raw + y * z.ry + x * z.rx;
    const bool inm = z.msk[y * z.my + x * z.mx] != 0;
    bool prev = false, found = false;
    float corr = 0.f, outv = 0.f, lsel = 0.f, wsel = 1.f, ssel = 1.f;
    isel = N - 1;
    for (int j = 0; j < N; ++j) {
        const float l = z.lg[(c0 + j) * z.lgc + p];
        const bool g = rp[(c0 + j) * z.rc] != 0;
        float sgn = (j >= 1 && prev) ? -1.f : 1.f;            // :140
        if (j < 2) sgn *= z.bf;                               // :141
        const float lp = l * sgn;                             // :142
        const bool pred = lp > 0.f;                           // :146
        const float wgt = static_cast<float>(1 << (N - 1 - j));
        if (pred) outv += wgt;                                // :147
        const bool err = (pred != g) || (j == N - 1);         // :149-150
        if (err && !found) { found = true; isel = j; lsel = lp; wsel = wgt; ssel = sgn; }   // :152-154 (bit idx dropped from the GT sum)
        else if (g) corr += wgt;                              // :156
        prev = g;
    }
    const float sg = sigmoidf_(lsel);
    dval = inm ? sg * (1.f - sg) * wsel * ssel : 0.f;
    return inm ? fmaf(sg, wsel, corr) : outv;                 // :157-158
}

struct SoftmaxWeights {
    const float* l0;   // logits plane a = 0 of this sample
    const float* l1;   // plane a = 1
    DenseGeom g;
    float m, k;        // max logit, scale / sum exp
    __device__ __forceinline__ void get(int i, float& s0, float& s1) const {
        const int p = g.pix(i);
        s0 = sm_exp(l0[p] - m) * k;
        s1 = sm_exp(l1[p] - m) * k;
    }
};

// per-point gradients parked in shared memory (over q and ec, which the owning thread has already consumed)
struct DenseSink {
    ResLayout l;
    float sgw;   // per-thread partial of sum_k gbar_k * w_k
    bool want;
    __device__ __forceinline__ bool want_any() const { return want; }
    __device__ __forceinline__ bool want_pts3d() const { return true; }
    __device__ __forceinline__ void weight_grad(int i, int c, float g, float w) {
        (c ? l.B1 : l.B0)[i] = g;
        sgw = fmaf(g, w, sgw);
    }
    __device__ __forceinline__ void pts2d_grad(int, int, float) const {}
    __device__ __forceinline__ void pts3d_grad(int i, float g0, float g1, float g2) const { l.A0[i] = g0; l.A1[i] = g1; l.A2[i] = g2; }
};

template <int NT>
__device__ __forceinline__ float block_max(float v, double* red, double* fin) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = static_cast<float>(red[0]);
        for (int w = 1; w < NT / 32; ++w) m = fmaxf(m, static_cast<float>(red[w]));
        fin[0] = m;
    }
    __syncthreads();
    const float m = static_cast<float>(fin[0]);
    __syncthreads();
    return m;
}

template <int NT, bool ZEBRA>
__global__ void __launch_bounds__(NT, 512 / NT) lc_dense_kernel(const lc_dense_args d, const lc_args a, int npad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PoseShared& s = *reinterpret_cast<PoseShared*>(smem_raw);
    const ResLayout l = res_layout(smem_raw, npad);
    const int b = blockIdx.x, tid = threadIdx.x;
    DenseGeom g;
    g.H = d.H; g.W = d.W; g.sample = d.sample; g.top = d.top; g.left = d.left;
    g.Hn = (d.H - d.top + d.sample - 1) / d.sample;
    g.Wn = (d.W - d.left + d.sample - 1) / d.sample;
    g.inv_wn = 1.0f / static_cast<float>(g.Wn);
    const int n = g.Hn * g.Wn;
    const int HW = d.H * d.W;

    // ---- pose constants ----
    if (tid < 9) s.K[tid] = ldf(d.K, b * d.K.stride[0] + (tid / 3) * d.K.stride[1] + (tid % 3) * d.K.stride[2]);
    else if (tid < 16) s.pose[tid - 9] = ldf(d.pose, b * d.pose.stride[0] + (tid - 9) * d.pose.stride[1]);
    for (int k = tid; k < 24; k += NT)
        s.bbox[k] = ldf(d.bbox, b * d.bbox.stride[0] + (k / 3) * d.bbox.stride[1] + (k % 3) * d.bbox.stride[2]);

    // ---- joint softmax statistics over the 2*H*W logits of this sample (losses.py:355) ----
    const float* lg = static_cast<const float*>(d.logits.ptr) + b * d.logits.stride[0];
    const int64_t lgc = d.logits.stride[1];
    float mx = -INFINITY;
    for (int j = tid; j < 2 * HW; j += NT) mx = fmaxf(mx, lg[(j >= HW ? lgc : 0) + (j >= HW ? j - HW : j)]);
    const float m = block_max<NT>(mx, s.red, s.fin);
    double se[1] = {0.0};
    {
        float acc = 0.f;
        for (int j = tid; j < 2 * HW; j += NT) acc += sm_exp(lg[(j >= HW ? lgc : 0) + (j >= HW ? j - HW : j)] - m);
        se[0] = acc;
    }
    block_reduce<1, NT>(se, s.red, s.fin);
    const double Z = s.fin[0];
    const float scale = ldf(d.weights_scale, b * d.weights_scale.stride[0]);
    const float kk = static_cast<float>(static_cast<double>(scale) / Z);
    __syncthreads();

    // ---- gather the sub-sampled correspondences (losses.py:142-161 / 163-184) ----
    const float n0 = ldf(d.noc_scale, b * d.noc_scale.stride[0]), n1 = ldf(d.noc_scale, b * d.noc_scale.stride[0] + d.noc_scale.stride[1]),
                n2 = ldf(d.noc_scale, b * d.noc_scale.stride[0] + 2 * d.noc_scale.stride[1]);
    ZebraSrc z{};
    float Tm[12] = {1.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f};   // rows a of [T[:3,:3] | T[:3,3]]
    float ihalf[3] = {0.f, 0.f, 0.f};
    if (ZEBRA) {
        z.lg = static_cast<const float*>(d.noc_bin_logits.ptr) + b * d.noc_bin_logits.stride[0];
        z.lgc = d.noc_bin_logits.stride[1];
        z.raw = static_cast<const unsigned char*>(d.noc_bin_raw.ptr) + b * d.noc_bin_raw.stride[0];
        z.rc = d.noc_bin_raw.stride[1]; z.ry = d.noc_bin_raw.stride[2]; z.rx = d.noc_bin_raw.stride[3];
        z.msk = static_cast<const unsigned char*>(d.msk_noc.ptr) + b * d.msk_noc.stride[0];
        z.my = d.msk_noc.stride[1]; z.mx = d.msk_noc.stride[2];
        z.W = d.W;
        z.bf = d.black_background ? -1.f : 1.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) { z.n[k] = d.bit_cnt[k]; ihalf[k] = 2.f / static_cast<float>((1 << d.bit_cnt[k]) - 1); }
        if (d.model_transform.ptr) {
#pragma unroll
            for (int k = 0; k < 12; ++k)
                Tm[k] = ldf(d.model_transform, b * d.model_transform.stride[0] + (k / 4) * d.model_transform.stride[1] + (k % 4) * d.model_transform.stride[2]);
        }
    }
    {
        const float* xyz = ZEBRA ? nullptr : static_cast<const float*>(d.xyz_noc.ptr) + b * d.xyz_noc.stride[0];
        const int64_t xc = ZEBRA ? 0 : d.xyz_noc.stride[1];
        for (int i = tid; i < npad; i += NT) {
            if (i < n) {
                const int p = g.pix(i);
                const int y = p / d.W, x = p - y * d.W;
                if (ZEBRA) {
                    int isel; float dv;
                    // noc = val / (max_val / 2) - 1 (floatbits.py:112), xyz_xformed = noc * noc_scale (losses.py:42)
                    const float f0 = fmaf(zebra_axis(z, 0, z.n[0], y, x, isel, dv), ihalf[0], -1.f) * n0 - Tm[3];
                    const float f1 = fmaf(zebra_axis(z, z.n[0], z.n[1], y, x, isel, dv), ihalf[1], -1.f) * n1 - Tm[7];
                    const float f2 = fmaf(zebra_axis(z, z.n[0] + z.n[1], z.n[2], y, x, isel, dv), ihalf[2], -1.f) * n2 - Tm[11];
                    // xyz = (xyz_xformed - T[:3,3]) @ T[:3,:3]  (losses.py:44)
                    l.A0[i] = fmaf(f0, Tm[0], fmaf(f1, Tm[4], f2 * Tm[8]));
                    l.A1[i] = fmaf(f0, Tm[1], fmaf(f1, Tm[5], f2 * Tm[9]));
                    l.A2[i] = fmaf(f0, Tm[2], fmaf(f1, Tm[6], f2 * Tm[10]));
                } else {
                    l.A0[i] = xyz[p] * n0; l.A1[i] = xyz[xc + p] * n1; l.A2[i] = xyz[2 * xc + p] * n2;
                }
                l.B0[i] = static_cast<float>(x); l.B1[i] = static_cast<float>(y);
            } else {
                l.A0[i] = 0.f; l.A1[i] = 0.f; l.A2[i] = 0.f; l.B0[i] = 0.f; l.B1[i] = 0.f;
            }
        }
    }
    __syncthreads();

    // ---- LC loss forward + per-point gradients ----
    const SoftmaxWeights wsrc{lg, lg + lgc, g, m, kk};
    DenseSink sink{l, 0.f, d.g_xyz_noc.ptr || d.g_logits.ptr || d.g_scale.ptr || (ZEBRA && d.g_noc_bin.ptr)};
    lc_phase_res<NT>(a, s, l, b, n, wsrc, sink, XAcc<false>{l, 0u});
    if (!sink.want) return;

    // ---- epilogue: softmax backward and the scatter, every pixel written once ----
    double sg[1] = {sink.sgw};
    block_reduce<1, NT>(sg, s.red, s.fin);
    // S = sum_k gbar_k p_k = (sum_k gbar_k w_k) / scale;   d/d logit_j = w_j (gbar_j - S);   d/d scale = S
    const float S = static_cast<float>(s.fin[0] / static_cast<double>(scale));
    if (tid == 0 && d.g_scale.ptr) stf(d.g_scale, b * d.g_scale.stride[0], S);
    float* gl = d.g_logits.ptr ? static_cast<float*>(d.g_logits.ptr) + b * d.g_logits.stride[0] : nullptr;
    float* gx = (!ZEBRA && d.g_xyz_noc.ptr) ? static_cast<float*>(d.g_xyz_noc.ptr) + b * d.g_xyz_noc.stride[0] : nullptr;
    float* gb = (ZEBRA && d.g_noc_bin.ptr) ? static_cast<float*>(d.g_noc_bin.ptr) + b * d.g_noc_bin.stride[0] : nullptr;
    const int64_t glc = d.g_logits.stride[1], gxc = d.g_xyz_noc.stride[1], gbc = d.g_noc_bin.stride[1];
    for (int p = tid; p < HW; p += NT) {
        const int y = p / d.W, x = p - y * d.W;
        const int dy = y - d.top, dx = x - d.left;
        const bool sampled = dy >= 0 && dx >= 0 && (dy % d.sample) == 0 && (dx % d.sample) == 0;
        const int i = sampled ? (dy / d.sample) * g.Wn + dx / d.sample : 0;
        if (gl) {
            const float w0 = sm_exp(lg[p] - m) * kk, w1 = sm_exp(lg[lgc + p] - m) * kk;
            gl[p] = w0 * ((sampled ? l.B0[i] : 0.f) - S);
            gl[glc + p] = w1 * ((sampled ? l.B1[i] : 0.f) - S);
        }
        if (gx) {
            gx[p] = sampled ? l.A0[i] * n0 : 0.f;
            gx[gxc + p] = sampled ? l.A1[i] * n1 : 0.f;
            gx[2 * gxc + p] = sampled ? l.A2[i] * n2 : 0.f;
        }
        if (ZEBRA && gb) {
            // d/d xyz_xformed_a = sum_k gX_k T[a][k]; one bit channel per axis carries it (floatbits.py:157)
            const float g0 = sampled ? l.A0[i] : 0.f, g1 = sampled ? l.A1[i] : 0.f, g2 = sampled ? l.A2[i] : 0.f;
            int c0 = 0;
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                const int N = z.n[ax];
                int isel = -1;
                float gsel = 0.f;
                if (sampled) {
                    float dv;
                    zebra_axis(z, c0, N, y, x, isel, dv);
                    const float gxf = fmaf(g0, Tm[ax * 4], fmaf(g1, Tm[ax * 4 + 1], g2 * Tm[ax * 4 + 2]));
                    gsel = gxf * (ax == 0 ? n0 : (ax == 1 ? n1 : n2)) * ihalf[ax] * dv;
                }
                for (int j = 0; j < N; ++j) gb[(c0 + j) * gbc + p] = (j == isel) ? gsel : 0.f;
                c0 += N;
            }
        }
    }
}

// Test-time decode: one thread per pixel.  mod_logits2float_bb (floatbits.py:197-224): hard Gray bits (leading two
// inverted under a black background) -> binary by a running XOR, LSB replaced by sigmoid(l_last * (1 - (val & 2))).
__global__ void lc_decode_kernel(const lc_decode_args d) {
    const int HW = d.H * d.W;
    const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gid >= static_cast<int64_t>(d.B) * HW) return;
    const int b = static_cast<int>(gid / HW), p = static_cast<int>(gid - static_cast<int64_t>(b) * HW);
    const float* lg = static_cast<const float*>(d.noc_bin_logits.ptr) + b * d.noc_bin_logits.stride[0] + p;
    const int64_t lgc = d.noc_bin_logits.stride[1];
    float xf[3];
    int c0 = 0;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        const int N = d.bit_cnt[ax];
        int val = 0, bin = 0;
        float last = 0.f;
        for (int j = 0; j < N; ++j) {
            last = lg[(c0 + j) * lgc];
            int bit = last > 0.f ? 1 : 0;
            if (d.black_background && j < 2) bit ^= 1;
            bin ^= bit;                       // binary bit j = XOR of the Gray bits 0..j
            val = (val << 1) | bin;
        }
        const float lsb_factor = static_cast<float>(1 - (val & 2));
        const float v = static_cast<float>(val & ~1) + sigmoidf_(last * lsb_factor);
        const float noc = v / (static_cast<float>((1 << N) - 1) * 0.5f) - 1.f;
        xf[ax] = noc * ldf(d.noc_scale, b * d.noc_scale.stride[0] + ax * d.noc_scale.stride[1]);
        c0 += N;
    }
    float o0 = xf[0], o1 = xf[1], o2 = xf[2];
    if (d.model_transform.ptr) {
        float T[12];
#pragma unroll
        for (int k = 0; k < 12; ++k)
            T[k] = ldf(d.model_transform, b * d.model_transform.stride[0] + (k / 4) * d.model_transform.stride[1] + (k % 4) * d.model_transform.stride[2]);
        const float f0 = xf[0] - T[3], f1 = xf[1] - T[7], f2 = xf[2] - T[11];
        o0 = fmaf(f0, T[0], fmaf(f1, T[4], f2 * T[8]));
        o1 = fmaf(f0, T[1], fmaf(f1, T[5], f2 * T[9]));
        o2 = fmaf(f0, T[2], fmaf(f1, T[6], f2 * T[10]));
    }
    const int y = p / d.W, x = p - y * d.W;
    const int64_t o = b * d.xyz.stride[0] + y * d.xyz.stride[1] + x * d.xyz.stride[2];
    stf(d.xyz, o, o0); stf(d.xyz, o + d.xyz.stride[3], o1); stf(d.xyz, o + 2 * d.xyz.stride[3], o2);
}

// Same decode, four consecutive pixels per thread: one 16-byte load per bit plane and thread (4x the bytes in flight of the
// scalar kernel at the same instruction count) and three 16-byte stores of the 4 x 3 output floats.  Needs H*W % 4 == 0,
// 16-byte aligned logit planes and a contiguous, 16-byte aligned (B,H,W,3) output.
__global__ void __launch_bounds__(256) lc_decode_kernel_v4(const lc_decode_args d) {
    const int HW = d.H * d.W, Q = HW / 4;
    const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gid >= static_cast<int64_t>(d.B) * Q) return;
    const int b = static_cast<int>(gid / Q), p = static_cast<int>(gid - static_cast<int64_t>(b) * Q) * 4;
    const float* lg = static_cast<const float*>(d.noc_bin_logits.ptr) + b * d.noc_bin_logits.stride[0] + p;
    const int64_t lgc = d.noc_bin_logits.stride[1];
    float xf[3][4];
    int c0 = 0;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        const int N = d.bit_cnt[ax];
        int val[4] = {0, 0, 0, 0}, bin[4] = {0, 0, 0, 0};
        float4 last = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int j = 0; j < N; ++j) {
            last = __ldcs(reinterpret_cast<const float4*>(lg + (c0 + j) * lgc));   // streamed once: evict-first
            const float lv[4] = {last.x, last.y, last.z, last.w};
            const int flip = (d.black_background && j < 2) ? 1 : 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                bin[k] ^= (lv[k] > 0.f ? 1 : 0) ^ flip;
                val[k] = (val[k] << 1) | bin[k];
            }
        }
        const float lv[4] = {last.x, last.y, last.z, last.w};
        const float ihalf = 1.f / (static_cast<float>((1 << N) - 1) * 0.5f);
        const float ns = ldf(d.noc_scale, b * d.noc_scale.stride[0] + ax * d.noc_scale.stride[1]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float v = static_cast<float>(val[k] & ~1) + sigmoidf_(lv[k] * static_cast<float>(1 - (val[k] & 2)));
            xf[ax][k] = (v / (static_cast<float>((1 << N) - 1) * 0.5f) - 1.f) * ns;
        }
        (void)ihalf;
        c0 += N;
    }
    float o[12];
    if (d.model_transform.ptr) {
        float T[12];
#pragma unroll
        for (int k = 0; k < 12; ++k)
            T[k] = ldf(d.model_transform, b * d.model_transform.stride[0] + (k / 4) * d.model_transform.stride[1] + (k % 4) * d.model_transform.stride[2]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float f0 = xf[0][k] - T[3], f1 = xf[1][k] - T[7], f2 = xf[2][k] - T[11];
            o[3 * k] = fmaf(f0, T[0], fmaf(f1, T[4], f2 * T[8]));
            o[3 * k + 1] = fmaf(f0, T[1], fmaf(f1, T[5], f2 * T[9]));
            o[3 * k + 2] = fmaf(f0, T[2], fmaf(f1, T[6], f2 * T[10]));
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) { o[3 * k] = xf[0][k]; o[3 * k + 1] = xf[1][k]; o[3 * k + 2] = xf[2][k]; }
    }
    float4* out = reinterpret_cast<float4*>(static_cast<float*>(d.xyz.ptr) + b * d.xyz.stride[0] + static_cast<int64_t>(p) * 3);
    __stcs(out, make_float4(o[0], o[1], o[2], o[3]));
    __stcs(out + 1, make_float4(o[4], o[5], o[6], o[7]));
    __stcs(out + 2, make_float4(o[8], o[9], o[10], o[11]));
}

// Target coding (floatbits.py:76-97): one thread per pixel, C bytes per output array, channel-last.
__global__ void __launch_bounds__(256) lc_encode_kernel(const lc_encode_args d) {
    const int HW = d.H * d.W;
    const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gid >= static_cast<int64_t>(d.B) * HW) return;
    const int b = static_cast<int>(gid / HW), p = static_cast<int>(gid - static_cast<int64_t>(b) * HW);
    const int y = p / d.W, x = p - y * d.W;
    const int C_ = d.bit_cnt[0] + d.bit_cnt[1] + d.bit_cnt[2];
    const int64_t o = b * d.noc.stride[0] + y * d.noc.stride[1] + x * d.noc.stride[2];
    unsigned char* mo = d.mod_bits ? d.mod_bits + gid * C_ : nullptr;
    unsigned char* ro = d.raw_bits ? d.raw_bits + gid * C_ : nullptr;
    int c0 = 0;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        const int N = d.bit_cnt[ax];
        const float mx = static_cast<float>((1 << N) - 1);
        // ints = torch.clamp_((numbers + 1) * (max_num * 0.5), 0, max_num).round_().to(int32)   (floatbits.py:87-88)
        const float v = __fmul_rn(__fadd_rn(ldf(d.noc, o + ax * d.noc.stride[3]), 1.f), mx * 0.5f);
        const int iv = static_cast<int>(rintf(fminf(fmaxf(v, 0.f), mx)));     // NaN -> 0 like clamp's NaN propagation + int cast on CPU
        int prev = 0;
        for (int j = 0; j < N; ++j) {
            const int bit = (iv >> (N - 1 - j)) & 1;
            int mod = bit ^ prev;                                              // floatbits.py:91-92
            if (d.black_background && j < 2) mod ^= 1;                         // :93-94
            if (ro) ro[c0 + j] = static_cast<unsigned char>(bit);
            if (mo) mo[c0 + j] = static_cast<unsigned char>(mod);
            prev = bit;
        }
        c0 += N;
    }
}

int launch_encode(const lc_encode_args& d, cudaStream_t st) {
    const int64_t total = static_cast<int64_t>(d.B) * d.H * d.W;
    if (total == 0) return 0;
    lc_encode_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(d);
    note_kernel("lc::lc_encode_kernel");
    return static_cast<int>(cudaGetLastError());
}

int launch_decode(const lc_decode_args& d, cudaStream_t st) {
    const int64_t HW = static_cast<int64_t>(d.H) * d.W, total = static_cast<int64_t>(d.B) * HW;
    if (total == 0) return 0;
    const lc_view& l = d.noc_bin_logits;
    const lc_view& o = d.xyz;
    const bool vec = HW % 4 == 0 && reinterpret_cast<uintptr_t>(l.ptr) % 16 == 0 && l.stride[0] % 4 == 0 && l.stride[1] % 4 == 0 &&
                     reinterpret_cast<uintptr_t>(o.ptr) % 16 == 0 && o.stride[3] == 1 && o.stride[2] == 3 && o.stride[1] == 3 * d.W &&
                     o.stride[0] % 4 == 0;
    if (vec) lc_decode_kernel_v4<<<static_cast<unsigned>((total / 4 + 255) / 256), 256, 0, st>>>(d);
    else lc_decode_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(d);
    note_kernel(vec ? "lc::lc_decode_kernel_v4" : "lc::lc_decode_kernel");
    return static_cast<int>(cudaGetLastError());
}

template <int NT, bool ZEBRA>
static int launch_dense_t(const lc_dense_args& d, const lc_args& a, int n, int max_smem, cudaStream_t st) {
    const size_t smem = resident_smem_bytes(n);
    static std::atomic<bool> configured[64];   // per instantiation and per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        const cudaError_t e = cudaFuncSetAttribute(lc_dense_kernel<NT, ZEBRA>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e != cudaSuccess) return static_cast<int>(e);
        if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    lc_dense_kernel<NT, ZEBRA><<<d.B, NT, smem, st>>>(d, a, round_up4(n));
    note_kernel("lc::lc_dense_kernel<%d,%s>", NT, ZEBRA ? "zebra" : "xyz");
    return static_cast<int>(cudaGetLastError());
}

// returns cudaError_t as int, or -1 when the sampled point count does not fit in shared memory
int launch_dense(const lc_dense_args& d, cudaStream_t st) {
    const int Hn = (d.H - d.top + d.sample - 1) / d.sample, Wn = (d.W - d.left + d.sample - 1) / d.sample;
    const int n = Hn * Wn;
    int dev = 0, max_smem = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) {
        return static_cast<int>(cudaGetLastError());
    }
    if (resident_smem_bytes(n) > static_cast<size_t>(max_smem)) return -1;
    // the LC phase reads its scalars and outputs through an lc_args
    lc_args a{};
    a.abi_version = LC_B200_ABI_VERSION; a.B = d.B; a.N = n; a.dtype = LC_F32;
    a.max_err_len = d.max_err_len; a.rel_thresh = d.rel_thresh; a.w_e_thresh = d.w_e_thresh; a.grad_scale = d.grad_scale;
    a.K = d.K; a.pose = d.pose; a.bbox = d.bbox; a.grad_out = d.grad_out; a.loss = d.loss; a.cov = d.cov; a.update_cov = d.update_cov;
    a.lc_flags = d.lc_flags; a.loss_sum = d.loss_sum;
    // the zebrapose epilogue writes C x H x W gradients per sample: 256 threads as soon as the map is large
    if (d.noc_bin_logits.ptr)
        return (n <= 2048 && d.H * d.W < 8192) ? launch_dense_t<128, true>(d, a, n, max_smem, st) : launch_dense_t<256, true>(d, a, n, max_smem, st);
    return n <= 2048 ? launch_dense_t<128, false>(d, a, n, max_smem, st) : launch_dense_t<256, false>(d, a, n, max_smem, st);
}

}  // namespace lc
