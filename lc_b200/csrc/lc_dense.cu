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

template <int NT>
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

    // ---- gather the sub-sampled correspondences (losses.py:142-161) ----
    {
        const float* xyz = static_cast<const float*>(d.xyz_noc.ptr) + b * d.xyz_noc.stride[0];
        const int64_t xc = d.xyz_noc.stride[1];
        const float n0 = ldf(d.noc_scale, b * d.noc_scale.stride[0]), n1 = ldf(d.noc_scale, b * d.noc_scale.stride[0] + d.noc_scale.stride[1]),
                    n2 = ldf(d.noc_scale, b * d.noc_scale.stride[0] + 2 * d.noc_scale.stride[1]);
        for (int i = tid; i < npad; i += NT) {
            if (i < n) {
                const int p = g.pix(i);
                l.A0[i] = xyz[p] * n0; l.A1[i] = xyz[xc + p] * n1; l.A2[i] = xyz[2 * xc + p] * n2;
                const int y = p / d.W, x = p - y * d.W;
                l.B0[i] = static_cast<float>(x); l.B1[i] = static_cast<float>(y);
            } else {
                l.A0[i] = 0.f; l.A1[i] = 0.f; l.A2[i] = 0.f; l.B0[i] = 0.f; l.B1[i] = 0.f;
            }
        }
    }
    __syncthreads();

    // ---- LC loss forward + per-point gradients ----
    const SoftmaxWeights wsrc{lg, lg + lgc, g, m, kk};
    DenseSink sink{l, 0.f, d.g_xyz_noc.ptr || d.g_logits.ptr || d.g_scale.ptr};
    lc_phase_res<NT>(a, s, l, b, n, wsrc, sink);
    if (!sink.want) return;

    // ---- epilogue: softmax backward and the scatter, every pixel written once ----
    double sg[1] = {sink.sgw};
    block_reduce<1, NT>(sg, s.red, s.fin);
    // S = sum_k gbar_k p_k = (sum_k gbar_k w_k) / scale;   d/d logit_j = w_j (gbar_j - S);   d/d scale = S
    const float S = static_cast<float>(s.fin[0] / static_cast<double>(scale));
    if (tid == 0 && d.g_scale.ptr) stf(d.g_scale, b * d.g_scale.stride[0], S);
    float* gl = d.g_logits.ptr ? static_cast<float*>(d.g_logits.ptr) + b * d.g_logits.stride[0] : nullptr;
    float* gx = d.g_xyz_noc.ptr ? static_cast<float*>(d.g_xyz_noc.ptr) + b * d.g_xyz_noc.stride[0] : nullptr;
    const int64_t glc = d.g_logits.stride[1], gxc = d.g_xyz_noc.stride[1];
    const float n0 = ldf(d.noc_scale, b * d.noc_scale.stride[0]), n1 = ldf(d.noc_scale, b * d.noc_scale.stride[0] + d.noc_scale.stride[1]),
                n2 = ldf(d.noc_scale, b * d.noc_scale.stride[0] + 2 * d.noc_scale.stride[1]);
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
    }
}

template <int NT>
static int launch_dense_t(const lc_dense_args& d, const lc_args& a, int n, int max_smem, cudaStream_t st) {
    const size_t smem = resident_smem_bytes(n);
    static bool configured[64] = {};   // per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        const cudaError_t e = cudaFuncSetAttribute(lc_dense_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e != cudaSuccess) return static_cast<int>(e);
        configured[dev] = true;
    }
    lc_dense_kernel<NT><<<d.B, NT, smem, st>>>(d, a, round_up4(n));
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
    return n <= 2048 ? launch_dense_t<128>(d, a, n, max_smem, st) : launch_dense_t<256>(d, a, n, max_smem, st);
}

}  // namespace lc
