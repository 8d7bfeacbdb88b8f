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
#include "lc_resident_kernel.cuh"

namespace lc {

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


// Which modes take the tensor-memory variant (measured on B200 at B = 1024 x N = 4096, see profiles/README.md): the loss-only
// kernel gains from four finer-grained CTAs per SM; LC_B200_TMEM = 0 | 1 | lc overrides for A/B runs.
static bool tmem_enabled(int mode) {
    if (const char* e = getenv("LC_B200_TMEM")) return e[0] != '0';
    return mode == MODE_LC;
}

// The vectorised point loops (lc_vec.cuh) need planar 16-byte aligned slabs for everything they touch in global memory with
// vector instructions: the weights (re-read from L2 every pass), the optional valid mask and the requested gradient outputs.
// pts3d / pts2d may have any layout (they are staged into planar shared memory either way).
static bool vec_ok(const lc_args& a, int mode) {
    if (getenv("LC_B200_NO_VEC")) return false;
    if ((a.N % 4) != 0 || !tma_ok(a.weights, a.N)) return false;
    if (a.weight_mode != LC_W_ICOV_DIAG && a.weight_mode != LC_W_INV_STD) return false;
    if (mode & MODE_LC) {
        if (a.g_pts3d.ptr && !tma_ok(a.g_pts3d, a.N)) return false;
        if (a.g_pts2d.ptr && !tma_ok(a.g_pts2d, a.N)) return false;
        if (a.g_weights.ptr && !tma_ok(a.g_weights, a.N)) return false;
        if (a.valid.ptr && !(a.valid.stride[1] == 1 && (reinterpret_cast<uintptr_t>(a.valid.ptr) % 16) == 0 && (a.valid.stride[0] % 4) == 0)) return false;
    }
    return true;
}

int launch_res_scalar(const lc_args& a, int mode, int nt, bool tm, cudaStream_t st, const ResLaunch& r) {
    return launch_res_any<false, 1>(a, mode, nt, tm, st, r);
}

static std::atomic<int> g_num_sms[64];
static int num_sms() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    const bool cached = dev >= 0 && dev < 64;
    if (cached && (v = g_num_sms[dev].load(std::memory_order_relaxed)) > 0) return v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cached) g_num_sms[dev].store(v, std::memory_order_relaxed);
    return v;
}

// Poses split over two-CTA clusters (lc_resident_kernel.cuh, CL = 2).  Measured on B200 at N = 4096 (profiles/README.md, "cluster
// split"): a CTA that has an SM to itself already runs a pose in 0.72x the time it takes next to a second CTA, and halving its
// points only shortens the point passes (the serial trust-region / 6x6 sections, the reductions and a cluster barrier per
// reduction stay), so
//   * B <= SMs / 2 (every half-pose CTA gets its own SM): P3 74 -> 67 us.  This is the default use;
//   * splitting the poses of the last, partly empty wave of a larger launch (B = 1024 on 296 slots: 136 poses as 272 half-pose
//     CTAs in front of 888 = 3 full waves) LOSES: 345 -> 390 us, because those 136 CTAs, alone on their SMs, take 76 us either way
//     and the second launch adds its ramp-up.  Kept reachable for A/B runs: LC_B200_SPLIT = 1 splits whenever the last wave fills
//     at most half of the slots, LC_B200_SPLIT = 0 never splits; LC_B200_PDL = 0 turns off the programmatic launch of the main
//     grid behind the cluster grid.
static int split_tail(const lc_args& a, int nt, bool vec, bool tm, int cap) {
    if (!vec || tm || cap != 0 || a.N < 512) return 0;
    const char* e = getenv("LC_B200_SPLIT");
    if (e && e[0] == '0') return 0;
    const int sms = num_sms();
    if (sms <= 0) return 0;
    if (!(e && e[0] == '1')) return (2 * a.B <= sms && a.N >= 2048) ? a.B : 0;
    const int slots = (nt <= 128 ? 4 : 2) * sms;
    if (a.B / slots > 8) return 0;   // many waves: the tail no longer matters
    const int tail = a.B % slots;
    return (tail > 0 && 2 * tail <= slots) ? tail : 0;
}

// cap > 0: ragged batch whose padded N exceeds the resident limit; only poses with n_points <= cap are processed here
int launch_resident_pose(const lc_args& a, int mode, cudaStream_t st, int cap) {
    // 256 threads x 2 CTAs per SM (20 B/point of shared memory): one CTA's 6x6 / trust-region sections and loads
    // overlap the other CTA's point passes
    // 2048 < N <= 4096: four 128-thread CTAs per SM with the model points in tensor memory (lc_resident.cuh: XAcc)
    const bool vec = vec_ok(a, mode);
    const bool tm = !vec && cap == 0 && a.N > 2048 && a.N <= kTmemMaxN && tmem_enabled(mode) && mode == MODE_LC;
    int nt = tm ? 128 : resident_threads_for(cap > 0 ? cap : a.N, mode);
    if (nt != 128 && nt != 192 && nt != 256) nt = 256;
    if (nt == 192 && mode != MODE_LM) nt = 256;
    ResLaunch r{cap, max_optin_smem(), tma_mask_for(a), 0, a.B, 1, 0};
    const int tail = split_tail(a, nt, vec, tm, cap);
    if (tail > 0) {
        const char* e = getenv("LC_B200_PDL");
        const bool pdl = !(e && e[0] == '0') && tail < a.B;
        ResLaunch rc = r;
        rc.pose_base = a.B - tail; rc.count = tail; rc.cl = 2; rc.pdl = pdl ? 1 : 0;
        const int rcode = launch_res_cluster2(a, mode, nt, st, rc);
        if (rcode != 0 || tail == a.B) return rcode;
        r.count = a.B - tail;
        r.pdl |= pdl ? 2 : 0;
    }
    return vec ? launch_res_vec(a, mode, nt, tm, st, r) : launch_res_scalar(a, mode, nt, tm, st, r);
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
