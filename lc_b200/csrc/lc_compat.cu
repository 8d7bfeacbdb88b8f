// lc_b200 — binary-compatible replacement of the reference's cffi boundary (lib/pnp/cxx/ext.h:1-14):
//
//   pnp_ceres_f32_omp(float** init_states, float** cam_Ks, float** pts2ds, float** pts3ds, float** icov_sqrtLs,
//                     int* ptCnts, int maxIterCnt, float function_tolerance, int printSummary,
//                     float* result_trs, int* rets, int job_count, int num_threads)
//
// is what lib/pnp/pnp_ceres.py:136-139 calls through `from ._ext import lib`.  Exporting the same symbol from
// liblc_b200.so lets the reference's UNMODIFIED pnp_ceres.py / cer_solver.py run on the sm_100a solver: only `_ext` is
// swapped (lc_b200/pnp/_ext.py).
//
// Unlike the device-pointer entry points of lc_b200.h this one takes HOST pointer tables (one ragged job per entry, as
// pnp_ceres.py:104-120 builds them), so it owns a staging path: the jobs are packed into ONE pinned host buffer in the
// planar layout the resident kernel stages by TMA (host threads = num_threads, the role OpenMP plays in ceres.cpp:161-169),
// copied to the device once, solved by one lc_b200_lm_solve launch and the three small result arrays copied back.
// Semantics kept from ceres.cpp: ptCnt < 3 -> ret = 1, tr = 1, state untouched (:84-91); ret = invalid flag and
// tr = last trust-region radius always written (:134-136); the state is written back only when valid (:137-144).
// The staging buffers (pinned host + device) are grow-only and cached per device; calls are serialised by a mutex.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/lc_b200.h"

namespace {

struct Staging {
    void* host = nullptr;
    void* dev = nullptr;
    size_t bytes = 0;
    cudaStream_t stream = nullptr;
};

std::mutex g_mutex;
Staging g_staging[64];

bool ensure(Staging& s, size_t need) {
    if (!s.stream && cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return false;
    if (need <= s.bytes) return true;
    if (s.host) cudaFreeHost(s.host);
    if (s.dev) cudaFree(s.dev);
    s.host = s.dev = nullptr;
    s.bytes = 0;
    const size_t cap = need + need / 4;
    if (cudaHostAlloc(&s.host, cap, cudaHostAllocDefault) != cudaSuccess) return false;
    if (cudaMalloc(&s.dev, cap) != cudaSuccess) { cudaFreeHost(s.host); s.host = nullptr; return false; }
    s.bytes = cap;
    return true;
}

size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace

extern "C" {

int lc_b200_lm_solve(const lc_args* a, void* stream);

// Returns 0 on success, else the failing CUDA / LC_E_* code (the reference's function is void; its cffi caller ignores
// the value).  On failure every job is reported invalid (rets[i] = 1) and no state is modified.
int pnp_ceres_f32_omp(float** init_states, float** cam_Ks, float** pts2ds, float** pts3ds, float** icov_sqrtLs, int* ptCnts,
                      int maxIterCnt, float function_tolerance, int printSummary, float* result_trs, int* rets, int job_count,
                      int num_threads) {
    if (job_count <= 0) return 0;
    const int B = job_count;
    int nmax = 0;
    for (int i = 0; i < B; ++i) nmax = std::max(nmax, ptCnts[i]);
    const int N = std::max(4, (nmax + 3) & ~3);   // planar slabs stay 16-byte aligned (TMA staging)
    const size_t n = static_cast<size_t>(N);

    // the weights are diagonal (L10 == 0 for every point, what cer_solver.py:37-38 produces from (.., N, 2) inverse
    // variances) -> two planar slabs; otherwise the full row-major 2x2 factors (ext.h ABI: element [1] ignored)
    bool diag = true;
    for (int i = 0; i < B && diag; ++i) {
        const float* L = icov_sqrtLs[i];
        for (int k = 0; k < ptCnts[i]; ++k)
            if (L[4 * k + 2] != 0.f) { diag = false; break; }
    }
    const size_t wper = diag ? 2 * n : 4 * n;

    // device / pinned layout (floats unless noted), every block 256-byte aligned:
    //   K (B,9) | state_in (B,7) | pts3d (B,3,N) planar | pts2d (B,2,N) planar | w (B,2,N) planar or (B,N,2,2) | n_points (B) i32
    //   | state_out (B,7) | radius (B) | invalid (B) i32 | iters (B) i32
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = align256(off + bytes); return o; };
    const size_t oK = take(sizeof(float) * 9 * B), oS = take(sizeof(float) * 7 * B), o3 = take(sizeof(float) * 3 * n * B),
                 o2 = take(sizeof(float) * 2 * n * B), oW = take(sizeof(float) * wper * B), oN = take(sizeof(int) * B);
    const size_t in_bytes = off;
    const size_t oSo = take(sizeof(float) * 7 * B), oR = take(sizeof(float) * B), oI = take(sizeof(int) * B), oIt = take(sizeof(int) * B);
    const size_t total = off;

    std::lock_guard<std::mutex> lock(g_mutex);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    auto fail = [&](int code) {
        for (int i = 0; i < B; ++i) { rets[i] = 1; result_trs[i] = 1.f; }
        fprintf(stderr, "lc_b200 pnp_ceres_f32_omp: failed with code %d (%s)\n", code, code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "argument error");
        return code;
    };
    if (e != cudaSuccess) return fail(static_cast<int>(e));
    if (dev < 0 || dev >= 64) return fail(LC_E_BADARG);
    Staging& s = g_staging[dev];
    if (!ensure(s, total)) return fail(static_cast<int>(cudaGetLastError()));

    char* h = static_cast<char*>(s.host);
    char* d = static_cast<char*>(s.dev);
    float* hK = reinterpret_cast<float*>(h + oK);
    float* hS = reinterpret_cast<float*>(h + oS);
    float* h3 = reinterpret_cast<float*>(h + o3);
    float* h2 = reinterpret_cast<float*>(h + o2);
    float* hW = reinterpret_cast<float*>(h + oW);
    int* hN = reinterpret_cast<int*>(h + oN);

    auto pack = [&](int lo, int hi) {
        for (int i = lo; i < hi; ++i) {
            memcpy(hK + 9 * i, cam_Ks[i], sizeof(float) * 9);
            memcpy(hS + 7 * i, init_states[i], sizeof(float) * 7);
            const int c = std::max(0, ptCnts[i]);
            hN[i] = c;
            float *x3 = h3 + 3 * n * i, *x2 = h2 + 2 * n * i, *w = hW + wper * i;
            const float *p3 = pts3ds[i], *p2 = pts2ds[i], *L = icov_sqrtLs[i];
            for (int k = 0; k < c; ++k) {
                x3[k] = p3[3 * k]; x3[n + k] = p3[3 * k + 1]; x3[2 * n + k] = p3[3 * k + 2];
                x2[k] = p2[2 * k]; x2[n + k] = p2[2 * k + 1];
            }
            if (diag) for (int k = 0; k < c; ++k) { w[k] = L[4 * k]; w[n + k] = L[4 * k + 3]; }
            else memcpy(w, L, sizeof(float) * 4 * c);
            for (size_t k = c; k < n; ++k) { x3[k] = 0.f; x3[n + k] = 0.f; x3[2 * n + k] = 0.f; x2[k] = 0.f; x2[n + k] = 0.f; }
            if (diag) for (size_t k = c; k < n; ++k) { w[k] = 0.f; w[n + k] = 0.f; }
            else memset(w + 4 * c, 0, sizeof(float) * 4 * (n - c));
        }
    };
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int nt = std::max(1, std::min<int>({num_threads, B, static_cast<int>(hw)}));
    if (nt == 1) pack(0, B);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(pack, static_cast<int>(static_cast<int64_t>(B) * t / nt), static_cast<int>(static_cast<int64_t>(B) * (t + 1) / nt));
        for (auto& t : th) t.join();
    }

    if ((e = cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, s.stream)) != cudaSuccess) return fail(static_cast<int>(e));

    lc_args a;
    memset(&a, 0, sizeof(a));
    a.abi_version = LC_B200_ABI_VERSION;
    a.B = B; a.N = N; a.dtype = LC_F32;
    a.flags = LC_FLAG_TOL_NEEDS_SUCCESS;
    a.max_iter = maxIterCnt;
    a.function_tolerance = static_cast<double>(function_tolerance);
    a.max_err_len = 32.0; a.rel_thresh = 3.0; a.w_e_thresh = 4.0; a.grad_scale = 1.0;
    auto view = [&](size_t o, int64_t s0, int64_t s1, int64_t s2, int64_t s3) {
        lc_view v; v.ptr = d + o; v.stride[0] = s0; v.stride[1] = s1; v.stride[2] = s2; v.stride[3] = s3; return v;
    };
    const int64_t ni = static_cast<int64_t>(n);
    a.K = view(oK, 9, 3, 1, 0);
    a.pose = view(oS, 7, 1, 0, 0);
    a.pts3d = view(o3, 3 * ni, 1, ni, 0);
    a.pts2d = view(o2, 2 * ni, 1, ni, 0);
    if (diag) { a.weight_mode = LC_W_INV_STD; a.weights = view(oW, 2 * ni, 1, ni, 0); }   // la = |L00|, lc = |L11|
    else { a.weight_mode = LC_W_SQRT_L; a.weights = view(oW, 4 * ni, 4, 2, 1); }
    a.n_points = reinterpret_cast<const int32_t*>(d + oN);
    a.state = view(oSo, 7, 1, 0, 0);
    a.radius = view(oR, 1, 0, 0, 0);
    a.invalid = reinterpret_cast<int32_t*>(d + oI);
    a.iters = reinterpret_cast<int32_t*>(d + oIt);
    const int rc = lc_b200_lm_solve(&a, s.stream);
    if (rc != 0) return fail(rc);
    if ((e = cudaMemcpyAsync(h + oSo, d + oSo, total - oSo, cudaMemcpyDeviceToHost, s.stream)) != cudaSuccess) return fail(static_cast<int>(e));
    if ((e = cudaStreamSynchronize(s.stream)) != cudaSuccess) return fail(static_cast<int>(e));

    const float* so = reinterpret_cast<const float*>(h + oSo);
    const float* ro = reinterpret_cast<const float*>(h + oR);
    const int* io = reinterpret_cast<const int*>(h + oI);
    const int* ito = reinterpret_cast<const int*>(h + oIt);
    for (int i = 0; i < B; ++i) {
        rets[i] = io[i] ? 1 : 0;
        result_trs[i] = ptCnts[i] < 3 ? 1.f : ro[i];
        if (!io[i]) memcpy(init_states[i], so + 7 * i, sizeof(float) * 7);
        if (printSummary)
            printf("lc_b200 job %d: %d correspondences, %d LM iterations, trust-region radius %g, %s\n", i, ptCnts[i], ito[i],
                   static_cast<double>(result_trs[i]), ptCnts[i] < 3 ? "skipped problem with less than 3 points" : (io[i] ? "NOT usable / no convergence" : "CONVERGENCE"));
    }
    return 0;
}

// single-problem form (ext.h does not declare it, ceres.cpp:72-83 exports it)
int pnp_ceres_f32(float* io_state_quat, const float* cam_K, const float* pts2d, const float* pts3d, const float* icov_sqrtL, int ptCnt,
                  int maxIterCnt, float function_tolerance, int printSummary, float* result_tr, int* ret) {
    float* st = io_state_quat;
    float* K = const_cast<float*>(cam_K);
    float* p2 = const_cast<float*>(pts2d);
    float* p3 = const_cast<float*>(pts3d);
    float* L = const_cast<float*>(icov_sqrtL);
    return pnp_ceres_f32_omp(&st, &K, &p2, &p3, &L, &ptCnt, maxIterCnt, function_tolerance, printSummary, result_tr, ret, 1, 1);
}

// release the cached staging buffers of every device (optional; they are also reclaimed at process exit)
void lc_b200_compat_release(void) {
    std::lock_guard<std::mutex> lock(g_mutex);
    for (auto& s : g_staging) {
        if (s.host) cudaFreeHost(s.host);
        if (s.dev) cudaFree(s.dev);
        if (s.stream) cudaStreamDestroy(s.stream);
        s = Staging();
    }
}

}  // extern "C"
