// Probe: tensor memory (TMEM) as per-thread scratch storage.  Each thread of a 128-thread CTA keeps 32 x 4 floats in its own
// TMEM lane (128 columns per CTA, 4 CTAs per SM = all 512 columns), written with tcgen05.st.32x32b.x4 and read back with
// tcgen05.ld.32x32b.x4.  Checks the values and reports cycles per load.   nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)),
                 "r"(__float_as_uint(c)), "r"(__float_as_uint(d)) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t addr, float& a, float& b, float& c, float& d) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    a = __uint_as_float(r0); b = __uint_as_float(r1); c = __uint_as_float(r2); d = __uint_as_float(r3);
}

__global__ void __launch_bounds__(128, 4) probe(float* out, long long* cyc, int reps, int* live, int* peak) {
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned smid;
    asm("mov.u32 %0, %smid;" : "=r"(smid));
    if (tid == 0) { const int c = atomicAdd(&live[smid], 1) + 1; atomicMax(&peak[smid], c); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)), "n"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tbase + ((uint32_t)(warp & 3) * 32u << 16);
    for (int k = 0; k < 32; ++k) tmem_st4(base + 4 * k, tid + 1000.f * k, blockIdx.x, k, -1.f);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    float acc = 0.f;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
        for (int k = 0; k < 32; ++k) {
            float a, b, c, d;
            tmem_ld4(base + 4 * k, a, b, c, d);
            acc += a + c;
            if (r == 0 && (a != tid + 1000.f * k || b != (float)blockIdx.x || c != (float)k || d != -1.f)) acc = NAN;
        }
    const long long t1 = clock64();
    out[blockIdx.x * 128 + tid] = acc;
    if (tid == 0) { cyc[blockIdx.x] = t1 - t0; atomicAdd(&live[smid], -1); }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(128));
}

int main() {
    const int blocks = 148 * 8, reps = 50;
    float* out; long long* cyc;
    cudaMalloc(&out, blocks * 128 * sizeof(float)); cudaMalloc(&cyc, blocks * sizeof(long long));
    int *live, *peak; cudaMalloc(&live, 256 * sizeof(int)); cudaMalloc(&peak, 256 * sizeof(int)); cudaMemset(live, 0, 1024); cudaMemset(peak, 0, 1024);
    probe<<<blocks, 128>>>(out, cyc, reps, live, peak);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    float* h = new float[blocks * 128]; long long* hc = new long long[blocks];
    cudaMemcpy(h, out, blocks * 128 * sizeof(float), cudaMemcpyDeviceToHost); cudaMemcpy(hc, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < blocks * 128; ++i) if (h[i] != h[i]) ++bad;
    double c = 0; for (int i = 0; i < blocks; ++i) c += hc[i];
    printf("mismatching threads: %d of %d;  cycles per tcgen05.ld.x4 + wait (4 CTAs/SM): %.1f\n", bad, blocks * 128, c / blocks / (reps * 32.0));
    int hp[256]; cudaMemcpy(hp, peak, 1024, cudaMemcpyDeviceToHost); int mx = 0; for (int i = 0; i < 256; ++i) mx = hp[i] > mx ? hp[i] : mx;
    printf("peak concurrent CTAs on one SM (measured): %d\n", mx);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe, 128, 0); printf("occupancy: %d CTAs/SM\n", occ);
    return bad != 0;
}
