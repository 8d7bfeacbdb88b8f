// Micro-benchmark: per-SM issue throughput of the instruction classes the LC kernels lean on.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/pipe_ubench tools/pipe_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096
template <int OP>
__global__ void k(float* out, float a, double da) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    double d0 = x0, d1 = x1, d2 = x2, d3 = x3, d4 = x4, d5 = x5, d6 = x6, d7 = x7;
    float2 p0 = {x0, x1}, p1 = {x2, x3}, p2 = {x4, x5}, p3 = {x6, x7}, p4 = {x1, x0}, p5 = {x3, x2}, p6 = {x5, x4}, p7 = {x7, x6};
    const float2 a2 = {a, a};
#pragma unroll 1
    for (int i = 0; i < ITER; ++i) {
        if (OP == 0) { // FFMA
            x0 = fmaf(x0, a, a); x1 = fmaf(x1, a, a); x2 = fmaf(x2, a, a); x3 = fmaf(x3, a, a);
            x4 = fmaf(x4, a, a); x5 = fmaf(x5, a, a); x6 = fmaf(x6, a, a); x7 = fmaf(x7, a, a);
        } else if (OP == 1) { // FFMA2
            p0 = __ffma2_rn(p0, a2, a2); p1 = __ffma2_rn(p1, a2, a2); p2 = __ffma2_rn(p2, a2, a2); p3 = __ffma2_rn(p3, a2, a2);
            p4 = __ffma2_rn(p4, a2, a2); p5 = __ffma2_rn(p5, a2, a2); p6 = __ffma2_rn(p6, a2, a2); p7 = __ffma2_rn(p7, a2, a2);
        } else if (OP == 2) { // DFMA
            d0 = fma(d0, da, da); d1 = fma(d1, da, da); d2 = fma(d2, da, da); d3 = fma(d3, da, da);
            d4 = fma(d4, da, da); d5 = fma(d5, da, da); d6 = fma(d6, da, da); d7 = fma(d7, da, da);
        } else if (OP == 3) { // F2F f32->f64 (+ DADD to consume)
            d0 += (double)x0; d1 += (double)x1; d2 += (double)x2; d3 += (double)x3;
            d4 += (double)x4; d5 += (double)x5; d6 += (double)x6; d7 += (double)x7;
            x0 += a; x1 += a; x2 += a; x3 += a; x4 += a; x5 += a; x6 += a; x7 += a;
        } else if (OP == 4) { // MUFU.RCP fp32 approx
            x0 = __frcp_rn(x0) + a; x1 = __fdividef(a, x1); x2 = __fdividef(a, x2); x3 = __fdividef(a, x3);
            x4 = __fdividef(a, x4); x5 = __fdividef(a, x5); x6 = __fdividef(a, x6); x7 = __fdividef(a, x7);
        } else if (OP == 5) { // fp64 IEEE division
            d0 = da / d0; d1 = da / d1; d2 = da / d2; d3 = da / d3; d4 = da / d4; d5 = da / d5; d6 = da / d6; d7 = da / d7;
        } else if (OP == 6) { // fp64 sqrt
            d0 = sqrt(d0) + da; d1 = sqrt(d1) + da; d2 = sqrt(d2) + da; d3 = sqrt(d3) + da;
            d4 = sqrt(d4) + da; d5 = sqrt(d5) + da; d6 = sqrt(d6) + da; d7 = sqrt(d7) + da;
        } else if (OP == 7) { // F2F f64->f32
            x0 += (float)d0; x1 += (float)d1; x2 += (float)d2; x3 += (float)d3;
            d0 += da; d1 += da; d2 += da; d3 += da;
        } else if (OP == 8) { // DADD
            d0 += da; d1 += da; d2 += da; d3 += da; d4 += da; d5 += da; d6 += da; d7 += da;
        } else if (OP == 9) { // FFMA with 3 distinct regs
            x0 = fmaf(x0, x1, x2); x1 = fmaf(x1, x2, x3); x2 = fmaf(x2, x3, x4); x3 = fmaf(x3, x4, x5);
            x4 = fmaf(x4, x5, x6); x5 = fmaf(x5, x6, x7); x6 = fmaf(x6, x7, x0); x7 = fmaf(x7, x0, x1);
        } else if (OP == 10) { // DFMA 3 distinct regs
            d0 = fma(d0, d1, d2); d1 = fma(d1, d2, d3); d2 = fma(d2, d3, d4); d3 = fma(d3, d4, d5);
            d4 = fma(d4, d5, d6); d5 = fma(d5, d6, d7); d6 = fma(d6, d7, d0); d7 = fma(d7, d0, d1);
        } else if (OP == 11) { // fp32 sqrt approx + rsqrt
            x0 = rsqrtf(x0) + a; x1 = rsqrtf(x1) + a; x2 = rsqrtf(x2) + a; x3 = rsqrtf(x3) + a;
            x4 = rsqrtf(x4) + a; x5 = rsqrtf(x5) + a; x6 = rsqrtf(x6) + a; x7 = rsqrtf(x7) + a;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + (float)(d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7) +
        p0.x + p0.y + p1.x + p1.y + p2.x + p2.y + p3.x + p3.y + p4.x + p4.y + p5.x + p5.y + p6.x + p6.y + p7.x + p7.y;
}

template <int OP>
void run(const char* name, int ops_per_iter, float* out) {
    int dev; cudaGetDevice(&dev);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    const int nsm = pr.multiProcessorCount, threads = 1024, blocks = nsm * 2;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, 1.0001f, 1.0001);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 1.0001f, 1.0001);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const double total_thread_ops = (double)blocks * threads * ITER * ops_per_iter;
    const double per_sm_per_clk = total_thread_ops / (ms * 1e-3) / nsm / (clk * 1e3);
    printf("%-28s %8.3f ms  %7.1f thread-ops/clk/SM (at %d MHz nominal)  %.2f Tops/s\n", name, ms, per_sm_per_clk, clk / 1000, total_thread_ops / (ms * 1e-3) / 1e12);
}

int main() {
    float* out; cudaMalloc(&out, 1 << 24);
    run<0>("FFMA (reg,imm-like)", 8, out);
    run<9>("FFMA (3 regs)", 8, out);
    run<1>("FFMA2 (instr)", 8, out);
    run<2>("DFMA", 8, out);
    run<10>("DFMA (3 regs)", 8, out);
    run<8>("DADD", 8, out);
    run<3>("F2F f32->f64 (+DADD,FADD)", 8, out);
    run<7>("F2F f64->f32 (+FADD,DADD)", 4, out);
    run<4>("MUFU rcp fp32 approx", 8, out);
    run<11>("rsqrtf", 8, out);
    run<5>("fp64 div IEEE", 8, out);
    run<6>("fp64 sqrt IEEE (+DADD)", 8, out);
    return 0;
}
