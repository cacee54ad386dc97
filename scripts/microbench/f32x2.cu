// Microbenchmark: scalar FP32 vs packed f32x2 (Blackwell add/mul/fma.f32x2) throughput per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}

template <int MODE>   // 0 scalar FFMA, 1 packed fma.f32x2, 2 scalar FADD, 3 packed add.f32x2
__global__ void k(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    if (MODE == 0 || MODE == 2) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = (MODE == 0) ? fmaf(x[i], a, b) : (x[i] + a);
        }
    } else {
        unsigned long long p[8], pa = pk(a, a), pb = pk(b, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = pk(x[2 * i], x[2 * i + 1]);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
                else asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) upk(p[i], x[2 * i], x[2 * i + 1]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int warps_per_sm) {
    int sms = 148;
    int block = 128, grid = sms * warps_per_sm * 32 / block;
    float* out;
    cudaMalloc(&out, (size_t)grid * block * 4);
    int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, block>>>(out, 100, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double elem_ops = (double)grid * block * iters * 16;   // per-element operations
    printf("%-18s warps/SM=%2d  %.3f ms  %.1f G elem-ops/s  (%.2f elem-ops/clk/SM @1.965GHz)\n", name, warps_per_sm, ms,
           elem_ops / ms / 1e6, elem_ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}

int main() {
    for (int w : {4, 8, 16, 32}) {
        run<0>("scalar FFMA", w);
        run<1>("fma.f32x2", w);
        run<2>("scalar FADD", w);
        run<3>("add.f32x2", w);
    }
    return 0;
}
