// Microbenchmark: FFMA vs FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_bench tools/micro/ffma2_bench.cu && /tmp/ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int ILP>
__global__ void k_ffma(float* out, float a, float b, int iters) {
    float v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_ffma2(float* out, float a, float b, int iters) {
    u64 v[ILP];
    const u64 A = pk(a, a), B = pk(b, b);
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = pk(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fma2(v[i], A, B);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float x, y; upk(v[i], x, y); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 200000, ILP = 8;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); k_ffma<ILP><<<148 * 8, 256>>>(out, 1.0001f, 0.5f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * ILP * iters * 148.0 * 8 * 256;
        printf("FFMA : %.2f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
        cudaEventRecord(e0); k_ffma2<ILP><<<148 * 8, 256>>>(out, 1.0001f, 0.5f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA2: %.2f ms  %.1f TFLOP/s\n", ms, 2 * fl / ms / 1e9);
    }
    return 0;
}
