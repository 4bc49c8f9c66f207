// Micro-benchmark: issue cost of packed FP32 (FFMA2, PTX fma.rn.f32x2) against scalar FFMA on sm_100a,
// alone and interleaved with ALU / MUFU work.  Decides whether two-pixels-per-lane blending pays.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/ffma2_bench tools/micro/ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
    float d;
    asm volatile("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ unsigned int iadd(unsigned int a, unsigned int b) {
    unsigned int d;
    asm volatile("lop3.b32 %0, %1, %2, %2, 0x96;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ float ex2(float a) {
    float d;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a));
    return d;
}

constexpr int ITERS = 4096;
// MODE 0: 8 FFMA / iter; 1: 8 FFMA2 / iter; 2: 8 FFMA + 8 LOP3; 3: 8 FFMA2 + 8 LOP3; 4: 8 FFMA2 + 16 LOP3;
// 5: 8 FFMA + 1 MUFU; 6: 8 FFMA2 + 2 MUFU; 7: 8 FFMA + 2 MUFU; 8: 16 LOP3
template <int MODE>
__global__ void k(float* out, float seed) {
    float a[8];
    unsigned long long p[8];
    unsigned int q[16];
    float m0 = seed, m1 = seed * 0.5f;
    const float b = seed * 1.0001f, c = seed * 0.001f;
    const unsigned long long bb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    const unsigned long long cc = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed + i; p[i] = ((unsigned long long)__float_as_uint(seed + i) << 32) | __float_as_uint(seed - i); }
#pragma unroll
    for (int i = 0; i < 16; i++) q[i] = threadIdx.x * 17 + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0 || MODE == 2 || MODE == 5 || MODE == 7) a[i] = fma1(a[i], b, c);
            if (MODE == 1 || MODE == 3 || MODE == 4 || MODE == 6) p[i] = fma2(p[i], bb, cc);
            if (MODE == 2 || MODE == 3 || MODE == 4 || MODE == 8) q[i] = iadd(q[i], q[(i + 1) & 15]);
            if (MODE == 4 || MODE == 8) q[i + 8] = iadd(q[i + 8], q[(i + 9) & 15]);
        }
        if (MODE == 5 || MODE == 6 || MODE == 7) m0 = ex2(m0);
        if (MODE == 6 || MODE == 7) m1 = ex2(m1);
    }
    float s = m0 + m1;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + __uint_as_float((unsigned int)p[i]) + __uint_as_float((unsigned int)(p[i] >> 32));
#pragma unroll
    for (int i = 0; i < 16; i++) s += (float)q[i];
    if (s == 12345.678f) out[0] = s;
}

template <int MODE>
void run(const char* name, int instr_per_iter) {
    float* out;
    cudaMalloc(&out, 4);
    const int blocks = 148 * 2, threads = 1024;
    k<MODE><<<blocks, threads>>>(out, 1.0f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k<MODE><<<blocks, threads>>>(out, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    const double warp_instr = (double)blocks * (threads / 32) * ITERS * instr_per_iter;
    // per SM sub-partition and cycle at 1.965 GHz (4 SMSPs x 148 SMs)
    const double per_smsp_clk = warp_instr / (ms * 1e-3) / (148.0 * 4) / 1.965e9;
    printf("%-28s %8.3f ms  %6.3f warp-instr/clk/SMSP (at 1965 MHz)\n", name, ms, per_smsp_clk);
    cudaFree(out);
}

int main() {
    run<0>("8 FFMA", 8);
    run<1>("8 FFMA2", 8);
    run<2>("8 FFMA + 8 LOP3", 16);
    run<3>("8 FFMA2 + 8 LOP3", 16);
    run<4>("8 FFMA2 + 16 LOP3", 24);
    run<5>("8 FFMA + 1 MUFU", 9);
    run<6>("8 FFMA2 + 2 MUFU", 10);
    run<7>("8 FFMA + 2 MUFU", 10);
    run<8>("16 LOP3", 16);
    return 0;
}
