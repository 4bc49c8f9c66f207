// Microbenchmark: cost of warp-level "peers with the same digit" for radix ranking on sm_100a.
// variants: 0 = none (load + store only), 1 = match.any, 2 = NBITS ballots, 3 = shared atomicOr match
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
constexpr int THREADS = 256, ITEMS = 16, TILE = THREADS * ITEMS;

template <int MODE, int NBITS>
__global__ void __launch_bounds__(THREADS) k(const unsigned* __restrict__ keys, unsigned* __restrict__ out, int N) {
    __shared__ unsigned masks[THREADS / 32][2][1 << NBITS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long base = (long long)blockIdx.x * TILE + warp * 32 * ITEMS + lane;
    unsigned key[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) key[i] = (base + i * 32 < N) ? keys[base + i * 32] : 0u;
    if (MODE == 3) {
        for (int i = lane; i < 2 * (1 << NBITS); i += 32) (&masks[warp][0][0])[i] = 0;
        __syncwarp();
    }
    unsigned acc = 0;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const unsigned d = key[i] & ((1u << NBITS) - 1u);
        unsigned peers;
        if (MODE == 0) peers = d;
        if (MODE == 1) peers = __match_any_sync(0xffffffffu, d);
        if (MODE == 2) {
            peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < NBITS; b++) {
                const bool bit = (d >> b) & 1u;
                const unsigned bal = __ballot_sync(0xffffffffu, bit);
                peers &= bit ? bal : ~bal;
            }
        }
        if (MODE == 3) {
            unsigned* m = masks[warp][i & 1];
            atomicOr(&m[d], 1u << lane);
            __syncwarp();
            peers = m[d];
            // clear the other buffer's entry of the previous item (safe: everyone has read it)
            if (i > 0) masks[warp][(i - 1) & 1][key[i - 1] & ((1u << NBITS) - 1u)] = 0;
        }
        acc += __popc(peers & lt) + (__ffs(peers) - 1);
    }
    out[blockIdx.x * THREADS + tid] = acc;
}

template <int MODE, int NBITS>
float run(const unsigned* keys, unsigned* out, int N) {
    const int T = (N + TILE - 1) / TILE;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) k<MODE, NBITS><<<T, THREADS>>>(keys, out, N);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; i++) k<MODE, NBITS><<<T, THREADS>>>(keys, out, N);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms * 100.f;  // us per launch
}

int main() {
    const int N = 11000000;
    unsigned* h = (unsigned*)malloc(N * 4);
    unsigned *d, *o;
    cudaMalloc(&d, N * 4); cudaMalloc(&o, N * 4);
    for (int pat = 0; pat < 2; pat++) {
        // pat 0: random digits; pat 1: runs of consecutive ids (tile ids as emitted)
        unsigned x = 12345u, cur = 0, left = 0;
        for (int i = 0; i < N; i++) {
            x = x * 1664525u + 1013904223u;
            if (pat == 0) h[i] = x >> 8;
            else { if (!left) { cur = (x >> 8) % 8000; left = 1 + (x >> 28) % 6; } h[i] = cur++; left--; }
        }
        cudaMemcpy(d, h, N * 4, cudaMemcpyHostToDevice);
        printf("pattern %d: none %.1f us | match7 %.1f match8 %.1f | ballot6 %.1f ballot7 %.1f ballot8 %.1f | atomicOr7 %.1f atomicOr8 %.1f\n", pat,
               run<0, 7>(d, o, N), run<1, 7>(d, o, N), run<1, 8>(d, o, N), run<2, 6>(d, o, N), run<2, 7>(d, o, N), run<2, 8>(d, o, N),
               run<3, 7>(d, o, N), run<3, 8>(d, o, N));
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
