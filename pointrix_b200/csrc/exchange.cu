// Gradient exchange of the view-sharded data-parallel render path over NVLink 5 /
// NVSwitch: an in-switch (NVLS) all-reduce of a symmetric buffer through its multicast mapping.
//
// Every rank holds the same-sized buffer  [ n_f32 floats | n_i32 int32 ]  registered as symmetric
// memory; `mc` is the multicast virtual address that aliases all of the replicas.  Rank r owns the
// r-th 1/world slice of each section: it pulls the slice with multimem.ld_reduce (the switch reads
// every replica and returns the SUM for the float section -- the parameter gradients and ndc.grad
// written by the fused backward -- and the MAX for the int section -- radii) and pushes the result
// back with multimem.st (the switch writes it into every replica).  Per GPU and direction that moves
// about one buffer size over NVLink instead of the 2(n-1)/n sizes of a ring, and no SM does the
// arithmetic.  The caller brackets the launch with a cross-rank barrier on both sides (the symmetric
// memory handle's barrier): before = every replica has been written, after = every slice has landed.
// No spinning happens here, so this kernel cannot hang on a peer.
#include "common.cuh"
#include "pointrix_b200.h"

namespace pxb {

__device__ __forceinline__ float4 mm_ld_reduce_add_f32x4(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mm_st_f32x4(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ int mm_ld_reduce_max_s32(const int* mc) {
    int v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.max.s32 %0, [%1];" : "=r"(v) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mm_st_s32(int* mc, int v) {
    asm volatile("multimem.st.relaxed.sys.global.s32 [%0], %1;" :: "l"(mc), "r"(v) : "memory");
}

constexpr int kXThreads = 512;
constexpr int kXUnroll = 4;

__global__ void __launch_bounds__(kXThreads)
nvls_allreduce_kernel(float* __restrict__ mc_f32, long long v4_begin, long long v4_end, int* __restrict__ mc_i32,
                      long long i_begin, long long i_end) {
    const long long stride = (long long)gridDim.x * kXThreads;
    const long long t = (long long)blockIdx.x * kXThreads + threadIdx.x;
    // float section: kXUnroll independent 16-byte reductions in flight per thread
    for (long long i = v4_begin + t; i < v4_end; i += stride * kXUnroll) {
        float4 v[kXUnroll];
#pragma unroll
        for (int u = 0; u < kXUnroll; u++)
            if (i + u * stride < v4_end) v[u] = mm_ld_reduce_add_f32x4(mc_f32 + 4 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < kXUnroll; u++)
            if (i + u * stride < v4_end) mm_st_f32x4(mc_f32 + 4 * (i + u * stride), v[u]);
    }
    for (long long i = i_begin + t; i < i_end; i += stride) mm_st_s32(mc_i32 + i, mm_ld_reduce_max_s32(mc_i32 + i));
}

// Two-GPU (and no-multicast) variant over plain peer mappings: rank r owns slice r, loads it from every
// replica (its own from HBM, the others over NVLink), and stores the result into every replica.
// Moves 2(n-1)/n buffer sizes per GPU and direction: less than the multicast path for n = 2, where
// multimem traffic loops the local replica through the switch.
constexpr int kMaxPeers = 16;
struct PeerPtrs {
    float* p[kMaxPeers];
};

__global__ void __launch_bounds__(kXThreads)
p2p_allreduce_kernel(PeerPtrs peers, int rank, int world, long long n_f32, long long v4_begin, long long v4_end,
                     long long i_begin, long long i_end) {
    const long long stride = (long long)gridDim.x * kXThreads;
    const long long t = (long long)blockIdx.x * kXThreads + threadIdx.x;
    for (long long i = v4_begin + t; i < v4_end; i += stride) {
        float4 acc = __ldcg(reinterpret_cast<const float4*>(peers.p[0]) + i);
        for (int q = 1; q < world; q++) {  // fixed rank order: every replica ends up with identical bits
            const float4 v = __ldcg(reinterpret_cast<const float4*>(peers.p[q]) + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        for (int q = 0; q < world; q++) __stcg(reinterpret_cast<float4*>(peers.p[q]) + i, acc);
    }
    for (long long i = i_begin + t; i < i_end; i += stride) {
        int m = __ldcg(reinterpret_cast<const int*>(peers.p[0] + n_f32) + i);
        for (int q = 1; q < world; q++) m = max(m, __ldcg(reinterpret_cast<const int*>(peers.p[q] + n_f32) + i));
        for (int q = 0; q < world; q++) __stcg(reinterpret_cast<int*>(peers.p[q] + n_f32) + i, m);
    }
}

}  // namespace pxb

using namespace pxb;

extern "C" int pxb_p2p_allreduce(const void* const* peer_ptrs, long long n_f32, long long n_i32, int rank, int world,
                                 void* stream) {
    if (peer_ptrs == nullptr || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || n_f32 < 0 || n_i32 < 0)
        return PXB_ERR_BAD_ARG;
    if ((n_f32 % (4ll * world)) || (n_i32 % world)) return PXB_ERR_ALIGN;
    PeerPtrs pp;
    for (int q = 0; q < kMaxPeers; q++) pp.p[q] = q < world ? (float*)peer_ptrs[q] : nullptr;
    for (int q = 0; q < world; q++)
        if (pp.p[q] == nullptr || (((uintptr_t)pp.p[q]) & 15)) return PXB_ERR_ALIGN;
    const long long nv4 = n_f32 / 4, per_v4 = nv4 / world, per_i = n_i32 / world;
    if (per_v4 == 0 && per_i == 0) return 0;
    const long long work = per_v4 > per_i ? per_v4 : per_i;
    long long blocks = (work + kXThreads - 1) / kXThreads;
    if (blocks > 148 * 4) blocks = 148 * 4;
    p2p_allreduce_kernel<<<(int)blocks, kXThreads, 0, (cudaStream_t)stream>>>(pp, rank, world, n_f32, rank * per_v4,
                                                                              (rank + 1) * per_v4, rank * per_i,
                                                                              (rank + 1) * per_i);
    return (int)cudaGetLastError();
}

extern "C" int pxb_nvls_allreduce(void* mc_ptr, long long n_f32, long long n_i32, int rank, int world, void* stream) {
    if (mc_ptr == nullptr || world < 1 || rank < 0 || rank >= world || n_f32 < 0 || n_i32 < 0) return PXB_ERR_BAD_ARG;
    if ((n_f32 % (4ll * world)) || (n_i32 % world) || (((uintptr_t)mc_ptr) & 15)) return PXB_ERR_ALIGN;
    const long long nv4 = n_f32 / 4, per_v4 = nv4 / world, per_i = n_i32 / world;
    float* mc_f = (float*)mc_ptr;
    int* mc_i = (int*)(mc_f + n_f32);
    if (per_v4 == 0 && per_i == 0) return 0;
    const long long work = per_v4 > per_i ? per_v4 : per_i;
    long long blocks = (work + (long long)kXThreads * kXUnroll - 1) / ((long long)kXThreads * kXUnroll);
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    nvls_allreduce_kernel<<<(int)blocks, kXThreads, 0, (cudaStream_t)stream>>>(
        mc_f, rank * per_v4, (rank + 1) * per_v4, mc_i, rank * per_i, (rank + 1) * per_i);
    return (int)cudaGetLastError();
}
