// Tile-based alpha blending, forward and backward
// (reference: msplat/msplat/src/alpha_blending.cu:16-110 forward, :112-246 backward).
//
// Data layout: blending consumes ONE packed record per Gaussian,
//     rec[g] = { u, v, A, B, C, opacity, f_0 .. f_{c-1}, pad }   (S floats, S % 4 == 0)
// so that the per-intersection gather is a single contiguous 16-byte-aligned
// segment instead of four scattered reads (uv / conic / opacity / feature row),
// and backward accumulates into a gradient record of the same shape,
//     grec[g] = { du, dv, dA, dB, dC, dop, df_0 .. }.
// The fused per-Gaussian kernels read/write these records directly; the
// operator-level API packs/unpacks them with two trivial streaming kernels.
//
// Forward: 16x16-pixel tile per CTA, warps own 8x4 pixel blocks (coherent early
// exit), Gaussians staged 256 at a time through shared memory with coalesced
// 16-byte copies, every warp skips the inner loop once all of its pixels are
// saturated, the CTA leaves when all 256 are.
// Backward: back-to-front replay with the reference's skip rules; the 6+C
// per-(pixel,Gaussian) partials are reduced across the warp with a
// value-halving butterfly (about one shuffle per value instead of five) and
// land as ONE contiguous vector atomic per (warp, Gaussian) on the gradient
// record.
#include <stdlib.h>

#include "common.cuh"
#include "pointrix_b200.h"

namespace pxb {

constexpr int kBlendThreads = 256;
constexpr int kBatch = 256;

// pixel owned by a thread: warp w covers the 8x4 block at (w&1, w>>1), lanes row-major inside it
__device__ __forceinline__ void thread_pixel(int& lx, int& ly) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    lx = ((w & 1) << 3) | (l & 7);
    ly = ((w >> 1) << 2) | (l >> 3);
}

// Which of the CTA's eight 8x4 pixel blocks can a Gaussian touch at all?  A pixel passes the
// reference's skip rules only if op*exp(power) >= 1/255, i.e. q(d) = A dx^2 + 2B dx dy + C dy^2 <= 2 ln(255 op);
// the bounding box of that ellipse (half extents sqrt(k C/det), sqrt(k A/det)), inflated by a safety
// margin that dwarfs the MUFU approximation error, is tested against each block.  Bit w of the
// result = block of warp w (thread_pixel).  Conservative: 0xFF whenever the conic is not a
// proper ellipse; never culls a pixel the exact per-pixel test would blend.
__device__ __forceinline__ unsigned int block_mask(float u, float v, float A, float B, float C, float op, float X0,
                                                   float Y0) {
    if (op < 1.0f / 255.0f) return 0u;  // alpha <= op < 1/255 for every pixel
    const float det = A * C - B * B;
    if (!(det > 0.f) || !(A > 0.f) || !(C > 0.f)) return 0xFFu;
    const float k = 2.02f * __logf(255.0f * op) + 0.02f;
    const float inv = 1.0f / det;
    const float hx = sqrtf(k * C * inv) + 0.51f, hy = sqrtf(k * A * inv) + 0.51f;
    if (!(hx == hx) || !(hy == hy)) return 0xFFu;
    const float xl = u - hx - X0, xh = u + hx - X0, yl = v - hy - Y0, yh = v + hy - Y0;  // tile-local
    const unsigned int mx = ((xl <= 7.f && xh >= 0.f) ? 1u : 0u) | ((xl <= 15.f && xh >= 8.f) ? 2u : 0u);
    unsigned int m = 0;
#pragma unroll
    for (int by = 0; by < 4; by++)
        if (yl <= (float)(4 * by + 3) && yh >= (float)(4 * by)) m |= mx << (2 * by);
    return m;
}

// Optional device counters (pxb_blend_counters): when the pointer is set, every warp adds the number of
// (8x4 block, Gaussian) candidates it executed to slot 0 (forward) / 1 (backward) as it leaves -- one
// 64-bit add per warp per launch.  bench.py derives the pair-test rate from them in an untimed pass.
__device__ unsigned long long* g_blend_counters = nullptr;
__device__ __forceinline__ void count_candidates(int slot, int n) {
    if ((threadIdx.x & 31) == 0) {
        unsigned long long* c = g_blend_counters;
        if (c != nullptr) atomicAdd(c + slot, (unsigned long long)n);
    }
}

#ifdef PXB_STATS
__device__ unsigned long long g_blend_stats[8];
#define PXB_STAT(i, v) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_blend_stats[i], (unsigned long long)(v)); } while (0)
#else
#define PXB_STAT(i, v) do { } while (0)
#endif

// cooperative gather of up to kBatch records into shared memory.
// Thread t copies float4 number t, t+256, ... of the flattened [kBatch][S/4] set.
template <int S>
__device__ __forceinline__ void stage_records(float* __restrict__ sm_rec, int* __restrict__ sm_id,
                                              const float* __restrict__ rec, const int* __restrict__ ids, int n) {
    constexpr int Q = S / 4;
#pragma unroll
    for (int it = 0; it < Q; it++) {
        const int f = it * kBlendThreads + threadIdx.x;
        const int j = f / Q, q = f - j * Q;
        if (j < n) {
            const int g = ids[j];
            if (sm_id != nullptr && q == 0) sm_id[j] = g;
            const float4 v = __ldg(reinterpret_cast<const float4*>(rec + (size_t)g * S) + q);
            reinterpret_cast<float4*>(sm_rec)[f] = v;
        }
    }
}

// shared-memory loads by 32-bit shared address (keeps the window base in one register; the generic
// form re-materialised it from SR_CgaCtaId on every access)
__device__ __forceinline__ float4 lds128(unsigned int a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
// stops the compiler from re-deriving a loop-invariant shared address inside the loop
__device__ __forceinline__ unsigned int opaque_u32(unsigned int x) {
    unsigned int y;
    asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ float2 lds64(unsigned int a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float lds32(unsigned int a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ int lds32i(unsigned int a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(unsigned int a, float x, float y) {
    asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ unsigned int lds_u16(unsigned int a) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
    return v;
}

// Per-warp candidate list of one staged batch: the batch entries whose block mask has this warp's bit
// (and, for the backward, whose list position is below `pos_limit`), in batch order, as u16 indices,
// padded to a multiple of kUnroll with the index of a never-blending sentinel record.  Returns the
// padded count.
constexpr int kUnroll = 4;
template <bool PAD>
__device__ __forceinline__ int build_candidates(const unsigned char* __restrict__ sm_mask, unsigned short* __restrict__ list,
                                                int n, int top, int pos_limit) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned int lt = (1u << lane) - 1u;
    int cnt = 0;
    for (int c0 = 0; c0 < n; c0 += 32) {
        const int jj = c0 + (int)lane;
        const bool bit = jj < n && (top - 1 - jj) < pos_limit && ((sm_mask[jj] >> warp) & 1u);
        const unsigned int m = __ballot_sync(0xffffffffu, bit);
        if (bit) list[cnt + __popc(m & lt)] = (unsigned short)jj;
        cnt += __popc(m);
    }
    const int padded = PAD ? ((cnt + kUnroll - 1) & ~(kUnroll - 1)) : cnt;
    if (PAD && (int)lane < padded - cnt) list[cnt + lane] = (unsigned short)kBatch;
    __syncwarp();
    return padded;
}

// record slot kBatch of the staging buffer: opacity 0 => alpha 0 => never blended
__device__ __forceinline__ void write_sentinel(float* sm_rec, int S) {
    if ((int)threadIdx.x < S) sm_rec[kBatch * S + threadIdx.x] = (threadIdx.x == 2 || threadIdx.x == 4) ? 1.f : 0.f;
}

// ---- bulk-copy staging (BULK = true) ----------------------------------------------------------------
// Every record of a batch is fetched by ONE cp.async.bulk (48 contiguous, 16-byte aligned bytes at C = 3)
// issued by the thread that owns the slot; the copies of batch b+1 are in flight while batch b is blended
// (two staging buffers, one mbarrier each, completion by transaction bytes).  The synchronous variant
// (BULK = false) stages with LDG.128 + STS.128 by all 256 threads and waits for them at a barrier.
__device__ __forceinline__ void mbar_init(unsigned int bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned int bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned int bar, unsigned int parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned int dst, const void* src, unsigned int bytes, unsigned int bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int CH, int S, bool BULK>
__global__ void __launch_bounds__(kBlendThreads)
blend_fwd_kernel(const float* __restrict__ rec, const int* __restrict__ idx_sorted, const int2* __restrict__ tile_range,
                 float bg, int C, int W, int H, int gx, float* __restrict__ final_T, int* __restrict__ ncontrib,
                 float* __restrict__ out) {
    pdl_wait();
    constexpr int NBUF = BULK ? 2 : 1;
    extern __shared__ __align__(16) float sm_f[];
    float* sm_rec0 = sm_f;                                                                        // [NBUF][kBatch + 1][S]
    unsigned short* sm_list = reinterpret_cast<unsigned short*>(sm_f + NBUF * (kBatch + 1) * S);  // [8 warps][kBatch + kUnroll]
    unsigned char* sm_mask = reinterpret_cast<unsigned char*>(sm_list + (kBlendThreads / 32) * (kBatch + kUnroll));
    __shared__ __align__(8) unsigned long long s_bar[2];
    // (taking the tiles heaviest-first instead of in raster order was measured: no gain, the dynamic CTA
    // scheduler already leaves a tail shorter than one tile)
    const int tile = blockIdx.y * gx + blockIdx.x;
    const int tx = blockIdx.x, ty = blockIdx.y;
    int lx, ly;
    thread_pixel(lx, ly);
    const int pxi = tx * PXB_TILE + lx, pyi = ty * PXB_TILE + ly;
    const bool inside = pxi < W && pyi < H;
    const float pxf = (float)pxi, pyf = (float)pyi;
    // a finished pixel (saturated, or outside the image) raises its alpha threshold above 0.99, the
    // largest alpha there is: every later pair test fails by itself, the hot loop carries no "done" flag
    float thr = inside ? 1.0f / 255.0f : 2.0f;
    const int2 range = tile_range[tile];
    int todo = range.y - range.x;
    float T = 1.0f;
    int last = 0;
    float F[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) F[k] = 0.f;
    const unsigned warp = threadIdx.x >> 5;
    const float X0 = (float)(tx * PXB_TILE), Y0 = (float)(ty * PXB_TILE);
    unsigned short* my_list = sm_list + warp * (kBatch + kUnroll);
    const unsigned int rec_s0 = opaque_u32((unsigned int)__cvta_generic_to_shared(sm_rec0));
    const unsigned int list_s = opaque_u32((unsigned int)__cvta_generic_to_shared(my_list));
    write_sentinel(sm_rec0, S);
    if (BULK) write_sentinel(sm_rec0 + (kBatch + 1) * S, S);
    int n_cand = 0;
    const unsigned int bar_s = (unsigned int)__cvta_generic_to_shared(&s_bar[0]);
    // BULK: thread t owns slot t of every batch; issue = one bulk copy of its record into buffer (batch & 1)
    auto issue = [&](int batch_base, int left, int buf) {
        const int n = min(kBatch, left);
        if (threadIdx.x == 0) mbar_expect_tx(bar_s + 8u * buf, (unsigned int)n * (S * 4u));
        if ((int)threadIdx.x < n) {
            const int g = idx_sorted[range.x + batch_base + threadIdx.x];
            bulk_g2s(rec_s0 + (unsigned int)(buf * (kBatch + 1) + threadIdx.x) * (S * 4u), rec + (size_t)g * S, S * 4u,
                     bar_s + 8u * buf);
        }
    };
    if (BULK) {
        if (threadIdx.x == 0) {
            mbar_init(bar_s, 1);
            mbar_init(bar_s + 8u, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (todo > 0) issue(0, todo, 0);
    }

    for (int base = 0, it = 0; todo > 0; base += kBatch, todo -= kBatch, it++) {
        const int buf = BULK ? (it & 1) : 0;
        float* sm_rec = sm_rec0 + buf * (kBatch + 1) * S;
        const unsigned int rec_s = rec_s0 + (unsigned int)(buf * (kBatch + 1)) * (S * 4u);
        // (the barrier also orders the previous batch's reads of the other buffer before its refill below)
        const bool all_done = __syncthreads_count(thr > 1.f) == kBlendThreads;
        const int n = min(kBatch, todo);
        if (BULK) {
            if (!all_done && todo > kBatch) issue(base + kBatch, todo - kBatch, buf ^ 1);
            mbar_wait(bar_s + 8u * buf, (unsigned int)((it >> 1) & 1));  // drain this buffer even when leaving
            if (all_done) break;
        } else {
            if (all_done) break;
            stage_records<S>(sm_rec, nullptr, rec, idx_sorted + range.x + base, n);
            __syncthreads();
        }
        if ((int)threadIdx.x < n) {
            const float4 r0 = *reinterpret_cast<const float4*>(sm_rec + threadIdx.x * S);
            const float2 r1 = *reinterpret_cast<const float2*>(sm_rec + threadIdx.x * S + 4);
            sm_mask[threadIdx.x] = (unsigned char)block_mask(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, X0, Y0);
        }
        __syncthreads();
        if (__all_sync(0xffffffffu, thr > 1.f)) continue;  // whole warp saturated: it only helps staging
        const int cnt = build_candidates<true>(sm_mask, my_list, n, 0x40000000, 0x7fffffff);
        PXB_STAT(5, cnt);
        const int last_base = base + 1;
        int c = 0;
        for (; c < cnt; c += kUnroll) {
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                const int j = (int)lds_u16(list_s + 2u * (unsigned)(c + u));
                const unsigned int ra = rec_s + (unsigned)j * (S * 4u);
                const float4 r0 = lds128(ra);        // u v A B
                const float4 r1 = lds128(ra + 16u);  // C op f0 f1
                const float dx = __fadd_rn(r0.x, -pxf), dy = __fadd_rn(r0.y, -pyf);
                const float power = blend_power(dx, dy, r0.z, r0.w, r1.x);
                const float alpha = fmin_nn(__fmul_rn(r1.y, blend_G(power)), 0.99f);
                // the reference's two skips (alpha_blending.cu:81-88) as one predicate
                if (power <= 0.f && alpha >= thr) {
                    const float next_T = __fmul_rn(T, __fadd_rn(-alpha, 1.0f));
                    if (next_T < 0.0001f) {
                        thr = 2.0f;  // done: this Gaussian is NOT blended
                    } else {
                        F[0] = __fmaf_rn(T, __fmul_rn(alpha, r1.z), F[0]);
                        if (CH > 1) F[1] = __fmaf_rn(T, __fmul_rn(alpha, r1.w), F[1]);
#pragma unroll
                        for (int q = 2; q < CH; q += 4) {
                            const float4 rf = lds128(ra + 4u * (6 + q));
                            F[q] = __fmaf_rn(T, __fmul_rn(alpha, rf.x), F[q]);
                            if (q + 1 < CH) F[q + 1] = __fmaf_rn(T, __fmul_rn(alpha, rf.y), F[q + 1]);
                            if (q + 2 < CH) F[q + 2] = __fmaf_rn(T, __fmul_rn(alpha, rf.z), F[q + 2]);
                            if (q + 3 < CH) F[q + 3] = __fmaf_rn(T, __fmul_rn(alpha, rf.w), F[q + 3]);
                        }
                        T = next_T;
                        last = last_base + j;  // 1-based list position of the last blended Gaussian (ncontrib)
                    }
                }
            }
            if (__all_sync(0xffffffffu, thr > 1.f)) { c += kUnroll; break; }
        }
        n_cand += min(c, cnt);
    }
    count_candidates(0, n_cand);
    if (inside) {
        const size_t pix = (size_t)pyi * W + pxi;
        final_T[pix] = T;
        ncontrib[pix] = last;
#pragma unroll
        for (int k = 0; k < CH; k++)
            if (k < C) out[(size_t)k * H * W + pix] = __fmaf_rn(T, bg, F[k]);
    }
}

// ---------------------------------------------------------------------------
// Backward.  Two phases per warp:
//   phase 1 (lane = pixel): back-to-front replay of the warp's 8x4 pixel block with the reference's
//     skip rules; for every (pixel, Gaussian) pair that was blended it produces the two scalars all
//     gradients are linear in,
//         q = G * dL/dalpha        w = alpha * T
//     and parks them in a warp-private shared-memory slot (row = Gaussian, column = pixel);
//   phase 2 (lane = Gaussian): when kSlots Gaussians are parked, lane pair (g, g+16) walks the 32
//     pixels of Gaussian g's row and accumulates the moments
//         Q0 = sum q, Q1 = sum q dx, Q2 = sum q dy, Q3 = sum q dx^2, Q4 = sum q dx dy, Q5 = sum q dy^2,
//         F_k = sum w dL/dpix_k
//     in registers, from which
//         dL/du = -op (A Q1 + B Q2)   dL/dv = -op (C Q2 + B Q1)   dL/dA = -op Q3 / 2   dL/dB = -op Q4
//         dL/dC = -op Q5 / 2          dL/dop = Q0                  dL/df_k = F_k
//     (the reference's per-pair expressions, alpha_blending.cu:206-243, summed over pixels).
// There is no cross-lane reduction tree: the reference issues 6+C scalar atomics per (pixel,
// Gaussian); this kernel issues ceil((6+C)/4) 16-byte vector atomics per (8x4 block, Gaussian) and
// spends one shuffle per value per 16 Gaussians.
// ---------------------------------------------------------------------------
constexpr int kSlots = 16;
constexpr int kSlotPitch = 33;  // float2 units: conflict-free row (phase 1) and column (phase 2) access

template <int CH, int S, int MINB>  // MINB: resident CTAs per SM the register allocation aims for
__global__ void __launch_bounds__(kBlendThreads, MINB)
blend_bwd_kernel(const float* __restrict__ rec, const int* __restrict__ idx_sorted, const int2* __restrict__ tile_range,
                 float bg, int C, int W, int H, int gx, const float* __restrict__ final_T,
                 const int* __restrict__ ncontrib, const float* __restrict__ dL_dout, float* __restrict__ grec) {
    pdl_wait();
    constexpr int CHP = (CH + 3) & ~3;  // dL/dpix row padded to float4s
    constexpr int NW = kBlendThreads / 32;
    extern __shared__ __align__(16) float sm_f[];
    float* sm_rec = sm_f;                                               // [kBatch][S]
    float* sm_dpix = sm_rec + kBatch * S;                               // [NW][32][CHP]
    float2* sm_slot = reinterpret_cast<float2*>(sm_dpix + NW * 32 * CHP);  // [NW][kSlots][kSlotPitch]
    int* sm_id = reinterpret_cast<int*>(sm_slot + NW * kSlots * kSlotPitch);  // [kBatch]
    unsigned short* sm_list = reinterpret_cast<unsigned short*>(sm_id + kBatch);  // [NW][kBatch]
    unsigned char* sm_mask = reinterpret_cast<unsigned char*>(sm_list + NW * kBatch);  // [kBatch]
    const int tile = blockIdx.y * gx + blockIdx.x;
    const int tx = blockIdx.x, ty = blockIdx.y;
    int lx, ly;
    thread_pixel(lx, ly);
    const int pxi = tx * PXB_TILE + lx, pyi = ty * PXB_TILE + ly;
    const float pxf = (float)pxi, pyf = (float)pyi;
    const bool inside = pxi < W && pyi < H;
    const size_t pix = (size_t)pyi * W + pxi;
    const int2 range = tile_range[tile];
    const int count = range.y - range.x;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const float T_final = inside ? final_T[pix] : 0.f;
    float T = T_final;
    const int last = inside ? ncontrib[pix] : 0;
    float dpix[CH];
    float bg_dot = 0.f;
    float* my_dpix = sm_dpix + (warp * 32 + lane) * CHP;
#pragma unroll
    for (int k = 0; k < CH; k++) {
        dpix[k] = (inside && k < C) ? dL_dout[(size_t)k * H * W + pix] : 0.f;
        bg_dot += bg * dpix[k];
        my_dpix[k] = dpix[k];
    }
    float Bacc = T_final * bg_dot;  // see the pair loop
    // the deepest contributor of any pixel of the CTA / of this warp bounds the work: list entries
    // at positions >= that bound were never blended by these pixels
    const int warp_last = __reduce_max_sync(0xffffffffu, last);
    __shared__ int s_cta_last;
    if (threadIdx.x == 0) s_cta_last = 0;
    __syncthreads();
    if (lane == 0) atomicMax(&s_cta_last, warp_last);
    __syncthreads();
    const int cta_last = s_cta_last;
    const float X0 = (float)(tx * PXB_TILE), Y0 = (float)(ty * PXB_TILE);

    // phase-2 role of this lane: Gaussian slot (lane & 15), pixel half (lane >> 4) = rows 2h, 2h+1 of the block
    float2* slots = sm_slot + warp * (kSlots * kSlotPitch);
    const float* wdpix = sm_dpix + warp * 32 * CHP;
    const int my_slot = lane & 15, my_half = lane >> 4;
    const float bx = X0 + (float)((warp & 1) << 3);                       // block origin (pixel centres are integers)
    const float by = Y0 + (float)(((warp >> 1) << 2) + 2 * my_half);
    int nslot = 0;   // warp-uniform
    int my_j = 0;    // batch entry parked in my_slot
    unsigned short* my_list = sm_list + warp * kBatch;
    const unsigned int rec_s = opaque_u32((unsigned int)__cvta_generic_to_shared(sm_rec));
    const unsigned int list_s = opaque_u32((unsigned int)__cvta_generic_to_shared(my_list));

    // every shared-memory access of the hot path goes through a 32-bit shared address held in a register
    // (generic pointers made the compiler rebuild the shared window base inside the loops)
    const unsigned int slot_s = opaque_u32((unsigned int)__cvta_generic_to_shared(slots));            // this warp's slots
    const unsigned int slot_w = slot_s + 8u * lane;                                                    // phase 1: my column
    const unsigned int slot_r = slot_s + 8u * (unsigned)(my_slot * kSlotPitch + 16 * my_half);         // phase 2: my half row
    const unsigned int dpix_r = opaque_u32((unsigned int)__cvta_generic_to_shared(wdpix + 16 * my_half * CHP));
    const unsigned int id_s = opaque_u32((unsigned int)__cvta_generic_to_shared(sm_id));
    unsigned int slot_cur = slot_w;  // advances by one slot row per parked Gaussian

    auto flush = [&](int n) {
        __syncwarp();
        const unsigned int ra = rec_s + (unsigned)my_j * (S * 4u);
        const float4 r0 = lds128(ra);       // u v A B
        const float2 r1 = lds64(ra + 16u);  // C op
        // moments of q over this lane's 16 pixels in block-local pixel coordinates x = i & 7, y = i >> 3
        // (compile-time constants: one FFMA each), shifted to the Gaussian's centre afterwards
        float M0 = 0.f, Mx = 0.f, My = 0.f, Mxx = 0.f, Mxy = 0.f, Myy = 0.f;
        float Fk[CH];
#pragma unroll
        for (int k = 0; k < CH; k++) Fk[k] = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const float2 qw = lds64(slot_r + 8u * i);
            const float x = (float)(i & 7), y = (float)(i >> 3);
            M0 += qw.x;
            Mx = fmaf(qw.x, x, Mx); My = fmaf(qw.x, y, My);
            Mxx = fmaf(qw.x, x * x, Mxx); Mxy = fmaf(qw.x, x * y, Mxy); Myy = fmaf(qw.x, y * y, Myy);
#pragma unroll
            for (int k4 = 0; k4 < CHP; k4 += 4) {
                const float4 d4 = lds128(dpix_r + 4u * (i * CHP + k4));
                Fk[k4] = fmaf(qw.y, d4.x, Fk[k4]);
                if (k4 + 1 < CH) Fk[k4 + 1] = fmaf(qw.y, d4.y, Fk[k4 + 1]);
                if (k4 + 2 < CH) Fk[k4 + 2] = fmaf(qw.y, d4.z, Fk[k4 + 2]);
                if (k4 + 3 < CH) Fk[k4 + 3] = fmaf(qw.y, d4.w, Fk[k4 + 3]);
            }
        }
        // dx = ub - x, dy = vb - y with (ub, vb) the centre relative to this half block's origin
        const float ub = r0.x - bx, vb = r0.y - by;
        float Q0 = M0;
        float Q1 = fmaf(ub, M0, -Mx), Q2 = fmaf(vb, M0, -My);
        float Q3 = fmaf(ub, fmaf(ub, M0, -2.f * Mx), Mxx);
        float Q4 = fmaf(ub, fmaf(vb, M0, -My), fmaf(-vb, Mx, Mxy));
        float Q5 = fmaf(vb, fmaf(vb, M0, -2.f * My), Myy);
        Q0 += __shfl_xor_sync(0xffffffffu, Q0, 16); Q1 += __shfl_xor_sync(0xffffffffu, Q1, 16);
        Q2 += __shfl_xor_sync(0xffffffffu, Q2, 16); Q3 += __shfl_xor_sync(0xffffffffu, Q3, 16);
        Q4 += __shfl_xor_sync(0xffffffffu, Q4, 16); Q5 += __shfl_xor_sync(0xffffffffu, Q5, 16);
#pragma unroll
        for (int k = 0; k < CH; k++) Fk[k] += __shfl_xor_sync(0xffffffffu, Fk[k], 16);
        if ((int)lane < n) {
            const float nop = -r1.y;
            float g[S];
            g[0] = nop * fmaf(r0.z, Q1, r0.w * Q2);
            g[1] = nop * fmaf(r1.x, Q2, r0.w * Q1);
            g[2] = 0.5f * nop * Q3;
            g[3] = nop * Q4;
            g[4] = 0.5f * nop * Q5;
            g[5] = Q0;
#pragma unroll
            for (int k = 0; k < S - 6; k++) g[6 + k] = (k < CH) ? Fk[k] : 0.f;
            float* dst = grec + (size_t)lds32i(id_s + 4u * (unsigned)my_j) * S;
#pragma unroll
            for (int q4 = 0; q4 < S; q4 += 4) {
                if (q4 + 1 >= 6 + CH) {          // one live value left in this float4
                    if (q4 < 6 + CH) atomicAdd(dst + q4, g[q4]);
                } else if (q4 < 6 + CH) {
                    atomicAdd(reinterpret_cast<float4*>(dst + q4), make_float4(g[q4], g[q4 + 1], g[q4 + 2], g[q4 + 3]));
                }
            }
        }
        slot_cur = slot_w;
        __syncwarp();
    };

    // walk the list back to front starting at the CTA's deepest contributor
    for (int top = min(count, cta_last); top > 0; top -= kBatch) {
        __syncthreads();
        const int n = min(kBatch, top);
        // entry j of the batch is list position top-1-j
        {
            constexpr int Q = S / 4;
#pragma unroll
            for (int it = 0; it < Q; it++) {
                const int f = it * kBlendThreads + threadIdx.x;
                const int j = f / Q, q = f - j * Q;
                if (j < n) {
                    const int g = idx_sorted[range.x + top - 1 - j];
                    if (q == 0) sm_id[j] = g;
                    reinterpret_cast<float4*>(sm_rec)[f] = __ldg(reinterpret_cast<const float4*>(rec + (size_t)g * S) + q);
                }
            }
        }
        __syncthreads();
        if ((int)threadIdx.x < n) {
            const float4 r0 = *reinterpret_cast<const float4*>(sm_rec + threadIdx.x * S);
            const float2 r1 = *reinterpret_cast<const float2*>(sm_rec + threadIdx.x * S + 4);
            sm_mask[threadIdx.x] = (unsigned char)block_mask(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, X0, Y0);
        }
        __syncthreads();
        // positions >= warp_last contribute to no pixel of this warp
        const int cnt = build_candidates<false>(sm_mask, my_list, n, top, warp_last);
        count_candidates(1, cnt);  // per batch: a running total would cost the kernel its 64th register
        PXB_STAT(0, cnt);
        for (int c = 0; c < cnt; c++) {
            const int j = (int)lds_u16(list_s + 2u * (unsigned)c);
            const unsigned int ra = rec_s + (unsigned)j * (S * 4u);
            const int pos = top - 1 - j;  // list position of this entry
            const float4 r0 = lds128(ra);        // u v A B
            const float4 r1 = lds128(ra + 16u);  // C op f0 f1
            const float dx = __fadd_rn(r0.x, -pxf), dy = __fadd_rn(r0.y, -pyf);
            const float power = blend_power(dx, dy, r0.z, r0.w, r1.x);
            const float G = blend_G(power);
            const float alpha = fmin_nn(__fmul_rn(r1.y, G), 0.99f);
            const bool valid = pos < last && power <= 0.f && alpha >= 1.0f / 255.0f;
#ifdef PXB_STATS
            const unsigned int vm = __ballot_sync(0xffffffffu, valid);
            if (vm == 0u) continue;
            PXB_STAT(1, 1); PXB_STAT(2, __popc(vm));
#else
            if (!__any_sync(0xffffffffu, valid)) continue;
#endif
            float qx = 0.f, qy = 0.f;
            if (valid) {
                // dL/dalpha = T * <f - accum, dpix> - T_final/(1-alpha) * bg * sum(dpix)   (alpha_blending.cu:206-222)
                // with accum the normalised colour behind this Gaussian.  Only dot products with the
                // pixel's dL/dpix enter, so the colour recurrence collapses to ONE scalar per pixel:
                //   B = T_final*bg*sum(dpix) + sum_{behind} alpha_j T_j <f_j, dpix>,
                //   dL/dalpha = T <f, dpix> - B / (1 - alpha).
                const float ra1 = __fdividef(1.f, 1.f - alpha);
                T = T * ra1;
                qy = alpha * T;
                float fd = r1.z * dpix[0];
                if (CH > 1) fd = fmaf(r1.w, dpix[1], fd);
#pragma unroll
                for (int k = 2; k < CH; k++) fd = fmaf(lds32(ra + 4u * (6 + k)), dpix[k], fd);
                const float dL_dalpha = fmaf(T, fd, -(ra1 * Bacc));
                Bacc = fmaf(qy, fd, Bacc);
                qx = G * dL_dalpha;
            }
            sts64(slot_cur, qx, qy);
            slot_cur += 8u * kSlotPitch;
            if (my_slot == nslot) my_j = j;
            nslot++;
            if (nslot == kSlots) {
                PXB_STAT(3, 1);
                flush(kSlots);
                nslot = 0;
            }
        }
        if (nslot) {  // the batch buffer is about to be restaged: drain
            PXB_STAT(3, 1); PXB_STAT(4, nslot);
            flush(nslot);
            nslot = 0;
        }
    }
}

// ---- operator-level pack / unpack -----------------------------------------
__global__ void pack_records_kernel(int P, const float2* __restrict__ uv, const float* __restrict__ conic,
                                    const float* __restrict__ opacity, const float* __restrict__ feature, int C,
                                    int c0, int cn, int S, float* __restrict__ rec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float* r = rec + (size_t)i * S;
    const float2 p = uv[i];
    r[0] = p.x; r[1] = p.y;
    r[2] = conic[3 * i]; r[3] = conic[3 * i + 1]; r[4] = conic[3 * i + 2];
    r[5] = opacity[i];
    for (int k = 0; k < S - 6; k++) r[6 + k] = (k < cn) ? feature[(size_t)i * C + c0 + k] : 0.f;
}

__global__ void unpack_grads_kernel(int P, const float* __restrict__ grec, int S, int C, int c0, int cn,
                                    int accumulate, float2* __restrict__ dL_duv, float* __restrict__ dL_dconic,
                                    float* __restrict__ dL_dopacity, float* __restrict__ dL_dfeature) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float* g = grec + (size_t)i * S;
    if (accumulate) {
        dL_duv[i].x += g[0]; dL_duv[i].y += g[1];
        dL_dconic[3 * i] += g[2]; dL_dconic[3 * i + 1] += g[3]; dL_dconic[3 * i + 2] += g[4];
        dL_dopacity[i] += g[5];
    } else {
        dL_duv[i] = make_float2(g[0], g[1]);
        dL_dconic[3 * i] = g[2]; dL_dconic[3 * i + 1] = g[3]; dL_dconic[3 * i + 2] = g[4];
        dL_dopacity[i] = g[5];
    }
    for (int k = 0; k < cn; k++) dL_dfeature[(size_t)i * C + c0 + k] = g[6 + k];
}

// PXB_BLEND_STAGING=bulk selects the cp.async.bulk + mbarrier double-buffered staging of the forward kernel
// (measured against the synchronous LDG/STS staging in profiles/)
static bool bulk_staging() {
    static const bool on = [] {
        const char* e = getenv("PXB_BLEND_STAGING");
        return e && e[0] == 'b';
    }();
    return on;
}

template <int CH, int S, bool BULK>
static int launch_fwd_v(const float* rec, const int* idx_sorted, const int* tile_range, float bg,
                        int C, int W, int H, float* final_T, int* ncontrib, float* out, cudaStream_t s) {
    const int gx = (W + PXB_TILE - 1) / PXB_TILE, gy = (H + PXB_TILE - 1) / PXB_TILE;
    const size_t smem = (size_t)(BULK ? 2 : 1) * (kBatch + 1) * S * 4 + (kBlendThreads / 32) * (kBatch + kUnroll) * 2 + kBatch;
    static bool attr = false;
    if (!attr) {
        if (smem > 48 * 1024)
            PXB_CUDA_OK(cudaFuncSetAttribute(blend_fwd_kernel<CH, S, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PXB_CUDA_OK(cudaFuncSetAttribute(blend_fwd_kernel<CH, S, BULK>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));
        attr = true;
    }
    PXB_CUDA_OK(launch_k(blend_fwd_kernel<CH, S, BULK>, dim3(gx, gy), dim3(kBlendThreads), smem, s, rec, idx_sorted,
                         (const int2*)tile_range, bg, C, W, H, gx, final_T, ncontrib, out));
    return (int)cudaGetLastError();
}

template <int CH, int S>
static int launch_fwd(const float* rec, const int* idx_sorted, const int* tile_range, float bg, int C,
                      int W, int H, float* final_T, int* ncontrib, float* out, cudaStream_t s) {
    if (bulk_staging() && S <= 16)  // two staging buffers: the wide strides keep the single-buffer variant
        return launch_fwd_v<CH, S, true>(rec, idx_sorted, tile_range, bg, C, W, H, final_T, ncontrib, out, s);
    return launch_fwd_v<CH, S, false>(rec, idx_sorted, tile_range, bg, C, W, H, final_T, ncontrib, out, s);
}

// C <= 4: 4 CTAs per SM with 64 registers (24 bytes spilled) -- measured on cfg4 0.557 ms, against 0.617 ms for
// 3 CTAs with 80 registers and no spill (PXB_BWD_OCC=3): the pair loop is issue bound and wants the warps
static int bwd_occupancy() {
    static const int v = [] {
        const char* e = getenv("PXB_BWD_OCC");
        return (e && e[0] == '3') ? 3 : 4;
    }();
    return v;
}

template <int CH, int S, int MINB>
static int launch_bwd_v(const float* rec, const int* idx_sorted, const int* tile_range, float bg, int C, int W, int H,
                        const float* final_T, const int* ncontrib, const float* dL_dout, float* grec, cudaStream_t s) {
    const int gx = (W + PXB_TILE - 1) / PXB_TILE, gy = (H + PXB_TILE - 1) / PXB_TILE;
    constexpr int CHP = (CH + 3) & ~3, NW = kBlendThreads / 32;
    const size_t smem = (size_t)(kBatch * S + NW * 32 * CHP + 2 * NW * kSlots * kSlotPitch) * 4 + kBatch * 4 + NW * kBatch * 2 + kBatch;
    static bool attr = false;
    if (!attr) {
        if (smem > 48 * 1024)
            PXB_CUDA_OK(cudaFuncSetAttribute(blend_bwd_kernel<CH, S, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PXB_CUDA_OK(cudaFuncSetAttribute(blend_bwd_kernel<CH, S, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));
        attr = true;
    }
    PXB_CUDA_OK(launch_k(blend_bwd_kernel<CH, S, MINB>, dim3(gx, gy), dim3(kBlendThreads), smem, s, rec, idx_sorted,
                         (const int2*)tile_range, bg, C, W, H, gx, final_T, ncontrib, dL_dout, grec));
    return (int)cudaGetLastError();
}

template <int CH, int S>
static int launch_bwd(const float* rec, const int* idx_sorted, const int* tile_range, float bg, int C, int W, int H,
                      const float* final_T, const int* ncontrib, const float* dL_dout, float* grec, cudaStream_t s) {
    if (CH > 4) return launch_bwd_v<CH, S, 1>(rec, idx_sorted, tile_range, bg, C, W, H, final_T, ncontrib, dL_dout, grec, s);
    if (bwd_occupancy() == 4)
        return launch_bwd_v<CH, S, (CH <= 4 ? 4 : 1)>(rec, idx_sorted, tile_range, bg, C, W, H, final_T, ncontrib, dL_dout, grec, s);
    return launch_bwd_v<CH, S, (CH <= 4 ? 3 : 1)>(rec, idx_sorted, tile_range, bg, C, W, H, final_T, ncontrib, dL_dout, grec, s);
}

// kernels are instantiated for the exact channel count up to 10 (rgb, rgb+depth, rgb+normal, ... the
// cfg4 set rgb+depth+normal+flow = 9), then for the two wide record strides
#define PXB_BLEND_DISPATCH(LAUNCH, ...)                              \
    switch (C) {                                                     \
        case 1: return LAUNCH<1, 8>(__VA_ARGS__);                    \
        case 2: return LAUNCH<2, 8>(__VA_ARGS__);                    \
        case 3: return LAUNCH<3, 12>(__VA_ARGS__);                   \
        case 4: return LAUNCH<4, 12>(__VA_ARGS__);                   \
        case 5: return LAUNCH<5, 12>(__VA_ARGS__);                   \
        case 6: return LAUNCH<6, 12>(__VA_ARGS__);                   \
        case 7: return LAUNCH<7, 16>(__VA_ARGS__);                   \
        case 8: return LAUNCH<8, 16>(__VA_ARGS__);                   \
        case 9: return LAUNCH<9, 16>(__VA_ARGS__);                   \
        case 10: return LAUNCH<10, 16>(__VA_ARGS__);                 \
        default:                                                     \
            if (C <= 18) return LAUNCH<18, 24>(__VA_ARGS__);         \
            return LAUNCH<26, 32>(__VA_ARGS__);                      \
    }

}  // namespace pxb

using namespace pxb;

extern "C" {

// record stride (floats) for C blended channels in one pass; C <= PXB_MAX_CHANNELS_PER_PASS
int pxb_record_stride(int C) {
    if (C <= 2) return 8;
    if (C <= 6) return 12;
    if (C <= 10) return 16;
    if (C <= 18) return 24;
    if (C <= 26) return 32;
    return -1;
}

int pxb_blend_forward(const float* rec, int S, int C, const int* idx_sorted, const int* tile_range, float bg, int W,
                      int H, float* final_T, int* ncontrib, float* out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (W <= 0 || H <= 0) return 0;
    if (C < 1 || C > PXB_MAX_CHANNELS_PER_PASS || S != pxb_record_stride(C)) return PXB_ERR_BAD_ARG;
    if (((uintptr_t)rec) & 15) return PXB_ERR_ALIGN;
    PXB_BLEND_DISPATCH(launch_fwd, rec, idx_sorted, tile_range, bg, C, W, H, final_T, ncontrib, out, s)
}

int pxb_blend_backward(const float* rec, int S, int C, const int* idx_sorted, const int* tile_range, float bg, int W,
                       int H, const float* final_T, const int* ncontrib, const float* dL_dout, float* grec,
                       void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (W <= 0 || H <= 0) return 0;
    if (C < 1 || C > PXB_MAX_CHANNELS_PER_PASS || S != pxb_record_stride(C)) return PXB_ERR_BAD_ARG;
    if ((((uintptr_t)rec) | ((uintptr_t)grec)) & 15) return PXB_ERR_ALIGN;
    PXB_BLEND_DISPATCH(launch_bwd, rec, idx_sorted, tile_range, bg, C, W, H, final_T, ncontrib, dL_dout, grec, s)
}

// counters: device pointer to >= 2 unsigned 64-bit words (NULL switches counting off).  Slot 0 accumulates
// the (8x4 block, Gaussian) candidates executed by blend forward launches, slot 1 by blend backward.
int pxb_blend_counters(unsigned long long* counters_dev) {
    return (int)cudaMemcpyToSymbol(g_blend_counters, &counters_dev, sizeof(counters_dev));
}

#ifdef PXB_STATS
int pxb_blend_stats(unsigned long long* host8, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host8, g_blend_stats, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_blend_stats, z, sizeof(z)); }
    return 0;
}
#endif

int pxb_pack_records(int P, const float* uv, const float* conic, const float* opacity, const float* feature, int C,
                     int c0, int cn, int S, float* rec, void* stream) {
    if (P <= 0) return 0;
    pack_records_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, (const float2*)uv, conic, opacity, feature,
                                                                          C, c0, cn, S, rec);
    return (int)cudaGetLastError();
}

int pxb_unpack_grads(int P, const float* grec, int S, int C, int c0, int cn, int accumulate, float* dL_duv,
                     float* dL_dconic, float* dL_dopacity, float* dL_dfeature, void* stream) {
    if (P <= 0) return 0;
    unpack_grads_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, grec, S, C, c0, cn, accumulate,
                                                                          (float2*)dL_duv, dL_dconic, dL_dopacity,
                                                                          dL_dfeature);
    return (int)cudaGetLastError();
}

}  // extern "C"
