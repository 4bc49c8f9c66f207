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
            if (q == 0) sm_id[j] = g;
            const float4 v = __ldg(reinterpret_cast<const float4*>(rec + (size_t)g * S) + q);
            reinterpret_cast<float4*>(sm_rec)[f] = v;
        }
    }
}

template <int CH, int S>
__global__ void __launch_bounds__(kBlendThreads)
blend_fwd_kernel(const float* __restrict__ rec, const int* __restrict__ idx_sorted, const int2* __restrict__ tile_range,
                 float bg, int C, int W, int H, int gx, float* __restrict__ final_T, int* __restrict__ ncontrib,
                 float* __restrict__ out) {
    extern __shared__ __align__(16) float sm_f[];
    float* sm_rec = sm_f;
    int* sm_id = reinterpret_cast<int*>(sm_f + kBatch * S);
    unsigned char* sm_mask = reinterpret_cast<unsigned char*>(sm_id + kBatch);
    const int tile = blockIdx.y * gx + blockIdx.x;
    int lx, ly;
    thread_pixel(lx, ly);
    const int pxi = blockIdx.x * PXB_TILE + lx, pyi = blockIdx.y * PXB_TILE + ly;
    const float pxf = (float)pxi, pyf = (float)pyi;
    const bool inside = pxi < W && pyi < H;
    bool done = !inside;
    const int2 range = tile_range[tile];
    int todo = range.y - range.x;
    float T = 1.0f;
    int last = 0;
    float F[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) F[k] = 0.f;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const float X0 = (float)(blockIdx.x * PXB_TILE), Y0 = (float)(blockIdx.y * PXB_TILE);

    for (int base = 0; todo > 0; base += kBatch, todo -= kBatch) {
        if (__syncthreads_count(done) == kBlendThreads) break;
        const int n = min(kBatch, todo);
        stage_records<S>(sm_rec, sm_id, rec, idx_sorted + range.x + base, n);
        __syncthreads();
        if ((int)threadIdx.x < n) {
            const float4 r0 = *reinterpret_cast<const float4*>(sm_rec + threadIdx.x * S);
            const float2 r1 = *reinterpret_cast<const float2*>(sm_rec + threadIdx.x * S + 4);
            sm_mask[threadIdx.x] = (unsigned char)block_mask(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, X0, Y0);
        }
        __syncthreads();
        if (__all_sync(0xffffffffu, done)) continue;  // whole warp saturated: it only helps staging
        for (int c0 = 0; c0 < n; c0 += 32) {
            const int jj = c0 + (int)lane;
            unsigned int m = __ballot_sync(0xffffffffu, jj < n && ((sm_mask[jj] >> warp) & 1u));
            while (m) {
                const int j = c0 + __ffs(m) - 1;
                m &= m - 1;
                if (done) continue;
                const float4 r0 = *reinterpret_cast<const float4*>(sm_rec + j * S);      // u v A B
                const float4 r1 = *reinterpret_cast<const float4*>(sm_rec + j * S + 4);  // C op f0 f1
                const float dx = __fadd_rn(r0.x, -pxf), dy = __fadd_rn(r0.y, -pyf);
                const float power = blend_power(dx, dy, r0.z, r0.w, r1.x);
                if (power > 0.f) continue;
                const float alpha = fmin_nn(__fmul_rn(r1.y, blend_G(power)), 0.99f);
                if (alpha < 1.0f / 255.0f) continue;
                const float next_T = __fmul_rn(T, __fadd_rn(-alpha, 1.0f));
                if (next_T < 0.0001f) { done = true; continue; }
                F[0] = __fmaf_rn(T, __fmul_rn(alpha, r1.z), F[0]);
                if (CH > 1) F[1] = __fmaf_rn(T, __fmul_rn(alpha, r1.w), F[1]);
#pragma unroll
                for (int q = 2; q < CH; q += 4) {
                    const float4 rf = *reinterpret_cast<const float4*>(sm_rec + j * S + 6 + q);
                    F[q] = __fmaf_rn(T, __fmul_rn(alpha, rf.x), F[q]);
                    if (q + 1 < CH) F[q + 1] = __fmaf_rn(T, __fmul_rn(alpha, rf.y), F[q + 1]);
                    if (q + 2 < CH) F[q + 2] = __fmaf_rn(T, __fmul_rn(alpha, rf.z), F[q + 2]);
                    if (q + 3 < CH) F[q + 3] = __fmaf_rn(T, __fmul_rn(alpha, rf.w), F[q + 3]);
                }
                T = next_T;
                last = base + j + 1;  // 1-based list position of the last blended Gaussian (ncontrib)
            }
            if (__all_sync(0xffffffffu, done)) break;
        }
    }
    if (inside) {
        const size_t pix = (size_t)pyi * W + pxi;
        final_T[pix] = T;
        ncontrib[pix] = last;
#pragma unroll
        for (int k = 0; k < CH; k++)
            if (k < C) out[(size_t)k * H * W + pix] = __fmaf_rn(T, bg, F[k]);
    }
}

// Sum NV per-lane values across the warp.  v[] is padded to NVP (power of two
// >= NV, <= 32).  On return lane L holds in v[0] the total of value index
// L / (32/NVP): each butterfly stage exchanges half of the remaining values.
template <int NVP>
__device__ __forceinline__ void warp_reduce_vec(float (&v)[NVP]) {
    const unsigned lane = threadIdx.x & 31u;
    int n = NVP;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        if (n > 1) {
            const int half = n / 2;
            const bool upper = (lane & step) != 0;
#pragma unroll
            for (int k = 0; k < NVP / 2; k++) {
                if (k < half) {
                    const float send = upper ? v[k] : v[k + half];
                    const float keep = upper ? v[k + half] : v[k];
                    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, step);
                }
            }
            n = half;
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], step);
        }
    }
}

template <int CH, int S, int NVP>
__global__ void __launch_bounds__(kBlendThreads)
blend_bwd_kernel(const float* __restrict__ rec, const int* __restrict__ idx_sorted, const int2* __restrict__ tile_range,
                 float bg, int C, int W, int H, int gx, const float* __restrict__ final_T,
                 const int* __restrict__ ncontrib, const float* __restrict__ dL_dout, float* __restrict__ grec) {
    extern __shared__ __align__(16) float sm_f[];
    float* sm_rec = sm_f;
    int* sm_id = reinterpret_cast<int*>(sm_f + kBatch * S);
    unsigned char* sm_mask = reinterpret_cast<unsigned char*>(sm_id + kBatch);
    const int tile = blockIdx.y * gx + blockIdx.x;
    int lx, ly;
    thread_pixel(lx, ly);
    const int pxi = blockIdx.x * PXB_TILE + lx, pyi = blockIdx.y * PXB_TILE + ly;
    const float pxf = (float)pxi, pyf = (float)pyi;
    const bool inside = pxi < W && pyi < H;
    const size_t pix = (size_t)pyi * W + pxi;
    const int2 range = tile_range[tile];
    const int count = range.y - range.x;
    const float T_final = inside ? final_T[pix] : 0.f;
    float T = T_final;
    const int last = inside ? ncontrib[pix] : 0;
    float accum[CH], dpix[CH], lastf[CH];
    float bg_dot = 0.f;
#pragma unroll
    for (int k = 0; k < CH; k++) {
        accum[k] = 0.f; lastf[k] = 0.f;
        dpix[k] = (inside && k < C) ? dL_dout[(size_t)k * H * W + pix] : 0.f;
        bg_dot += bg * dpix[k];
    }
    float last_alpha = 0.f;
    // the deepest contributor of any pixel of the CTA / of this warp bounds the work: list entries
    // at positions >= that bound were never blended by these pixels
    const int warp_last = __reduce_max_sync(0xffffffffu, last);
    __shared__ int s_cta_last;
    if (threadIdx.x == 0) s_cta_last = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) atomicMax(&s_cta_last, warp_last);
    __syncthreads();
    const int cta_last = s_cta_last;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const float X0 = (float)(blockIdx.x * PXB_TILE), Y0 = (float)(blockIdx.y * PXB_TILE);

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
        for (int c0 = 0; c0 < n; c0 += 32) {
            const int jj = c0 + (int)lane;
            // positions >= warp_last contribute to no pixel of this warp
            unsigned int m = __ballot_sync(0xffffffffu, jj < n && (top - 1 - jj) < warp_last && ((sm_mask[jj] >> warp) & 1u));
            while (m) {
                const int j = c0 + __ffs(m) - 1;
                m &= m - 1;
                const int pos = top - 1 - j;  // list position of this entry
                bool valid = pos < last;
                float dx = 0.f, dy = 0.f, G = 0.f, alpha = 0.f;
                const float4 r0 = *reinterpret_cast<const float4*>(sm_rec + j * S);
                const float4 r1 = *reinterpret_cast<const float4*>(sm_rec + j * S + 4);
                if (valid) {
                    dx = __fadd_rn(r0.x, -pxf); dy = __fadd_rn(r0.y, -pyf);
                    const float power = blend_power(dx, dy, r0.z, r0.w, r1.x);
                    G = blend_G(power);
                    alpha = fmin_nn(__fmul_rn(r1.y, G), 0.99f);
                    valid = !(power > 0.f) && !(alpha < 1.0f / 255.0f);
                }
                if (!__any_sync(0xffffffffu, valid)) continue;
                float v[NVP];
#pragma unroll
                for (int k = 0; k < NVP; k++) v[k] = 0.f;
                if (valid) {
                    T = __fdividef(T, 1.f - alpha);
                    const float w = alpha * T;
                    float dL_dalpha = 0.f;
#pragma unroll
                    for (int k = 0; k < CH; k++) {
                        const float f = sm_rec[j * S + 6 + k];
                        accum[k] = last_alpha * lastf[k] + (1.f - last_alpha) * accum[k];
                        lastf[k] = f;
                        dL_dalpha += (f - accum[k]) * dpix[k];
                        v[6 + k] = w * dpix[k];
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                    const float dL_dG = r1.y * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    v[0] = dL_dG * (-gdx * r0.z - gdy * r0.w);
                    v[1] = dL_dG * (-gdy * r1.x - gdx * r0.w);
                    v[2] = -0.5f * gdx * dx * dL_dG;
                    v[3] = -gdx * dy * dL_dG;
                    v[4] = -0.5f * gdy * dy * dL_dG;
                    v[5] = G * dL_dalpha;
                }
                warp_reduce_vec<NVP>(v);
                constexpr int REP = 32 / NVP;
                const int idx = lane / REP;
                if ((lane % REP) == 0 && idx < 6 + C && v[0] != 0.f) atomicAdd(grec + (size_t)sm_id[j] * S + idx, v[0]);
            }
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

template <int CH, int S>
static int launch_fwd(const float* rec, const int* idx_sorted, const int* tile_range, float bg, int C, int W, int H,
                      float* final_T, int* ncontrib, float* out, cudaStream_t s) {
    const int gx = (W + PXB_TILE - 1) / PXB_TILE, gy = (H + PXB_TILE - 1) / PXB_TILE;
    const size_t smem = (size_t)kBatch * S * 4 + kBatch * 4 + kBatch;
    static bool attr = false;
    if (!attr && smem > 48 * 1024) {
        PXB_CUDA_OK(cudaFuncSetAttribute(blend_fwd_kernel<CH, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    blend_fwd_kernel<CH, S><<<dim3(gx, gy), kBlendThreads, smem, s>>>(rec, idx_sorted, (const int2*)tile_range, bg, C, W,
                                                                      H, gx, final_T, ncontrib, out);
    return (int)cudaGetLastError();
}

template <int CH, int S, int NVP>
static int launch_bwd(const float* rec, const int* idx_sorted, const int* tile_range, float bg, int C, int W, int H,
                      const float* final_T, const int* ncontrib, const float* dL_dout, float* grec, cudaStream_t s) {
    const int gx = (W + PXB_TILE - 1) / PXB_TILE, gy = (H + PXB_TILE - 1) / PXB_TILE;
    const size_t smem = (size_t)kBatch * S * 4 + kBatch * 4 + kBatch;
    static bool attr = false;
    if (!attr && smem > 48 * 1024) {
        PXB_CUDA_OK(cudaFuncSetAttribute(blend_bwd_kernel<CH, S, NVP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    blend_bwd_kernel<CH, S, NVP><<<dim3(gx, gy), kBlendThreads, smem, s>>>(
        rec, idx_sorted, (const int2*)tile_range, bg, C, W, H, gx, final_T, ncontrib, dL_dout, grec);
    return (int)cudaGetLastError();
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
    if (((uintptr_t)rec) & 15) return PXB_ERR_ALIGN;
    switch (S) {
        case 8: return launch_fwd<2, 8>(rec, idx_sorted, tile_range, bg, C, W, H, final_T, ncontrib, out, s);
        case 12: return launch_fwd<6, 12>(rec, idx_sorted, tile_range, bg, C, W, H, final_T, ncontrib, out, s);
        case 16: return launch_fwd<10, 16>(rec, idx_sorted, tile_range, bg, C, W, H, final_T, ncontrib, out, s);
        case 24: return launch_fwd<18, 24>(rec, idx_sorted, tile_range, bg, C, W, H, final_T, ncontrib, out, s);
        case 32: return launch_fwd<26, 32>(rec, idx_sorted, tile_range, bg, C, W, H, final_T, ncontrib, out, s);
    }
    return PXB_ERR_BAD_ARG;
}

int pxb_blend_backward(const float* rec, int S, int C, const int* idx_sorted, const int* tile_range, float bg, int W,
                       int H, const float* final_T, const int* ncontrib, const float* dL_dout, float* grec,
                       void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (W <= 0 || H <= 0) return 0;
    if ((((uintptr_t)rec) | ((uintptr_t)grec)) & 15) return PXB_ERR_ALIGN;
    switch (S) {
        case 8: return launch_bwd<2, 8, 8>(rec, idx_sorted, tile_range, bg, C, W, H, final_T, ncontrib, dL_dout, grec, s);
        case 12: return launch_bwd<6, 12, 16>(rec, idx_sorted, tile_range, bg, C, W, H, final_T, ncontrib, dL_dout, grec, s);
        case 16: return launch_bwd<10, 16, 16>(rec, idx_sorted, tile_range, bg, C, W, H, final_T, ncontrib, dL_dout, grec, s);
        case 24: return launch_bwd<18, 24, 32>(rec, idx_sorted, tile_range, bg, C, W, H, final_T, ncontrib, dL_dout, grec, s);
        case 32: return launch_bwd<26, 32, 32>(rec, idx_sorted, tile_range, bg, C, W, H, final_T, ncontrib, dL_dout, grec, s);
    }
    return PXB_ERR_BAD_ARG;
}

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
