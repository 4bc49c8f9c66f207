// Photometric loss at the render boundary: the step that follows the render path
// in every training iteration, `(1-l)*L1 + l*(1-SSIM)` of the rendered image against the ground truth
// (pointrix/model/base_model.py:113-120; l1_loss / l2_loss / ssim / _ssim / create_window of
// pointrix/model/loss.py:27-124).  The reference runs it as five depthwise 11x11 cuDNN convolutions plus
// a dozen elementwise kernels over [B,C,H,W], and autograd replays all of them; here it is
//
//   l1_ssim_fwd_kernel   one pass: a 32x32 tile (+5 halo) of both images in shared memory, the five
//                        Gaussian-window moments E[x] E[y] E[xx] E[yy] E[xy] by a separable 11-tap filter
//                        (horizontal into shared memory, vertical in registers), the SSIM map, the |x-y|
//                        map, their per-CTA sums, and -- when a backward will follow -- the three partial
//                        derivative maps dS/dE[x], dS/dE[xx], dS/dE[xy] the gradient is linear in;
//   loss_finalize_kernel fixed-order (deterministic) double-precision sum of the CTA partials per image;
//   l1_ssim_bwd_kernel   one pass: the same separable filter over the three derivative maps,
//                        dL/dx = w_ssim*(F*Dm + 2x F*D11 + y F*D12) + w_l1*sign(x-y), the weights read from
//                        device memory (the upstream gradient never visits the host).
//
// Zero padding like F.conv2d(padding=5) (loss.py:100-108): out-of-image pixels count as 0 in the moments
// and out-of-image derivative maps as 0 in the backward.  The window is separable by construction
// (create_window builds it as an outer product, loss.py:119-121), so the separable filter is the same
// operator up to fp32 summation order.
// Algorithmic bytes per pixel-channel: forward 8 read + 12 written (+0 without a backward), backward
// 20 read + 4 written.
#include <cmath>

#include "common.cuh"
#include "pointrix_b200.h"

namespace pxb {

constexpr int kLT = 32;              // tile edge (outputs)
constexpr int kLR = 5;               // window radius (window_size 11)
constexpr int kLI = kLT + 2 * kLR;   // 42: tile + halo
constexpr int kLP = 44;              // shared row pitch of the halo tiles (floats, 16-byte multiple)
constexpr int kLThreads = 256;

struct GaussWin {
    float w[11];
};

// gaussian(11, 1.5) of pointrix/model/loss.py:69-71: exp in double, stored as fp32, normalised in fp32
static GaussWin make_window() {
    GaussWin g;
    float s = 0.f;
    for (int i = 0; i < 11; i++) {
        g.w[i] = (float)std::exp(-(double)((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5));
        s += g.w[i];
    }
    for (int i = 0; i < 11; i++) g.w[i] = g.w[i] / s;
    return g;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;  // valid in thread 0
}

// loads the (42 x 42) halo tile of one image plane into shared memory, zero outside the image: warp = row,
// lane = column (two columns per lane), all 12 loads of a thread issued before the first store
__device__ __forceinline__ void load_halo(float (*dst)[kLP], const float* __restrict__ plane, int H, int W, int x0,
                                          int y0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kRounds = (kLI + kLThreads / 32 - 1) / (kLThreads / 32);  // 6
    const int gx0 = x0 + lane - kLR, gx1 = gx0 + 32;
    const bool ok0 = gx0 >= 0 && gx0 < W, ok1 = lane < kLP - 32 && gx1 < W;  // gx1 >= 27 always
    float v0[kRounds], v1[kRounds];
#pragma unroll
    for (int i = 0; i < kRounds; i++) {
        const int r = warp + i * (kLThreads / 32);
        const int gy = y0 + r - kLR;
        const bool rok = r < kLI && gy >= 0 && gy < H;
        const float* row = plane + (size_t)(rok ? gy : 0) * W;
        v0[i] = rok && ok0 ? __ldg(row + gx0) : 0.f;
        v1[i] = rok && ok1 && lane < kLI - 32 ? __ldg(row + gx1) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kRounds; i++) {
        const int r = warp + i * (kLThreads / 32);
        if (r < kLI) {
            dst[r][lane] = v0[i];
            if (lane < kLP - 32) dst[r][lane + 32] = v1[i];
        }
    }
}

template <bool STORE>
__global__ void __launch_bounds__(kLThreads)
l1_ssim_fwd_kernel(int H, int W, const float* __restrict__ pred, const float* __restrict__ gt,
                   float* __restrict__ dmaps, size_t map_stride, float* __restrict__ partial, GaussWin gw) {
    pdl_wait();
    __shared__ __align__(16) float sx[kLI][kLP];
    __shared__ __align__(16) float sy[kLI][kLP];
    __shared__ __align__(16) float hs[5][kLI][kLT];
    __shared__ float red[kLThreads / 32];

    const int plane = blockIdx.z, x0 = blockIdx.x * kLT, y0 = blockIdx.y * kLT;
    const size_t pofs = (size_t)plane * H * W;
    load_halo(sx, pred + pofs, H, W, x0, y0);
    load_halo(sy, gt + pofs, H, W, x0, y0);
    __syncthreads();

    // horizontal pass: one strip = 4 consecutive outputs of one halo row, 5 moments each
    for (int s = threadIdx.x; s < kLI * (kLT / 4); s += kLThreads) {
        const int r = s >> 3, c0 = (s & 7) * 4;
        float a[16], b[16];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float4 va = *reinterpret_cast<const float4*>(&sx[r][c0 + 4 * q]);
            const float4 vb = *reinterpret_cast<const float4*>(&sy[r][c0 + 4 * q]);
            a[4 * q] = va.x, a[4 * q + 1] = va.y, a[4 * q + 2] = va.z, a[4 * q + 3] = va.w;
            b[4 * q] = vb.x, b[4 * q + 1] = vb.y, b[4 * q + 2] = vb.z, b[4 * q + 3] = vb.w;
        }
        float xx[14], yy[14], xy[14];
#pragma unroll
        for (int j = 0; j < 14; j++) xx[j] = a[j] * a[j], yy[j] = b[j] * b[j], xy[j] = a[j] * b[j];
        float o[5][4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
            for (int k = 0; k < 11; k++) {
                const float w = gw.w[k];
                m1 = fmaf(w, a[q + k], m1);
                m2 = fmaf(w, b[q + k], m2);
                e11 = fmaf(w, xx[q + k], e11);
                e22 = fmaf(w, yy[q + k], e22);
                e12 = fmaf(w, xy[q + k], e12);
            }
            o[0][q] = m1, o[1][q] = m2, o[2][q] = e11, o[3][q] = e22, o[4][q] = e12;
        }
#pragma unroll
        for (int m = 0; m < 5; m++)
            *reinterpret_cast<float4*>(&hs[m][r][c0]) = make_float4(o[m][0], o[m][1], o[m][2], o[m][3]);
    }
    __syncthreads();

    // vertical pass: lane = column, each thread 4 consecutive rows
    const int col = threadIdx.x & 31, r0 = (threadIdx.x >> 5) * 4;
    float acc[5][4];
#pragma unroll
    for (int m = 0; m < 5; m++) {
        float v[14];
#pragma unroll
        for (int j = 0; j < 14; j++) v[j] = hs[m][r0 + j][col];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < 11; k++) t = fmaf(gw.w[k], v[q + k], t);
            acc[m][q] = t;
        }
    }

    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;  // loss.py:110-111
    float ssim_sum = 0.f, l1_sum = 0.f;
    const int gx = x0 + col;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int gy = y0 + r0 + q;
        if (gx < W && gy < H) {
            const float mu1 = acc[0][q], mu2 = acc[1][q];
            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
            const float s1 = acc[2][q] - mu1_sq, s2 = acc[3][q] - mu2_sq, s12 = acc[4][q] - mu12;
            const float A1 = 2.f * mu12 + C1, A2 = 2.f * s12 + C2;
            const float B1 = mu1_sq + mu2_sq + C1, B2 = s1 + s2 + C2;
            const float i1 = 1.f / B1, i2 = 1.f / B2, inv = i1 * i2;
            const float S = A1 * A2 * inv;  // loss.py:113
            ssim_sum += S;
            l1_sum += fabsf(sx[r0 + q + kLR][col + kLR] - sy[r0 + q + kLR][col + kLR]);
            if (STORE) {
                const float dS_ds1 = -S * i2;
                const float dS_ds12 = 2.f * A1 * inv;
                const float dS_dmu1 = 2.f * A2 * inv * (mu2 - mu1 * A1 * i1);
                const size_t o = pofs + (size_t)gy * W + gx;
                dmaps[o] = dS_dmu1 - 2.f * mu1 * dS_ds1 - mu2 * dS_ds12;  // dS/dE[x]
                dmaps[map_stride + o] = dS_ds1;                            // dS/dE[xx]
                dmaps[2 * map_stride + o] = dS_ds12;                       // dS/dE[xy]
            }
        }
    }
    const size_t nblk = (size_t)gridDim.x * gridDim.y * gridDim.z;
    const size_t bid = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const float ts = block_sum(ssim_sum, red);
    const float tl = block_sum(l1_sum, red);
    if (threadIdx.x == 0) {
        partial[bid] = tl;
        partial[nblk + bid] = ts;
    }
}

// per image b: out_k[b] = sum(partials of image b) / count, in a fixed order
__global__ void __launch_bounds__(256)
loss_finalize_kernel(const float* __restrict__ partial, size_t nblk, size_t per_image, double inv_count, float* out0,
                     float* out1) {
    pdl_wait();
    __shared__ double red[2][256];
    const int b = blockIdx.x;
    double s0 = 0.0, s1 = 0.0;
    for (size_t i = threadIdx.x; i < per_image; i += 256) {
        s0 += (double)partial[b * per_image + i];
        if (out1) s1 += (double)partial[nblk + b * per_image + i];
    }
    red[0][threadIdx.x] = s0, red[1][threadIdx.x] = s1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[0][threadIdx.x] += red[0][threadIdx.x + o], red[1][threadIdx.x] += red[1][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out0[b] = (float)(red[0][0] * inv_count);
        if (out1) out1[b] = (float)(red[1][0] * inv_count);
    }
}

// the training loss of base_model.py:117-120 straight from the CTA partials:
// out3 = { (1-l)*L1 + l*(1-SSIM), L1, 1-SSIM }, means over the whole batch
__global__ void __launch_bounds__(256)
loss_finalize_fused_kernel(const float* __restrict__ partial, size_t nblk, double inv_count, float lambda,
                           float* __restrict__ out3) {
    pdl_wait();
    __shared__ double red[2][256];
    double s0 = 0.0, s1 = 0.0;
    for (size_t i = threadIdx.x; i < nblk; i += 256) s0 += (double)partial[i], s1 += (double)partial[nblk + i];
    red[0][threadIdx.x] = s0, red[1][threadIdx.x] = s1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[0][threadIdx.x] += red[0][threadIdx.x + o], red[1][threadIdx.x] += red[1][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float l1 = (float)(red[0][0] * inv_count), sl = 1.0f - (float)(red[1][0] * inv_count);
        out3[0] = (1.0f - lambda) * l1 + lambda * sl;
        out3[1] = l1;
        out3[2] = sl;
    }
}

__global__ void __launch_bounds__(kLThreads)
l1_ssim_bwd_kernel(int C, int H, int W, const float* __restrict__ pred, const float* __restrict__ gt,
                   const float* __restrict__ dmaps, size_t map_stride, const float* __restrict__ g_l1,
                   const float* __restrict__ g_ssim, int g_stride, float s_l1, float s_ssim,
                   float* __restrict__ d_pred, GaussWin gw) {
    pdl_wait();
    __shared__ __align__(16) float sm[3][kLI][kLP];
    __shared__ __align__(16) float hs[3][kLI][kLT];

    const int plane = blockIdx.z, x0 = blockIdx.x * kLT, y0 = blockIdx.y * kLT;
    const size_t pofs = (size_t)plane * H * W;
#pragma unroll
    for (int m = 0; m < 3; m++) load_halo(sm[m], dmaps + m * map_stride + pofs, H, W, x0, y0);
    __syncthreads();

    for (int s = threadIdx.x; s < kLI * (kLT / 4); s += kLThreads) {
        const int r = s >> 3, c0 = (s & 7) * 4;
#pragma unroll
        for (int m = 0; m < 3; m++) {
            float a[16];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 va = *reinterpret_cast<const float4*>(&sm[m][r][c0 + 4 * q]);
                a[4 * q] = va.x, a[4 * q + 1] = va.y, a[4 * q + 2] = va.z, a[4 * q + 3] = va.w;
            }
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                float t = 0.f;
#pragma unroll
                for (int k = 0; k < 11; k++) t = fmaf(gw.w[k], a[q + k], t);
                o[q] = t;
            }
            *reinterpret_cast<float4*>(&hs[m][r][c0]) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
    __syncthreads();

    const int col = threadIdx.x & 31, r0 = (threadIdx.x >> 5) * 4;
    float acc[3][4];
#pragma unroll
    for (int m = 0; m < 3; m++) {
        float v[14];
#pragma unroll
        for (int j = 0; j < 14; j++) v[j] = hs[m][r0 + j][col];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < 11; k++) t = fmaf(gw.w[k], v[q + k], t);
            acc[m][q] = t;
        }
    }
    const int b = plane / C;
    const float ws = g_ssim ? s_ssim * __ldg(g_ssim + (size_t)b * g_stride) : 0.f;
    const float wl = g_l1 ? s_l1 * __ldg(g_l1 + (size_t)b * g_stride) : 0.f;
    const int gx = x0 + col;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int gy = y0 + r0 + q;
        if (gx < W && gy < H) {
            const size_t o = pofs + (size_t)gy * W + gx;
            const float x = __ldg(pred + o), y = __ldg(gt + o);
            const float sgn = (float)(x > y) - (float)(x < y);
            d_pred[o] = ws * (acc[0][q] + 2.f * x * acc[1][q] + y * acc[2][q]) + wl * sgn;
        }
    }
}

// ---- plain per-pixel losses (l1_loss / l2_loss of loss.py:27-67) ----------------------------------
template <int MODE>
__device__ __forceinline__ float pix_loss(float d) {
    return MODE == 1 ? fabsf(d) : d * d;
}

template <int MODE>
__global__ void __launch_bounds__(256)
pixel_loss_fwd_kernel(long long n, const float* __restrict__ pred, const float* __restrict__ gt,
                      float* __restrict__ map_out, float* __restrict__ partial) {
    __shared__ float red[8];
    const size_t base = (size_t)blockIdx.y * n;
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float v = pix_loss<MODE>(__ldg(pred + base + i) - __ldg(gt + base + i));
        if (map_out) map_out[base + i] = v;
        s += v;
    }
    const float t = block_sum(s, red);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
}

template <int MODE>
__global__ void __launch_bounds__(256)
pixel_loss_bwd_kernel(long long n, const float* __restrict__ pred, const float* __restrict__ gt,
                      const float* __restrict__ w, const float* __restrict__ g_map, float* __restrict__ d_pred) {
    const size_t base = (size_t)blockIdx.y * n;
    const float wb = w ? __ldg(w + blockIdx.y) : 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float d = __ldg(pred + base + i) - __ldg(gt + base + i);
        const float g = wb + (g_map ? __ldg(g_map + base + i) : 0.f);
        const float dv = MODE == 1 ? (float)(d > 0.f) - (float)(d < 0.f) : 2.f * d;
        d_pred[base + i] = g * dv;
    }
}

static inline int pixel_grid(long long n) {
    long long g = (n + 255) / 256;
    const long long cap = 148 * 8;  // 8 resident CTAs of 256 threads per SM
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace pxb

using namespace pxb;

extern "C" {

size_t pxb_loss_workspace_bytes(int B, int C, int H, int W) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    const size_t tiles = (size_t)((W + kLT - 1) / kLT) * ((H + kLT - 1) / kLT) * (size_t)B * C;
    const size_t pix = (size_t)B * pixel_grid((long long)C * H * W);
    const size_t n = 2 * tiles > pix ? 2 * tiles : pix;
    return n * sizeof(float);
}

static int l1_ssim_launch(int B, int C, int H, int W, const float* pred, const float* gt, float* dmaps, void* ws,
                          size_t ws_bytes, cudaStream_t st, size_t* nblk_out) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || !pred || !gt || !ws) return PXB_ERR_BAD_ARG;
    if ((long long)B * C > 65535) return PXB_ERR_UNSUPPORTED;
    if (ws_bytes < pxb_loss_workspace_bytes(B, C, H, W)) return PXB_ERR_WORKSPACE;
    static const GaussWin gw = make_window();
    const dim3 grid((W + kLT - 1) / kLT, (H + kLT - 1) / kLT, B * C);
    const size_t map_stride = (size_t)B * C * H * W;
    float* partial = (float*)ws;
    if (dmaps)
        PXB_CUDA_OK(launch_k(l1_ssim_fwd_kernel<true>, grid, dim3(kLThreads), 0, st, H, W, pred, gt, dmaps, map_stride, partial, gw));
    else
        PXB_CUDA_OK(launch_k(l1_ssim_fwd_kernel<false>, grid, dim3(kLThreads), 0, st, H, W, pred, gt, (float*)nullptr, map_stride, partial, gw));
    *nblk_out = (size_t)grid.x * grid.y * grid.z;
    return 0;
}

int pxb_l1_ssim_forward(int B, int C, int H, int W, const float* pred, const float* gt, float* dmaps, float* l1_mean,
                        float* ssim_mean, void* ws, size_t ws_bytes, void* stream) {
    if (!l1_mean || !ssim_mean) return PXB_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    size_t nblk = 0;
    const int rc = l1_ssim_launch(B, C, H, W, pred, gt, dmaps, ws, ws_bytes, st, &nblk);
    if (rc) return rc;
    PXB_CUDA_OK(launch_k(loss_finalize_kernel, dim3(B), dim3(256), 0, st, (const float*)ws, nblk, nblk / B,
                         1.0 / ((double)C * H * W), l1_mean, ssim_mean));
    return (int)cudaGetLastError();
}

int pxb_l1_ssim_loss_forward(int B, int C, int H, int W, const float* pred, const float* gt, float lambda_ssim,
                             float* dmaps, float* loss3, void* ws, size_t ws_bytes, void* stream) {
    if (!loss3) return PXB_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    size_t nblk = 0;
    const int rc = l1_ssim_launch(B, C, H, W, pred, gt, dmaps, ws, ws_bytes, st, &nblk);
    if (rc) return rc;
    PXB_CUDA_OK(launch_k(loss_finalize_fused_kernel, dim3(1), dim3(256), 0, st, (const float*)ws, nblk,
                         1.0 / ((double)B * C * H * W), lambda_ssim, loss3));
    return (int)cudaGetLastError();
}

int pxb_l1_ssim_backward(int B, int C, int H, int W, const float* pred, const float* gt, const float* dmaps,
                         const float* g_l1, const float* g_ssim, int g_stride, float s_l1, float s_ssim, float* d_pred,
                         void* stream) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || !pred || !gt || !dmaps || !d_pred) return PXB_ERR_BAD_ARG;
    if ((long long)B * C > 65535) return PXB_ERR_UNSUPPORTED;
    static const GaussWin gw = make_window();
    const dim3 grid((W + kLT - 1) / kLT, (H + kLT - 1) / kLT, B * C);
    PXB_CUDA_OK(launch_k(l1_ssim_bwd_kernel, grid, dim3(kLThreads), 0, (cudaStream_t)stream, C, H, W, pred, gt, dmaps,
                         (size_t)B * C * H * W, g_l1, g_ssim, g_stride, s_l1, s_ssim, d_pred, gw));
    return (int)cudaGetLastError();
}

int pxb_pixel_loss_forward(int mode, int B, long long n, const float* pred, const float* gt, float* map_out,
                           float* mean_out, void* ws, size_t ws_bytes, void* stream) {
    if ((mode != 1 && mode != 2) || B <= 0 || B > 65535 || n <= 0 || !pred || !gt || !mean_out || !ws)
        return PXB_ERR_BAD_ARG;
    const int g = pixel_grid(n);
    if (ws_bytes < (size_t)B * g * sizeof(float)) return PXB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float* partial = (float*)ws;
    if (mode == 1)
        pixel_loss_fwd_kernel<1><<<dim3(g, B), 256, 0, st>>>(n, pred, gt, map_out, partial);
    else
        pixel_loss_fwd_kernel<2><<<dim3(g, B), 256, 0, st>>>(n, pred, gt, map_out, partial);
    loss_finalize_kernel<<<B, 256, 0, st>>>(partial, (size_t)B * g, (size_t)g, 1.0 / (double)n, mean_out, nullptr);
    return (int)cudaGetLastError();
}

int pxb_pixel_loss_backward(int mode, int B, long long n, const float* pred, const float* gt, const float* w,
                            const float* g_map, float* d_pred, void* stream) {
    if ((mode != 1 && mode != 2) || B <= 0 || B > 65535 || n <= 0 || !pred || !gt || !d_pred) return PXB_ERR_BAD_ARG;
    const int g = pixel_grid(n);
    if (mode == 1)
        pixel_loss_bwd_kernel<1><<<dim3(g, B), 256, 0, (cudaStream_t)stream>>>(n, pred, gt, w, g_map, d_pred);
    else
        pixel_loss_bwd_kernel<2><<<dim3(g, B), 256, 0, (cudaStream_t)stream>>>(n, pred, gt, w, g_map, d_pred);
    return (int)cudaGetLastError();
}

}  // extern "C"
