// Per-Gaussian stages of the render path: operator-level kernels with the
// reference's tensor layouts (project_point / compute_cov3d / ewa_project /
// compute_sh, forward + backward) and the fused plugin-level kernels that run
// the whole per-Gaussian forward (SH colour + projection + covariance + EWA +
// record packing) and backward in one HBM pass each.
//
// All of them are HBM streams: one thread per Gaussian, inputs read once,
// outputs written once, camera gradients reduced per block before touching
// global atomics.
#include <stdlib.h>

#include "common.cuh"
#include "sh.cuh"
#include "pointrix_b200.h"

namespace pxb {

constexpr int kThreads = 256;
static inline int blocks_for(int n, int t = kThreads) { return (n + t - 1) / t; }

// Block-wide sum of NV per-thread values -> one atomicAdd per value per block.
template <int NV>
__device__ __forceinline__ void block_reduce_atomic(float (&v)[NV], float* __restrict__ dst, float* smem /*[NV*warps]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        float x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) smem[k * nwarps + warp] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        float s = 0.f;
        for (int w = 0; w < nwarps; w++) s += smem[threadIdx.x * nwarps + w];
        if (s != 0.f) atomicAdd(dst + threadIdx.x, s);
    }
}

// ---------------------------------------------------------------------------
// project_point  (msplat/msplat/src/project_point.cu:13-57, 59-145)
// ---------------------------------------------------------------------------
__global__ void project_fwd_kernel(int P, const float* __restrict__ xyz, const float* __restrict__ intr,
                                   const float* __restrict__ extr, int W, int H, float nearest, float extent,
                                   float2* __restrict__ uv, float* __restrict__ depth) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const Cam c = load_cam(intr, extr);
    const float3 t = cam_transform(c, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    float u, v;
    const bool keep = project_uv(c, t, W, H, nearest, extent, u, v);
    uv[i] = keep ? make_float2(u, v) : make_float2(0.f, 0.f);
    depth[i] = keep ? t.z : 0.f;
}

// dL/dt from (dL/duv, dL/ddepth): the chain of project_point.cu:86-104
__device__ __forceinline__ float3 project_dt(const Cam& c, const float3 t, float du, float dv, float dd) {
    const float n1 = 1.0f / t.z, n2 = n1 * n1;
    float3 g;
    g.x = c.fx * n1 * du;
    g.y = c.fy * n1 * dv;
    g.z = dd - (c.fx * t.x * du + c.fy * t.y * dv) * n2;
    return g;
}

__global__ void project_bwd_kernel(int P, const float* __restrict__ xyz, const float* __restrict__ intr,
                                   const float* __restrict__ extr, const float* __restrict__ depth,
                                   const float2* __restrict__ dL_duv, const float* __restrict__ dL_ddepth,
                                   float* __restrict__ dL_dxyz, float* __restrict__ dL_dintr,
                                   float* __restrict__ dL_dextr) {
    __shared__ float red[16 * (kThreads / 32)];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const Cam c = load_cam(intr, extr);
    float cg[16];
#pragma unroll
    for (int k = 0; k < 16; k++) cg[k] = 0.f;
    if (i < P) {
        float3 gx = make_float3(0.f, 0.f, 0.f);
        if (depth[i] != 0.f) {  // depth == 0 means culled (project_point.cu:74-76)
            const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
            const float3 t = cam_transform(c, px, py, pz);
            const float2 duv = dL_duv[i];
            const float3 dt = project_dt(c, t, duv.x, duv.y, dL_ddepth[i]);
            gx.x = c.e[0] * dt.x + c.e[4] * dt.y + c.e[8] * dt.z;
            gx.y = c.e[1] * dt.x + c.e[5] * dt.y + c.e[9] * dt.z;
            gx.z = c.e[2] * dt.x + c.e[6] * dt.y + c.e[10] * dt.z;
            const float n1 = 1.0f / t.z;
            cg[0] = t.x * n1 * duv.x; cg[1] = t.y * n1 * duv.y; cg[2] = duv.x; cg[3] = duv.y;
            cg[4] = px * dt.x; cg[5] = py * dt.x; cg[6] = pz * dt.x; cg[7] = dt.x;
            cg[8] = px * dt.y; cg[9] = py * dt.y; cg[10] = pz * dt.y; cg[11] = dt.y;
            cg[12] = px * dt.z; cg[13] = py * dt.z; cg[14] = pz * dt.z; cg[15] = dt.z;
        }
        dL_dxyz[3 * i] = gx.x; dL_dxyz[3 * i + 1] = gx.y; dL_dxyz[3 * i + 2] = gx.z;
    }
    if (dL_dintr != nullptr || dL_dextr != nullptr) {
        // one reduction, two destinations
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            float x = cg[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) red[k * nwarps + warp] = x;
        }
        __syncthreads();
        if (threadIdx.x < 16) {
            float s = 0.f;
            for (int w = 0; w < nwarps; w++) s += red[threadIdx.x * nwarps + w];
            if (s != 0.f) {
                if (threadIdx.x < 4) { if (dL_dintr) atomicAdd(dL_dintr + threadIdx.x, s); }
                else if (dL_dextr) atomicAdd(dL_dextr + threadIdx.x - 4, s);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// compute_cov3d  (msplat/msplat/src/compute_cov3d.cu:42-58, 60-117)
// ---------------------------------------------------------------------------
__global__ void cov3d_fwd_kernel(int P, const float* __restrict__ scales, const float4* __restrict__ quats,
                                 const uint8_t* __restrict__ visible, float* __restrict__ cov3d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float cov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (visible == nullptr || visible[i]) {
        const float4 q = quats[i];
        cov3d_from_scale_quat(scales[3 * i], scales[3 * i + 1], scales[3 * i + 2], q.x, q.y, q.z, q.w, cov);
    }
#pragma unroll
    for (int k = 0; k < 6; k++) cov3d[6 * i + k] = cov[k];
}

// dL/dscale, dL/dquat from dL/dcov3d[6]; Sigma = R S^2 R^T, q un-normalised.
__device__ __forceinline__ void cov3d_backward(float sx, float sy, float sz, float r, float x, float y, float z,
                                               const float g[6], float ds[3], float dq[4]) {
    float R[3][3];
    R[0][0] = 1.f - 2.f * (y * y + z * z); R[0][1] = 2.f * (x * y - r * z); R[0][2] = 2.f * (x * z + r * y);
    R[1][0] = 2.f * (x * y + r * z); R[1][1] = 1.f - 2.f * (x * x + z * z); R[1][2] = 2.f * (y * z - r * x);
    R[2][0] = 2.f * (x * z - r * y); R[2][1] = 2.f * (y * z + r * x); R[2][2] = 1.f - 2.f * (x * x + y * y);
    const float s[3] = {sx, sy, sz};
    // G symmetric with halved off-diagonals; dN = 2 G N, N = R S
    const float G[3][3] = {{g[0], 0.5f * g[1], 0.5f * g[2]}, {0.5f * g[1], g[3], 0.5f * g[4]}, {0.5f * g[2], 0.5f * g[4], g[5]}};
    float dR[3][3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const float dN = 2.f * s[k] * (G[i][0] * R[0][k] + G[i][1] * R[1][k] + G[i][2] * R[2][k]);
            acc += dN * R[i][k];
            dR[i][k] = dN * s[k];
        }
        ds[k] = acc;
    }
    dq[0] = 2.f * (-z * dR[0][1] + y * dR[0][2] + z * dR[1][0] - x * dR[1][2] - y * dR[2][0] + x * dR[2][1]);
    dq[1] = 2.f * (y * dR[0][1] + z * dR[0][2] + y * dR[1][0] - 2.f * x * dR[1][1] - r * dR[1][2] + z * dR[2][0] + r * dR[2][1] - 2.f * x * dR[2][2]);
    dq[2] = 2.f * (-2.f * y * dR[0][0] + x * dR[0][1] + r * dR[0][2] + x * dR[1][0] + z * dR[1][2] - r * dR[2][0] + z * dR[2][1] - 2.f * y * dR[2][2]);
    dq[3] = 2.f * (-2.f * z * dR[0][0] - r * dR[0][1] + x * dR[0][2] + r * dR[1][0] - 2.f * z * dR[1][1] + y * dR[1][2] + x * dR[2][0] + y * dR[2][1]);
}

__global__ void cov3d_bwd_kernel(int P, const float* __restrict__ scales, const float4* __restrict__ quats,
                                 const uint8_t* __restrict__ visible, const float* __restrict__ dL_dcov3d,
                                 float* __restrict__ dL_dscales, float4* __restrict__ dL_dquats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float ds[3] = {0.f, 0.f, 0.f}, dq[4] = {0.f, 0.f, 0.f, 0.f};
    if (visible == nullptr || visible[i]) {
        float g[6];
#pragma unroll
        for (int k = 0; k < 6; k++) g[k] = dL_dcov3d[6 * i + k];
        const float4 q = quats[i];
        cov3d_backward(scales[3 * i], scales[3 * i + 1], scales[3 * i + 2], q.x, q.y, q.z, q.w, g, ds, dq);
    }
    dL_dscales[3 * i] = ds[0]; dL_dscales[3 * i + 1] = ds[1]; dL_dscales[3 * i + 2] = ds[2];
    dL_dquats[i] = make_float4(dq[0], dq[1], dq[2], dq[3]);
}

// ---------------------------------------------------------------------------
// ewa_project  (msplat/msplat/src/ewa_project.cu:16-83, 85-252)
// ---------------------------------------------------------------------------
__global__ void ewa_fwd_kernel(int P, const float* __restrict__ xyz, const float* __restrict__ cov3d,
                               const float* __restrict__ intr, const float* __restrict__ extr,
                               const float2* __restrict__ uv, int gx, int gy, const uint8_t* __restrict__ visible,
                               float* __restrict__ conic, int* __restrict__ radius, int* __restrict__ tiles) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float cn[3] = {0.f, 0.f, 0.f};
    int rad = 0, nt = 0;
    if (visible == nullptr || visible[i]) {
        const Cam c = load_cam(intr, extr);
        const float3 t = cam_transform(c, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        const EwaT T = ewa_T(c, t);
        float v[6];
#pragma unroll
        for (int k = 0; k < 6; k++) v[k] = cov3d[6 * i + k];
        float a, b, cc, det;
        ewa_cov2d(T, v, a, b, cc);
        const float2 p = uv[i];
        int r_, t_;
        float cn_[3];
        if (ewa_finish(a, b, cc, p.x, p.y, gx, gy, det, r_, t_, cn_)) {
            rad = r_; nt = t_; cn[0] = cn_[0]; cn[1] = cn_[1]; cn[2] = cn_[2];
        }
    }
    conic[3 * i] = cn[0]; conic[3 * i + 1] = cn[1]; conic[3 * i + 2] = cn[2];
    radius[i] = rad;
    tiles[i] = nt;
}

// Gradients of conic wrt cov3d, t (-> xyz), intr[0:2], extr.  cg[] layout:
// [0..1] dfx,dfy ; [2..13] dextr row-major.
__device__ __forceinline__ void ewa_backward(const Cam& c, float px, float py, float pz, const float v[6],
                                             const float dcn[3], float dV[6], float3& dxyz, float cg[14],
                                             bool& live) {
    const float3 t = cam_transform(c, px, py, pz);
    const float iz = 1.0f / t.z, iz2 = iz * iz, iz3 = iz2 * iz;
    const float J00 = c.fx * iz, J11 = c.fy * iz, J20 = -c.fx * t.x * iz2, J21 = -c.fy * t.y * iz2;
    float T0[3], T1[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        T0[k] = J00 * c.e[k] + J20 * c.e[8 + k];
        T1[k] = J11 * c.e[4 + k] + J21 * c.e[8 + k];
    }
    const float V[3][3] = {{v[0], v[1], v[2]}, {v[1], v[3], v[4]}, {v[2], v[4], v[5]}};
    float Ax[3], Ay[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        Ax[k] = V[k][0] * T0[0] + V[k][1] * T0[1] + V[k][2] * T0[2];
        Ay[k] = V[k][0] * T1[0] + V[k][1] * T1[1] + V[k][2] * T1[2];
    }
    const float a = T0[0] * Ax[0] + T0[1] * Ax[1] + T0[2] * Ax[2] + 0.3f;
    const float b = T0[0] * Ay[0] + T0[1] * Ay[1] + T0[2] * Ay[2];
    const float cc = T1[0] * Ay[0] + T1[1] * Ay[1] + T1[2] * Ay[2] + 0.3f;
    const float det = a * cc - b * b;
    live = (det != 0.f);
    if (!live) return;
    const float nom = 1.0f / (det * det);
    const float da = nom * (-cc * cc * dcn[0] + b * cc * dcn[1] + (det - a * cc) * dcn[2]);
    const float db = nom * (2.f * b * cc * dcn[0] - (det + 2.f * b * b) * dcn[1] + 2.f * a * b * dcn[2]);
    const float dc = nom * ((det - a * cc) * dcn[0] + a * b * dcn[1] - a * a * dcn[2]);
    dV[0] = T0[0] * T0[0] * da + T0[0] * T1[0] * db + T1[0] * T1[0] * dc;
    dV[1] = 2.f * T0[0] * T0[1] * da + (T0[0] * T1[1] + T1[0] * T0[1]) * db + 2.f * T1[0] * T1[1] * dc;
    dV[2] = 2.f * T0[0] * T0[2] * da + (T0[0] * T1[2] + T1[0] * T0[2]) * db + 2.f * T1[0] * T1[2] * dc;
    dV[3] = T0[1] * T0[1] * da + T0[1] * T1[1] * db + T1[1] * T1[1] * dc;
    dV[4] = 2.f * T0[1] * T0[2] * da + (T0[1] * T1[2] + T1[1] * T0[2]) * db + 2.f * T1[1] * T1[2] * dc;
    dV[5] = T0[2] * T0[2] * da + T0[2] * T1[2] * db + T1[2] * T1[2] * dc;
    float dT0[3], dT1[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        dT0[k] = 2.f * Ax[k] * da + Ay[k] * db;
        dT1[k] = Ax[k] * db + 2.f * Ay[k] * dc;
    }
    const float dJ00 = c.e[0] * dT0[0] + c.e[1] * dT0[1] + c.e[2] * dT0[2];
    const float dJ20 = c.e[8] * dT0[0] + c.e[9] * dT0[1] + c.e[10] * dT0[2];
    const float dJ11 = c.e[4] * dT1[0] + c.e[5] * dT1[1] + c.e[6] * dT1[2];
    const float dJ21 = c.e[8] * dT1[0] + c.e[9] * dT1[1] + c.e[10] * dT1[2];
    const float dtx = -c.fx * iz2 * dJ20;
    const float dty = -c.fy * iz2 * dJ21;
    const float dtz = -c.fx * iz2 * dJ00 - c.fy * iz2 * dJ11 + 2.f * c.fx * t.x * iz3 * dJ20 + 2.f * c.fy * t.y * iz3 * dJ21;
    dxyz.x = c.e[0] * dtx + c.e[4] * dty + c.e[8] * dtz;
    dxyz.y = c.e[1] * dtx + c.e[5] * dty + c.e[9] * dtz;
    dxyz.z = c.e[2] * dtx + c.e[6] * dty + c.e[10] * dtz;
    cg[0] = iz * dJ00 - t.x * iz2 * dJ20;
    cg[1] = iz * dJ11 - t.y * iz2 * dJ21;
    const float p[3] = {px, py, pz};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        cg[2 + k] = J00 * dT0[k] + p[k] * dtx;
        cg[2 + 4 + k] = J11 * dT1[k] + p[k] * dty;
        cg[2 + 8 + k] = J20 * dT0[k] + J21 * dT1[k] + p[k] * dtz;
    }
    cg[2 + 3] = dtx; cg[2 + 7] = dty; cg[2 + 11] = dtz;
}

__global__ void ewa_bwd_kernel(int P, const float* __restrict__ xyz, const float* __restrict__ cov3d,
                               const float* __restrict__ intr, const float* __restrict__ extr,
                               const int* __restrict__ radius, const float* __restrict__ dL_dconic,
                               float* __restrict__ dL_dxyz, float* __restrict__ dL_dcov3d,
                               float* __restrict__ dL_dintr, float* __restrict__ dL_dextr) {
    __shared__ float red[14 * (kThreads / 32)];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const Cam c = load_cam(intr, extr);
    float cg[14];
#pragma unroll
    for (int k = 0; k < 14; k++) cg[k] = 0.f;
    if (i < P) {
        float dV[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float3 dx = make_float3(0.f, 0.f, 0.f);
        if (radius[i] > 0) {
            float v[6], dcn[3];
#pragma unroll
            for (int k = 0; k < 6; k++) v[k] = cov3d[6 * i + k];
#pragma unroll
            for (int k = 0; k < 3; k++) dcn[k] = dL_dconic[3 * i + k];
            bool live;
            float cgl[14];
#pragma unroll
            for (int k = 0; k < 14; k++) cgl[k] = 0.f;
            ewa_backward(c, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], v, dcn, dV, dx, cgl, live);
            if (live) {
#pragma unroll
                for (int k = 0; k < 14; k++) cg[k] = cgl[k];
            } else {
#pragma unroll
                for (int k = 0; k < 6; k++) dV[k] = 0.f;
                dx = make_float3(0.f, 0.f, 0.f);
            }
        }
        dL_dxyz[3 * i] = dx.x; dL_dxyz[3 * i + 1] = dx.y; dL_dxyz[3 * i + 2] = dx.z;
#pragma unroll
        for (int k = 0; k < 6; k++) dL_dcov3d[6 * i + k] = dV[k];
    }
    if (dL_dintr != nullptr || dL_dextr != nullptr) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
        for (int k = 0; k < 14; k++) {
            float x = cg[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) red[k * nwarps + warp] = x;
        }
        __syncthreads();
        if (threadIdx.x < 14) {
            float s = 0.f;
            for (int w = 0; w < nwarps; w++) s += red[threadIdx.x * nwarps + w];
            if (s != 0.f) {
                if (threadIdx.x < 2) { if (dL_dintr) atomicAdd(dL_dintr + threadIdx.x, s); }
                else if (dL_dextr) atomicAdd(dL_dextr + threadIdx.x - 2, s);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// compute_sh  (msplat/msplat/src/compute_sh.cu:1600-1694), layout shs[P,C,D]
// ---------------------------------------------------------------------------
template <int K>  // K = min(D,16) basis functions in the closed-form block
__global__ void sh_fwd_kernel(int P, int C, int D, int deg, const float* __restrict__ shs,
                              const float* __restrict__ dirs, const uint8_t* __restrict__ visible,
                              float* __restrict__ value) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const bool vis = (visible == nullptr || visible[i]);
    const float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
    float B[K];
    sh_basis<K>(x, y, z, B);
    for (int ch = 0; ch < C; ch++) {
        float acc = 0.f;
        if (vis) {
            const float* s = shs + ((size_t)i * C + ch) * D;
#pragma unroll
            for (int k = 0; k < K; k++) acc = fmaf(B[k], s[k], acc);
            if (D > 16) sh_high_visit(deg, x, y, z, [&](int idx, float b, float, float, float) { acc = fmaf(b, s[idx], acc); });
        }
        value[(size_t)i * C + ch] = acc;
    }
}

template <int K>
__global__ void sh_bwd_kernel(int P, int C, int D, int deg, const float* __restrict__ shs,
                              const float* __restrict__ dirs, const uint8_t* __restrict__ visible,
                              const float* __restrict__ dL_dval, float* __restrict__ dL_dshs,
                              float* __restrict__ dL_ddirs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const bool vis = (visible == nullptr || visible[i]);
    const float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
    float B[K], gx[K], gy[K], gz[K];
    sh_basis<K>(x, y, z, B);
    sh_basis_grad<K>(x, y, z, gx, gy, gz);
    float dx = 0.f, dy = 0.f, dz = 0.f;
    for (int ch = 0; ch < C; ch++) {
        const float* s = shs + ((size_t)i * C + ch) * D;
        float* ds = dL_dshs + ((size_t)i * C + ch) * D;
        if (!vis) {
            for (int k = 0; k < D; k++) ds[k] = 0.f;
            continue;
        }
        const float g = dL_dval[(size_t)i * C + ch];
#pragma unroll
        for (int k = 0; k < K; k++) {
            ds[k] = g * B[k];
            const float sg = s[k] * g;
            dx = fmaf(sg, gx[k], dx); dy = fmaf(sg, gy[k], dy); dz = fmaf(sg, gz[k], dz);
        }
        if (D > 16)
            sh_high_visit(deg, x, y, z, [&](int idx, float b, float bx, float by, float bz) {
                ds[idx] = g * b;
                const float sg = s[idx] * g;
                dx = fmaf(sg, bx, dx); dy = fmaf(sg, by, dy); dz = fmaf(sg, bz, dz);
            });
    }
    dL_ddirs[3 * i] = dx; dL_ddirs[3 * i + 1] = dy; dL_ddirs[3 * i + 2] = dz;
}

// ---------------------------------------------------------------------------
// Fused plugin-level forward: everything MsplatRender.render_iter does per
// Gaussian before binning (pointrix/model/renderer/msplat.py:94-139) in one
// pass.  Reads position/scaling/rotation/opacity/shs[P,K,3] (+ extra feature
// columns), writes the packed blend record [P,S] = {u,v,A,B,C,opacity,f0..},
// depth, radius, tiles.  SH rows are staged per warp through padded shared
// memory with 16-byte cp.async copies so that the 192-byte rows are read with
// fully coalesced requests and evaluated bank-conflict free.
// ---------------------------------------------------------------------------
constexpr int kFThreads = 128;
constexpr int kShPitch = 52;  // words per staged SH row (48 + 4 pad): conflict-free LDS.128

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Stage the first nq float4 of each of this warp's 32 SH rows (row = 48 floats).
__device__ __forceinline__ void stage_sh_rows(float* warp_smem, const float* __restrict__ shs, int g0, int P,
                                              int nq) {
    const int lane = threadIdx.x & 31;
    const int total = 32 * nq;
    for (int f = lane; f < total; f += 32) {
        const int g = f / nq, q = f - g * nq;
        if (g0 + g < P) cp_async16(warp_smem + g * kShPitch + 4 * q, shs + (size_t)(g0 + g) * 48 + 4 * q);
    }
}

// ---- parameter activations folded into the fused kernels (RAW = true) -------------------
// The point cloud stores log-scales, un-normalised quaternions, opacity logits and the SH coefficients as two
// tensors features[P,1,3] / features_rest[P,15,3]; every iteration the reference materialises exp / normalize /
// sigmoid / cat of them before the renderer (pointrix/model/point_cloud/gaussian_points.py:70-86,
// base_model.py:79-85) and their gradients behind it.  With RAW the fused forward reads the raw tensors and
// activates in registers, and the fused backward emits the gradients of the raw tensors.
// expf_acc: accurate exponential (the TU is compiled with --use_fast_math, which maps expf to ex2.approx).
extern "C" __device__ float __nv_expf(float);  // libdevice's accurate expf (what torch.exp evaluates per element)
__device__ __forceinline__ float expf_acc(float x) { return __nv_expf(x); }
__device__ __forceinline__ float sigmoid_acc(float x) { return __fdiv_rn(1.0f, 1.0f + expf_acc(-x)); }
// F.normalize(q, dim = 1): q / max(||q||, 1e-12); returns 1 / max(||q||, eps)
__device__ __forceinline__ float4 normalize_quat(float4 q, float& inv_n) {
    const float n = __fsqrt_rn(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    inv_n = __fdiv_rn(1.0f, fmaxf(n, 1e-12f));
    return make_float4(q.x * inv_n, q.y * inv_n, q.z * inv_n, q.w * inv_n);
}

constexpr int kRestRow = 45;  // floats per features_rest row (15 coefficients x 3 channels)
// Stage this warp's 32 features_rest rows (contiguous: 1440 floats) and 32 features rows (96 floats) with flat
// 16-byte cp.async copies.  Layout in the warp's staging area: rest at [0, 1440), DC at [1440, 1536).
// Lane l then reads row l with scalar loads at stride 45 / 3 words (odd strides: conflict free).
__device__ __forceinline__ void stage_raw_sh(float* warp_smem, const float* __restrict__ features,
                                             const float* __restrict__ features_rest, int g0, int P, int k_active) {
    const int lane = threadIdx.x & 31;
    const int rows = min(32, P - g0);
    if (rows <= 0) return;
    if (k_active > 1) {
        const int n = rows * kRestRow;
        const float* src = features_rest + (size_t)g0 * kRestRow;
        for (int f = lane * 4; f < n; f += 128) {
            if (f + 4 <= n) cp_async16(warp_smem + f, src + f);
            else for (int e = f; e < n; e++) warp_smem[e] = src[e];
        }
    }
    const int n = rows * 3;
    const float* src = features + (size_t)g0 * 3;
    for (int f = lane * 4; f < n; f += 128) {
        if (f + 4 <= n) cp_async16(warp_smem + 32 * kRestRow + f, src + f);
        else for (int e = f; e < n; e++) warp_smem[32 * kRestRow + e] = src[e];
    }
}
// SH coefficient (k, ch) of this lane's Gaussian from the raw staging area
__device__ __forceinline__ float raw_sh(const float* warp_smem, int lane, int k, int ch) {
    return k == 0 ? warp_smem[32 * kRestRow + 3 * lane + ch] : warp_smem[kRestRow * lane + 3 * (k - 1) + ch];
}

template <int KA, bool RAW>  // KA: active SH basis count 1,4,9,16; RAW: inputs are the point cloud's raw parameters
__global__ void __launch_bounds__(kFThreads)
fused_fwd_kernel(int P, const float* __restrict__ pos, const float* __restrict__ scales,
                 const float4* __restrict__ quats, const float* __restrict__ opacity,
                 const float* __restrict__ shs /*[P,16,3]; RAW: features[P,1,3]*/,
                 const float* __restrict__ shs_rest /*RAW: features_rest[P,15,3]*/, const float* __restrict__ extra,
                 int n_extra, int with_depth, const float* __restrict__ intr, const float* __restrict__ extr,
                 const float* __restrict__ cam_center, int W, int H, int gx, int gy, float nearest,
                 float extent, int S, int tight, float* __restrict__ rec, float* __restrict__ depth,
                 int* __restrict__ radius, int* __restrict__ tiles, int2* __restrict__ rect /*nullable*/,
                 unsigned int* __restrict__ vis_cnt /*nullable: visible Gaussians per 1024 ids, accumulated*/) {
    pdl_wait();
    extern __shared__ __align__(16) float sh_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float* wsm = sh_smem + warp * 32 * kShPitch;
    constexpr int NQ = (3 * KA + 3) / 4;
    if (RAW) stage_raw_sh(wsm, shs, shs_rest, blockIdx.x * blockDim.x + warp * 32, P, KA);
    else stage_sh_rows(wsm, shs, blockIdx.x * blockDim.x + warp * 32, P, NQ);

    const Cam c = load_cam(intr, extr);
    float px = 0.f, py = 0.f, pz = 0.f;
    bool keep = false;
    float u = 0.f, v = 0.f, cn[3] = {0.f, 0.f, 0.f};
    int rad = 0, nt = 0;
    float3 t = make_float3(0.f, 0.f, 0.f);
    if (i < P) {
        px = pos[3 * i]; py = pos[3 * i + 1]; pz = pos[3 * i + 2];
        t = cam_transform(c, px, py, pz);
        keep = project_uv(c, t, W, H, nearest, extent, u, v);
        keep = keep && (t.z != 0.f);  // visible = depth != 0 (msplat.py:115)
        if (keep) {
            float4 q = quats[i];
            float sx = scales[3 * i], sy = scales[3 * i + 1], sz = scales[3 * i + 2];
            if (RAW) {  // gaussian_points.py:70-82: scaling = exp, rotation = normalize
                float inv_n;
                q = normalize_quat(q, inv_n);
                sx = expf_acc(sx); sy = expf_acc(sy); sz = expf_acc(sz);
            }
            float cov[6];
            cov3d_from_scale_quat(sx, sy, sz, q.x, q.y, q.z, q.w, cov);
            const EwaT T = ewa_T(c, t);
            float a, b, cc, det;
            ewa_cov2d(T, cov, a, b, cc);
            int r_, t_;
            float cn_[3];
            if (ewa_finish(a, b, cc, u, v, gx, gy, det, r_, t_, cn_)) {
                rad = r_; nt = t_; cn[0] = cn_[0]; cn[1] = cn_[1]; cn[2] = cn_[2];
            }
        }
    }
    // the tile rectangle binning will expand -- computed ONCE, here: the key emission reads it back instead of
    // re-deriving it (tight: only the tiles the alpha >= 1/255 ellipse can reach, common.cuh)
    float op_i = i < P ? opacity[i] : 0.f;
    if (RAW) op_i = sigmoid_acc(op_i);  // gaussian_points.py:66-68
    int2 rc = make_int2(0, 0);
    if (rad > 0) {
        int x0, y0, x1, y1;
        if (tight) tight_tile_rect(u, v, rad, cn[0], cn[1], cn[2], op_i, gx, gy, x0, y0, x1, y1);
        else tile_rect(u, v, rad, gx, gy, x0, y0, x1, y1);
        nt = (x1 - x0) * (y1 - y0);
        if (nt > 0) rc = make_int2(x0 | (y0 << 16), (x1 - x0) | ((y1 - y0) << 16));
    }
    if (vis_cnt != nullptr) {  // a warp's 32 ids share one 1024-id chunk
        const unsigned int vm = __ballot_sync(0xffffffffu, rad > 0 && nt > 0);
        if (lane == 0 && vm) atomicAdd(&vis_cnt[(blockIdx.x * blockDim.x + warp * 32) >> 10], (unsigned int)__popc(vm));
    }
    cp_async_wait_all();
    __syncwarp();
    if (i >= P) return;
    if (rect != nullptr) rect[i] = rc;
    // SH colour: rgb = max(sum + 0.5, 0)   (msplat.py:94-105)
    float rgb[3];
    {
        float dx = px - cam_center[0], dy = py - cam_center[1], dz = pz - cam_center[2];
        const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
        dx *= inv; dy *= inv; dz *= inv;
        float B[KA];
        sh_basis<KA>(dx, dy, dz, B);
        const float* row = wsm + lane * kShPitch;
        float acc[3] = {0.f, 0.f, 0.f};
        float w[4 * NQ];
        if (RAW) {
#pragma unroll
            for (int k = 0; k < KA; k++)
#pragma unroll
                for (int ch = 0; ch < 3; ch++) w[3 * k + ch] = raw_sh(wsm, lane, k, ch);
        } else {
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                const float4 f = *reinterpret_cast<const float4*>(row + 4 * q);
                w[4 * q] = f.x; w[4 * q + 1] = f.y; w[4 * q + 2] = f.z; w[4 * q + 3] = f.w;
            }
        }
#pragma unroll
        for (int k = 0; k < KA; k++) {
            acc[0] = fmaf(B[k], w[3 * k], acc[0]);
            acc[1] = fmaf(B[k], w[3 * k + 1], acc[1]);
            acc[2] = fmaf(B[k], w[3 * k + 2], acc[2]);
        }
#pragma unroll
        for (int ch = 0; ch < 3; ch++) rgb[ch] = fmaxf(acc[ch] + 0.5f, 0.f);
    }
    // culled Gaussians: uv = 0, depth = 0, as project_point leaves them
    const float dep = keep ? t.z : 0.f;
    if (!keep) { u = 0.f; v = 0.f; }
    depth[i] = dep;
    radius[i] = rad;
    tiles[i] = nt;
    float* r = rec + (size_t)i * S;
    *reinterpret_cast<float4*>(r) = make_float4(u, v, cn[0], cn[1]);
    *reinterpret_cast<float4*>(r + 4) = make_float4(cn[2], op_i, rgb[0], rgb[1]);
    // remaining feature columns: rgb[2], [depth], extra[P,n_extra], zero pad
    auto feat = [&](int f) -> float {
        if (f == 2) return rgb[2];
        if (with_depth && f == 3) return dep;
        const int e = f - 3 - with_depth;
        return (e >= 0 && e < n_extra) ? extra[(size_t)i * n_extra + e] : 0.f;
    };
    for (int base = 8; base < S; base += 4)
        *reinterpret_cast<float4*>(r + base) =
            make_float4(feat(base - 6), feat(base - 5), feat(base - 4), feat(base - 3));
}

// Fused plugin-level backward: consumes the packed gradient record
// [P,S] = {du,dv,dA,dB,dC,dop,df0..} accumulated by blend backward and produces
// every parameter gradient of render_iter in one pass:
//   position  <- project.bwd(duv, ddepth) + ewa.bwd(dconic) + normalize.bwd(sh.bwd)
//   scaling, rotation <- cov3d.bwd(ewa.bwd)
//   opacity, shs (clamp- and degree-masked), extra features
//   ndc.grad  = duv * (W/2, H/2)                (msplat/msplat/alpha_blending.py:106-110)
//   camera: dintr[4], dextr[12], dcam_center[3] block-reduced then atomically added.
template <int KA, int MINB, bool RAW>  // MINB: resident CTAs per SM the register allocation aims for; RAW: see above
__global__ void __launch_bounds__(kFThreads, MINB)
fused_bwd_kernel(int P, const float* __restrict__ pos, const float* __restrict__ scales,
                 const float4* __restrict__ quats, const float* __restrict__ opacity /*RAW only: the logits*/,
                 const float* __restrict__ shs, const float* __restrict__ shs_rest /*RAW: features_rest*/, int n_extra,
                 int with_depth,
                 const float* __restrict__ intr, const float* __restrict__ extr,
                 const float* __restrict__ cam_center, int W, int H, int S, const float* __restrict__ depth,
                 const int* __restrict__ radius, const float* __restrict__ grec, float* __restrict__ d_pos,
                 float* __restrict__ d_scales, float4* __restrict__ d_quats, float* __restrict__ d_opacity,
                 float* __restrict__ d_shs /*RAW: d_features[P,3]*/, float* __restrict__ d_shs_rest /*RAW: [P,45]*/,
                 float* __restrict__ d_rgb /*[P,3] or null*/, float* __restrict__ d_extra,
                 float* __restrict__ d_ndc, float* __restrict__ d_cam /*[19]: intr4, extr12, center3 or null*/) {
    pdl_wait();
    extern __shared__ __align__(16) float sh_smem[];
    __shared__ float red[19 * (kFThreads / 32)];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float* wsm = sh_smem + warp * 32 * kShPitch;
    constexpr int NQ = (3 * KA + 3) / 4;
    if (RAW) stage_raw_sh(wsm, shs, shs_rest, blockIdx.x * blockDim.x + warp * 32, P, KA);
    else stage_sh_rows(wsm, shs, blockIdx.x * blockDim.x + warp * 32, P, NQ);
    const Cam c = load_cam(intr, extr);
    float cg[19];
#pragma unroll
    for (int k = 0; k < 19; k++) cg[k] = 0.f;
    float gpos[3] = {0.f, 0.f, 0.f}, ds[3] = {0.f, 0.f, 0.f}, dq[4] = {0.f, 0.f, 0.f, 0.f};
    float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float g2 = 0.f, gdepth = 0.f;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (i < P) {
        px = pos[3 * i]; py = pos[3 * i + 1]; pz = pos[3 * i + 2];
        const float* gr = grec + (size_t)i * S;
        const float4 g0 = *reinterpret_cast<const float4*>(gr);
        const float4 g1 = *reinterpret_cast<const float4*>(gr + 4);
        g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
        const float4 g2v = *reinterpret_cast<const float4*>(gr + 8);
        g2 = g2v.x;
        if (with_depth) gdepth = g2v.y;
        if (d_extra != nullptr) {
            for (int e = 0; e < n_extra; e++) d_extra[(size_t)i * n_extra + e] = gr[6 + 3 + with_depth + e];
        }
        if (RAW) {  // sigmoid': s (1 - s)
            const float sg = sigmoid_acc(opacity[i]);
            d_opacity[i] = g[5] * sg * (1.0f - sg);
        } else {
            d_opacity[i] = g[5];
        }
        d_ndc[2 * i] = g[0] * (0.5f * (float)W);
        d_ndc[2 * i + 1] = g[1] * (0.5f * (float)H);
        const float dep = depth[i];
        if (dep != 0.f) {
            const float3 t = cam_transform(c, px, py, pz);
            // project_point backward (uv and depth paths)
            const float3 dt = project_dt(c, t, g[0], g[1], gdepth);
            gpos[0] = c.e[0] * dt.x + c.e[4] * dt.y + c.e[8] * dt.z;
            gpos[1] = c.e[1] * dt.x + c.e[5] * dt.y + c.e[9] * dt.z;
            gpos[2] = c.e[2] * dt.x + c.e[6] * dt.y + c.e[10] * dt.z;
            const float n1 = 1.0f / t.z;
            cg[0] = t.x * n1 * g[0]; cg[1] = t.y * n1 * g[1]; cg[2] = g[0]; cg[3] = g[1];
            cg[4] = px * dt.x; cg[5] = py * dt.x; cg[6] = pz * dt.x; cg[7] = dt.x;
            cg[8] = px * dt.y; cg[9] = py * dt.y; cg[10] = pz * dt.y; cg[11] = dt.y;
            cg[12] = px * dt.z; cg[13] = py * dt.z; cg[14] = pz * dt.z; cg[15] = dt.z;
            if (radius[i] > 0) {
                float4 q = quats[i];
                float sx = scales[3 * i], sy = scales[3 * i + 1], sz = scales[3 * i + 2];
                float inv_n = 1.0f;
                if (RAW) {
                    q = normalize_quat(q, inv_n);
                    sx = expf_acc(sx); sy = expf_acc(sy); sz = expf_acc(sz);
                }
                float cov[6];
                cov3d_from_scale_quat(sx, sy, sz, q.x, q.y, q.z, q.w, cov);
                const float dcn[3] = {g[2], g[3], g[4]};
                float dV[6], cgl[14];
                float3 dx;
                bool live;
                ewa_backward(c, px, py, pz, cov, dcn, dV, dx, cgl, live);
                if (live) {
                    gpos[0] += dx.x; gpos[1] += dx.y; gpos[2] += dx.z;
                    cg[0] += cgl[0]; cg[1] += cgl[1];
#pragma unroll
                    for (int k = 0; k < 12; k++) cg[4 + k] += cgl[2 + k];
                    cov3d_backward(sx, sy, sz, q.x, q.y, q.z, q.w, dV, ds, dq);
                    if (RAW) {
                        // exp': d/draw = d/ds * s;  normalize': (dq - q <q, dq>) / ||raw||
                        ds[0] *= sx; ds[1] *= sy; ds[2] *= sz;
                        const float dot = q.x * dq[0] + q.y * dq[1] + q.z * dq[2] + q.w * dq[3];
                        dq[0] = (dq[0] - q.x * dot) * inv_n; dq[1] = (dq[1] - q.y * dot) * inv_n;
                        dq[2] = (dq[2] - q.z * dot) * inv_n; dq[3] = (dq[3] - q.w * dot) * inv_n;
                    }
                }
            }
        }
    }
    cp_async_wait_all();
    __syncwarp();
    if (i < P) {
        // SH backward.  The reference evaluates SH for every Gaussian (visible
        // defaults to all ones in msplat.py:104), culled ones simply receive a
        // zero dL/drgb.
        float dx = px - cam_center[0], dy = py - cam_center[1], dz = pz - cam_center[2];
        const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
        dx *= inv; dy *= inv; dz *= inv;
        float B[KA], bx[KA], by[KA], bz[KA];
        sh_basis<KA>(dx, dy, dz, B);
        sh_basis_grad<KA>(dx, dy, dz, bx, by, bz);
        const float* row = wsm + lane * kShPitch;
        float w[4 * NQ];
        if (RAW) {
#pragma unroll
            for (int k = 0; k < KA; k++)
#pragma unroll
                for (int ch = 0; ch < 3; ch++) w[3 * k + ch] = raw_sh(wsm, lane, k, ch);
        } else {
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                const float4 f = *reinterpret_cast<const float4*>(row + 4 * q);
                w[4 * q] = f.x; w[4 * q + 1] = f.y; w[4 * q + 2] = f.z; w[4 * q + 3] = f.w;
            }
        }
        float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < KA; k++) {
            acc[0] = fmaf(B[k], w[3 * k], acc[0]);
            acc[1] = fmaf(B[k], w[3 * k + 1], acc[1]);
            acc[2] = fmaf(B[k], w[3 * k + 2], acc[2]);
        }
        // clamp(min=0) gate: gradient passes where rgb + 0.5 > 0
        const float gr3[3] = {(acc[0] + 0.5f > 0.f) ? g[6] : 0.f, (acc[1] + 0.5f > 0.f) ? g[7] : 0.f,
                              (acc[2] + 0.5f > 0.f) ? g2 : 0.f};
        float ddx = 0.f, ddy = 0.f, ddz = 0.f;
        float o[48];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (k < KA) {
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    o[3 * k + ch] = gr3[ch] * B[k];
                    const float sg = w[3 * k + ch] * gr3[ch];
                    ddx = fmaf(sg, bx[k], ddx); ddy = fmaf(sg, by[k], ddy); ddz = fmaf(sg, bz[k], ddz);
                }
            } else {
                o[3 * k] = o[3 * k + 1] = o[3 * k + 2] = 0.f;
            }
        }
        if (d_rgb != nullptr) {
            // factored form for the data-parallel exchange (exchange.cu, sh_grad_gather_kernel): dL/dshs of
            // this view is the outer product basis(dir) x gated dL/drgb, so only the 3-vector leaves the kernel
            d_rgb[3 * i] = gr3[0]; d_rgb[3 * i + 1] = gr3[1]; d_rgb[3 * i + 2] = gr3[2];
        } else if (RAW) {
            // two outputs: d_features[P,3] directly, d_features_rest[P,45] through the staging area (below)
            d_shs[3 * i] = o[0]; d_shs[3 * i + 1] = o[1]; d_shs[3 * i + 2] = o[2];
        } else {
            float4* dst = reinterpret_cast<float4*>(d_shs + (size_t)i * 48);
#pragma unroll
            for (int q = 0; q < 12; q++) dst[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
        // normalize backward: d = v/|v|
        const float dot = ddx * dx + ddy * dy + ddz * dz;
        const float vx = (ddx - dx * dot) * inv, vy = (ddy - dy * dot) * inv, vz = (ddz - dz * dot) * inv;
        gpos[0] += vx; gpos[1] += vy; gpos[2] += vz;
        cg[16] = -vx; cg[17] = -vy; cg[18] = -vz;
        d_pos[3 * i] = gpos[0]; d_pos[3 * i + 1] = gpos[1]; d_pos[3 * i + 2] = gpos[2];
        d_scales[3 * i] = ds[0]; d_scales[3 * i + 1] = ds[1]; d_scales[3 * i + 2] = ds[2];
        d_quats[i] = make_float4(dq[0], dq[1], dq[2], dq[3]);
        if (RAW && d_rgb == nullptr) {
            // this lane's features_rest gradient row into the staging area (every active lane has read its
            // coefficients out of it by now), conflict free at stride 45
            __syncwarp(__activemask());
#pragma unroll
            for (int j = 0; j < kRestRow; j++) wsm[kRestRow * lane + j] = o[3 + j];
        }
    }
    if (RAW && d_rgb == nullptr) {
        // the warp's 32 rows are contiguous in d_features_rest: flat, coalesced 16-byte stores
        __syncwarp();
        const int g0 = blockIdx.x * blockDim.x + warp * 32;
        const int n = min(32, P - g0) * kRestRow;
        float* dst = d_shs_rest + (size_t)g0 * kRestRow;
        for (int f = lane * 4; f < n; f += 128) {
            if (f + 4 <= n) *reinterpret_cast<float4*>(dst + f) = *reinterpret_cast<const float4*>(wsm + f);
            else for (int e = f; e < n; e++) dst[e] = wsm[e];
        }
    }
    if (d_cam != nullptr) block_reduce_atomic<19>(cg, d_cam, red);
}

// ---------------------------------------------------------------------------------------------------
// SH gradient of a data-parallel step, summed over the views of all ranks WITHOUT exchanging it:
// dL/dshs[g] of one view is the outer product  basis(dir(g, camera)) (x) gated dL/drgb[g]  (fused_bwd_kernel
// above), so instead of all-reducing 48 floats per Gaussian the ranks publish 3 (their d_rgb, in symmetric
// memory) plus their camera centre, and every rank rebuilds  sum_q basis(dir_q) (x) d_rgb_q  itself: this
// kernel reads the peers' d_rgb straight over NVLink (plain peer loads; the sum runs in rank order, so all
// ranks hold identical bits) while it does the math, and writes the finished [P,16,3] gradient locally.
// Bytes over NVLink per Gaussian: 12 (world - 1) in, against 192 * 2 (world - 1) / world for the all-reduce.
struct GatherPeers {
    const float* p[16];
};

template <int KA>
__global__ void __launch_bounds__(256, 3)
sh_grad_gather_kernel(GatherPeers peers, long long rgb_off, long long cam_off, int world, int P,
                      const float* __restrict__ pos, float* __restrict__ d_shs) {
    // Persistent CTAs walk blocks of 256 Gaussians (= 768 consecutive d_rgb floats per peer), 4 peers per
    // round.  Peer loads are NVLink round trips, so they are few and wide (coalesced 16 bytes, staged through
    // shared memory) and those of the NEXT (block, round) are already in flight while this one is computed.
    __shared__ __align__(16) float s_rgb[4][768];
    __shared__ float s_cam[16][3];
    const int tid = threadIdx.x;
    const long long n3 = 3ll * P;
    const int nblk = (P + 255) / 256;
    if (tid < 3 * world) s_cam[tid / 3][tid % 3] = __ldcg(peers.p[tid / 3] + cam_off + tid % 3);

    float4 v[4];
    auto issue = [&](int blk_i, int q0) {
        const long long e = 768ll * blk_i + 4 * tid;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tid < 192 && q0 + u < world) {
                const float* src = peers.p[q0 + u] + rgb_off + e;
                if (e + 3 < n3) {
                    v[u] = __ldcg(reinterpret_cast<const float4*>(src));
                } else {  // tail of the last block
                    if (e < n3) v[u].x = __ldcg(src);
                    if (e + 1 < n3) v[u].y = __ldcg(src + 1);
                    if (e + 2 < n3) v[u].z = __ldcg(src + 2);
                }
            }
        }
    };

    int blk_i = blockIdx.x, q0 = 0;
    if (blk_i < nblk) issue(blk_i, q0);
    float px = 0.f, py = 0.f, pz = 0.f;
    float acc[3 * KA];
    while (blk_i < nblk) {
        if (tid < 192) {
#pragma unroll
            for (int u = 0; u < 4; u++) *reinterpret_cast<float4*>(&s_rgb[u][4 * tid]) = v[u];
        }
        __syncthreads();
        int nb = blk_i, nq = q0 + 4;
        if (nq >= world) { nq = 0; nb = blk_i + gridDim.x; }
        if (nb < nblk) issue(nb, nq);
        const int i = blk_i * 256 + tid;
        if (q0 == 0) {
            if (i < P) { px = pos[3 * i]; py = pos[3 * i + 1]; pz = pos[3 * i + 2]; }
#pragma unroll
            for (int k = 0; k < 3 * KA; k++) acc[k] = 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int q = q0 + u;
            if (q < world) {
                const float r0 = s_rgb[u][3 * tid], r1 = s_rgb[u][3 * tid + 1], r2 = s_rgb[u][3 * tid + 2];
                if (r0 != 0.f || r1 != 0.f || r2 != 0.f) {
                    // same direction arithmetic as fused_bwd_kernel
                    float dx = px - s_cam[q][0], dy = py - s_cam[q][1], dz = pz - s_cam[q][2];
                    const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
                    dx *= inv; dy *= inv; dz *= inv;
                    float B[KA];
                    sh_basis<KA>(dx, dy, dz, B);
#pragma unroll
                    for (int k = 0; k < KA; k++) {
                        acc[3 * k] = fmaf(r0, B[k], acc[3 * k]);
                        acc[3 * k + 1] = fmaf(r1, B[k], acc[3 * k + 1]);
                        acc[3 * k + 2] = fmaf(r2, B[k], acc[3 * k + 2]);
                    }
                }
            }
        }
        if (q0 + 4 >= world && i < P) {
            float4* dst = reinterpret_cast<float4*>(d_shs + (size_t)i * 48);
#pragma unroll
            for (int q = 0; q < 12; q++) {
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; e++) o[e] = (4 * q + e) < 3 * KA ? acc[(4 * q + e) < 3 * KA ? 4 * q + e : 0] : 0.f;
                dst[q] = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
        __syncthreads();
        blk_i = nb; q0 = nq;
    }
}

}  // namespace pxb

using namespace pxb;

// ------------------------------- C ABI -------------------------------------
extern "C" {

int pxb_project_point_forward(int P, const float* xyz, const float* intr, const float* extr, int W, int H,
                              float nearest, float extent, float* uv, float* depth, void* stream) {
    if (P <= 0) return 0;
    project_fwd_kernel<<<blocks_for(P), kThreads, 0, (cudaStream_t)stream>>>(P, xyz, intr, extr, W, H, nearest, extent,
                                                                           (float2*)uv, depth);
    return (int)cudaGetLastError();
}

int pxb_project_point_backward(int P, const float* xyz, const float* intr, const float* extr, const float* depth,
                               const float* dL_duv, const float* dL_ddepth, float* dL_dxyz, float* dL_dintr,
                               float* dL_dextr, void* stream) {
    if (P <= 0) return 0;
    project_bwd_kernel<<<blocks_for(P), kThreads, 0, (cudaStream_t)stream>>>(
        P, xyz, intr, extr, depth, (const float2*)dL_duv, dL_ddepth, dL_dxyz, dL_dintr, dL_dextr);
    return (int)cudaGetLastError();
}

int pxb_compute_cov3d_forward(int P, const float* scales, const float* uquats, const uint8_t* visible,
                              float* cov3d, void* stream) {
    if (P <= 0) return 0;
    cov3d_fwd_kernel<<<blocks_for(P), kThreads, 0, (cudaStream_t)stream>>>(P, scales, (const float4*)uquats, visible,
                                                                         cov3d);
    return (int)cudaGetLastError();
}

int pxb_compute_cov3d_backward(int P, const float* scales, const float* uquats, const uint8_t* visible,
                               const float* dL_dcov3d, float* dL_dscales, float* dL_duquats, void* stream) {
    if (P <= 0) return 0;
    cov3d_bwd_kernel<<<blocks_for(P), kThreads, 0, (cudaStream_t)stream>>>(P, scales, (const float4*)uquats, visible,
                                                                         dL_dcov3d, dL_dscales, (float4*)dL_duquats);
    return (int)cudaGetLastError();
}

int pxb_ewa_project_forward(int P, const float* xyz, const float* cov3d, const float* intr, const float* extr,
                            const float* uv, int W, int H, const uint8_t* visible, float* conic, int* radius,
                            int* tiles, void* stream) {
    if (P <= 0) return 0;
    const int gx = (W + PXB_TILE - 1) / PXB_TILE, gy = (H + PXB_TILE - 1) / PXB_TILE;
    ewa_fwd_kernel<<<blocks_for(P), kThreads, 0, (cudaStream_t)stream>>>(P, xyz, cov3d, intr, extr, (const float2*)uv,
                                                                       gx, gy, visible, conic, radius, tiles);
    return (int)cudaGetLastError();
}

int pxb_ewa_project_backward(int P, const float* xyz, const float* cov3d, const float* intr, const float* extr,
                             const int* radius, const float* dL_dconic, float* dL_dxyz, float* dL_dcov3d,
                             float* dL_dintr, float* dL_dextr, void* stream) {
    if (P <= 0) return 0;
    ewa_bwd_kernel<<<blocks_for(P), kThreads, 0, (cudaStream_t)stream>>>(P, xyz, cov3d, intr, extr, radius, dL_dconic,
                                                                       dL_dxyz, dL_dcov3d, dL_dintr, dL_dextr);
    return (int)cudaGetLastError();
}

static int sh_degree_of(int D) {
    for (int d = 0; d <= 10; d++)
        if ((d + 1) * (d + 1) == D) return d;
    return -1;
}

int pxb_compute_sh_forward(int P, int C, int D, const float* shs, const float* dirs, const uint8_t* visible,
                           float* value, void* stream) {
    if (P <= 0 || C <= 0) return 0;
    const int deg = sh_degree_of(D);
    if (deg < 0) return PXB_ERR_BAD_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int nb = blocks_for(P);
    switch (deg) {
        case 0: sh_fwd_kernel<1><<<nb, kThreads, 0, s>>>(P, C, D, deg, shs, dirs, visible, value); break;
        case 1: sh_fwd_kernel<4><<<nb, kThreads, 0, s>>>(P, C, D, deg, shs, dirs, visible, value); break;
        case 2: sh_fwd_kernel<9><<<nb, kThreads, 0, s>>>(P, C, D, deg, shs, dirs, visible, value); break;
        default: sh_fwd_kernel<16><<<nb, kThreads, 0, s>>>(P, C, D, deg, shs, dirs, visible, value); break;
    }
    return (int)cudaGetLastError();
}

int pxb_compute_sh_backward(int P, int C, int D, const float* shs, const float* dirs, const uint8_t* visible,
                            const float* dL_dval, float* dL_dshs, float* dL_ddirs, void* stream) {
    if (P <= 0 || C <= 0) return 0;
    const int deg = sh_degree_of(D);
    if (deg < 0) return PXB_ERR_BAD_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int nb = blocks_for(P);
    switch (deg) {
        case 0: sh_bwd_kernel<1><<<nb, kThreads, 0, s>>>(P, C, D, deg, shs, dirs, visible, dL_dval, dL_dshs, dL_ddirs); break;
        case 1: sh_bwd_kernel<4><<<nb, kThreads, 0, s>>>(P, C, D, deg, shs, dirs, visible, dL_dval, dL_dshs, dL_ddirs); break;
        case 2: sh_bwd_kernel<9><<<nb, kThreads, 0, s>>>(P, C, D, deg, shs, dirs, visible, dL_dval, dL_dshs, dL_ddirs); break;
        default: sh_bwd_kernel<16><<<nb, kThreads, 0, s>>>(P, C, D, deg, shs, dirs, visible, dL_dval, dL_dshs, dL_ddirs); break;
    }
    return (int)cudaGetLastError();
}

int pxb_init(void) {
    // c_nm normalisation for SH degrees 4..10 (double precision on the host)
    static float h[11][11];
    for (int n = 0; n <= 10; n++)
        for (int m = 0; m <= 10; m++) {
            double v = 0.0;
            if (m <= n) {
                double r = 1.0;  // (n-m)!/(n+m)!
                for (int k = n - m + 1; k <= n + m; k++) r /= (double)k;
                v = sqrt((2.0 * n + 1.0) / (4.0 * 3.14159265358979323846) * r);
                if (m > 0) v *= ((m & 1) ? -1.0 : 1.0) * 1.4142135623730951;
            }
            h[n][m] = (float)v;
        }
    return (int)cudaMemcpyToSymbol(c_sh_norm, h, sizeof(h));
}

int pxb_fused_forward(int P, int sh_degree, const float* pos, const float* scales, const float* quats,
                      const float* opacity, const float* shs, const float* shs_rest, const float* extra, int n_extra,
                      int with_depth, const float* intr, const float* extr, const float* cam_center, int W, int H,
                      float nearest, float extent, int S, int tight, float* rec, float* depth, int* radius, int* tiles,
                      void* stream) {
    return pxb::fused_forward(P, sh_degree, pos, scales, quats, opacity, shs, shs_rest, extra, n_extra, with_depth, intr,
                              extr, cam_center, W, H, nearest, extent, S, tight, rec, depth, radius, tiles, nullptr, nullptr,
                              stream);
}

}  // extern "C"

int pxb::fused_forward(int P, int sh_degree, const float* pos, const float* scales, const float* quats,
                       const float* opacity, const float* shs, const float* shs_rest, const float* extra, int n_extra,
                       int with_depth,
                       const float* intr, const float* extr, const float* cam_center, int W, int H, float nearest,
                       float extent, int S, int tight, float* rec, float* depth, int* radius, int* tiles, int* rect,
                       unsigned int* vis_cnt, void* stream) {
    if (P <= 0) return 0;
    if (sh_degree < 0 || sh_degree > 3) return PXB_ERR_UNSUPPORTED;
    if (S % 4 != 0 || S < 6 + 3 + with_depth + n_extra) return PXB_ERR_BAD_ARG;
    if ((((uintptr_t)shs) | ((uintptr_t)rec) | ((uintptr_t)quats)) & 15) return PXB_ERR_ALIGN;
    const int gx = (W + PXB_TILE - 1) / PXB_TILE, gy = (H + PXB_TILE - 1) / PXB_TILE;
    cudaStream_t s = (cudaStream_t)stream;
    const int nb = blocks_for(P, kFThreads);
    const size_t smem = (size_t)kFThreads * kShPitch * sizeof(float);
    if (gx > 0xffff || gy > 0x7fff) return PXB_ERR_BAD_ARG;  // the packed rectangle holds 16-bit tile coordinates
    const bool raw = shs_rest != nullptr;  // the point cloud's raw parameters: activations happen in the kernel
#define PXB_LAUNCH_FWD_V(KA, RAW)                                                                                        \
    PXB_CUDA_OK(launch_k(fused_fwd_kernel<KA, RAW>, dim3(nb), dim3(kFThreads), smem, s, P, pos, scales,                   \
                         (const float4*)quats, opacity, shs, shs_rest, extra, n_extra, with_depth, intr, extr, cam_center, \
                         W, H, gx, gy, nearest, extent, S, tight, rec, depth, radius, tiles, (int2*)rect, vis_cnt))
#define PXB_LAUNCH_FWD(KA)                   \
    if (raw) { PXB_LAUNCH_FWD_V(KA, true); } \
    else { PXB_LAUNCH_FWD_V(KA, false); }
    switch (sh_degree) {
        case 0: PXB_LAUNCH_FWD(1); break;
        case 1: PXB_LAUNCH_FWD(4); break;
        case 2: PXB_LAUNCH_FWD(9); break;
        default: PXB_LAUNCH_FWD(16); break;
    }
#undef PXB_LAUNCH_FWD
#undef PXB_LAUNCH_FWD_V
    return (int)cudaGetLastError();
}

extern "C" {

int pxb_fused_backward(int P, int sh_degree, const float* pos, const float* scales, const float* quats,
                       const float* opacity_raw, const float* shs, const float* shs_rest, int n_extra, int with_depth,
                       const float* intr, const float* extr, const float* cam_center, int W, int H, int S,
                       const float* depth, const int* radius, const float* grec, float* d_pos, float* d_scales,
                       float* d_quats, float* d_opacity, float* d_shs, float* d_shs_rest, float* d_rgb, float* d_extra,
                       float* d_ndc, float* d_cam, void* stream) {
    if (P <= 0) return 0;
    if (sh_degree < 0 || sh_degree > 3) return PXB_ERR_UNSUPPORTED;
    if (d_shs == nullptr && d_rgb == nullptr) return PXB_ERR_BAD_ARG;
    const bool raw = shs_rest != nullptr;
    if (raw && (opacity_raw == nullptr || (d_rgb == nullptr && d_shs_rest == nullptr))) return PXB_ERR_BAD_ARG;
    if ((((uintptr_t)shs) | ((uintptr_t)shs_rest) | ((uintptr_t)grec) | ((uintptr_t)quats) | ((uintptr_t)d_quats) |
         ((uintptr_t)d_shs_rest) | (raw ? 0 : (uintptr_t)d_shs)) & 15)
        return PXB_ERR_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    const int nb = blocks_for(P, kFThreads);
    const size_t smem = (size_t)kFThreads * kShPitch * sizeof(float);
    // 121 registers (4 CTAs of 128 threads per SM).  Capping at 96 for a fifth CTA was measured: 0.130 vs 0.133 ms,
    // not worth the 52 spilled bytes
#define PXB_LAUNCH_BWD_V(KA, RAW)                                                                                       \
    PXB_CUDA_OK(launch_k(fused_bwd_kernel<KA, 4, RAW>, dim3(nb), dim3(kFThreads), smem, s, P, pos, scales,                \
                         (const float4*)quats, opacity_raw, shs, shs_rest, n_extra, with_depth, intr, extr, cam_center,  \
                         W, H, S, depth, radius, grec, d_pos, d_scales, (float4*)d_quats, d_opacity, d_shs, d_shs_rest,  \
                         d_rgb, d_extra, d_ndc, d_cam))
#define PXB_LAUNCH_BWD(KA)                   \
    if (raw) { PXB_LAUNCH_BWD_V(KA, true); } \
    else { PXB_LAUNCH_BWD_V(KA, false); }
    switch (sh_degree) {
        case 0: PXB_LAUNCH_BWD(1); break;
        case 1: PXB_LAUNCH_BWD(4); break;
        case 2: PXB_LAUNCH_BWD(9); break;
        default: PXB_LAUNCH_BWD(16); break;
    }
#undef PXB_LAUNCH_BWD
#undef PXB_LAUNCH_BWD_V
    return (int)cudaGetLastError();
}

int pxb_sh_grad_gather(const void* const* peer_ptrs, long long rgb_offset, long long cam_offset, int world, int P,
                       int sh_degree, const float* pos, float* d_shs, void* stream) {
    if (peer_ptrs == nullptr || world < 1 || world > 16 || rgb_offset < 0 || cam_offset < 0 || !pos || !d_shs)
        return PXB_ERR_BAD_ARG;
    if (P <= 0) return 0;
    if (sh_degree < 0 || sh_degree > 3) return PXB_ERR_UNSUPPORTED;
    if (((uintptr_t)d_shs) & 15) return PXB_ERR_ALIGN;
    GatherPeers gp;
    for (int q = 0; q < 16; q++) gp.p[q] = q < world ? (const float*)peer_ptrs[q] : nullptr;
    for (int q = 0; q < world; q++)
        if (gp.p[q] == nullptr) return PXB_ERR_BAD_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int nb = (P + 255) / 256;
    if (nb > 148 * 3) nb = 148 * 3;  // persistent: 3 resident CTAs per SM
#define PXB_LAUNCH_GATHER(KA) \
    sh_grad_gather_kernel<KA><<<nb, 256, 0, s>>>(gp, rgb_offset, cam_offset, world, P, pos, d_shs)
    switch (sh_degree) {
        case 0: PXB_LAUNCH_GATHER(1); break;
        case 1: PXB_LAUNCH_GATHER(4); break;
        case 2: PXB_LAUNCH_GATHER(9); break;
        default: PXB_LAUNCH_GATHER(16); break;
    }
#undef PXB_LAUNCH_GATHER
    return (int)cudaGetLastError();
}

}  // extern "C"
