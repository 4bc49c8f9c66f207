// pointrix_b200 -- shared device helpers (sm_100a).
//
// The integer outputs of the render path (radius, tiles, keys, sorted ids, tile
// ranges) must equal the reference msplat's bit for bit.  They are functions of
// fp32 intermediates, so the integer-critical chain below is written with
// explicit single-rounding intrinsics (__fmul_rn/__fadd_rn/__fmaf_rn are never
// re-contracted by nvcc) and the MUFU approximations the reference's
// `--use_fast_math` build issues (rcp/sqrt/ex2 .approx.ftz).  The operation
// order was decoded from the SASS of the reference's sm_100 build
// (re-checked against oracle/_ref/build/*.o) and is
// cited per function against the reference source it restates.
//
// Compile with -ftz=true (the reference's fast-math build flushes denormals on
// every fp32 op).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PXB_TILE 16            // msplat/msplat/include/config.h:7-8
#define PXB_TILE_PIX 256

namespace pxb {

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// FMNMX semantics (NaN-suppressing max), as the reference's max() compiles to.
__device__ __forceinline__ float fmax_nn(float a, float b) {
    float r;
    asm("max.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float fmin_nn(float a, float b) {
    float r;
    asm("min.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// Camera block handed to every per-Gaussian kernel: extr[3x4] row-major, intr
// (fx,fy,cx,cy).  Read from device memory (they may be autograd leaves that
// never visit the host).
struct Cam {
    float e[12];
    float fx, fy, cx, cy;
};
__device__ __forceinline__ Cam load_cam(const float* __restrict__ intr, const float* __restrict__ extr) {
    Cam c;
#pragma unroll
    for (int i = 0; i < 12; i++) c.e[i] = __ldg(extr + i);
    c.fx = __ldg(intr + 0);
    c.fy = __ldg(intr + 1);
    c.cx = __ldg(intr + 2);
    c.cy = __ldg(intr + 3);
    return c;
}

// t = E.[p;1]    (project_point.cu:28-30, ewa_project.cu:35-39)
// SASS: FMUL(py,e1); FFMA(px,e0,.); FFMA(pz,e2,.); FADD(.,e3)
__device__ __forceinline__ float3 cam_transform(const Cam& c, float px, float py, float pz) {
    float3 t;
    t.x = __fadd_rn(__fmaf_rn(pz, c.e[2], __fmaf_rn(px, c.e[0], __fmul_rn(py, c.e[1]))), c.e[3]);
    t.y = __fadd_rn(__fmaf_rn(pz, c.e[6], __fmaf_rn(px, c.e[4], __fmul_rn(py, c.e[5]))), c.e[7]);
    t.z = __fadd_rn(__fmaf_rn(pz, c.e[10], __fmaf_rn(px, c.e[8], __fmul_rn(py, c.e[9]))), c.e[11]);
    return t;
}

// Pixel coordinates + frustum cull (project_point.cu:31-53).
// norm1 is an fp64 reciprocal rounded to fp32; u = fma(norm1, tx*fx, cx) - 0.5.
// Returns false when culled (outputs stay zero in the reference).
__device__ __forceinline__ bool project_uv(const Cam& c, const float3 t, int W, int H, float nearest,
                                           float extent, float& u, float& v) {
    const float n1 = (float)(1.0 / ((double)t.z + 1e-7));
    u = __fadd_rn(__fmaf_rn(n1, __fmul_rn(t.x, c.fx), c.cx), -0.5f);
    v = __fadd_rn(__fmaf_rn(n1, __fmul_rn(t.y, c.fy), c.cy), -0.5f);
    bool cull = false;
    if (nearest > 0.f) cull = (t.z <= nearest);  // NaN is not culled, as in the reference
    if (extent > 0.f) {
        const float om = __fadd_rn(1.0f, -extent), op = __fadd_rn(extent, 1.0f);
        const float Wf = (float)W, Hf = (float)H;
        const float xmin = __fmul_rn(__fmul_rn(om, Wf), 0.5f), xmax = __fmul_rn(__fmul_rn(Wf, op), 0.5f);
        const float ymin = __fmul_rn(__fmul_rn(om, Hf), 0.5f), ymax = __fmul_rn(__fmul_rn(op, Hf), 0.5f);
        cull = cull || (u < xmin) || (u > xmax) || (v < ymin) || (v > ymax);
    }
    return !cull;
}

// Sigma = M^T M, M = S*R(q) in GLM column-major (compute_cov3d.cu:24-57).
// q = (r,x,y,z).  Rotation entries: off-diagonals 2*fma(a,b,+-(c*d)) as t+t,
// diagonals 1 - 2*s with s00 = fadd(y*y, z*z), s11 = fma(x,x,z*z),
// s22 = fma(x,x,y*y).  M[c][k] = s_k*Rg[c][k];
// Sigma[c][r] = fma(M[r][2],M[c][2], fma(M[r][0],M[c][0], M[r][1]*M[c][1])).
__device__ __forceinline__ void cov3d_from_scale_quat(float sx, float sy, float sz, float r, float x, float y,
                                                      float z, float cov[6]) {
    const float xz = __fmul_rn(x, z), zz = __fmul_rn(z, z), rx = __fmul_rn(r, x), rz = __fmul_rn(r, z),
                yy = __fmul_rn(y, y);
    float a;
    // GLM columns c0,c1,c2 (each a vec3 indexed by k)
    float R[3][3];
    a = __fadd_rn(yy, zz);            R[0][0] = __fadd_rn(-__fadd_rn(a, a), 1.0f);
    a = __fmaf_rn(x, y, -rz);         R[0][1] = __fadd_rn(a, a);
    a = __fmaf_rn(r, y, xz);          R[0][2] = __fadd_rn(a, a);
    a = __fmaf_rn(x, y, rz);          R[1][0] = __fadd_rn(a, a);
    a = __fmaf_rn(x, x, zz);          R[1][1] = __fadd_rn(-__fadd_rn(a, a), 1.0f);
    a = __fmaf_rn(y, z, -rx);         R[1][2] = __fadd_rn(a, a);
    a = __fmaf_rn(-r, y, xz);         R[2][0] = __fadd_rn(a, a);
    a = __fmaf_rn(y, z, rx);          R[2][1] = __fadd_rn(a, a);
    a = __fmaf_rn(x, x, yy);          R[2][2] = __fadd_rn(-__fadd_rn(a, a), 1.0f);
    float M[3][3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        M[c][0] = __fmul_rn(sx, R[c][0]);
        M[c][1] = __fmul_rn(sy, R[c][1]);
        M[c][2] = __fmul_rn(sz, R[c][2]);
    }
#define PXB_SIG(c, rr) __fmaf_rn(M[rr][2], M[c][2], __fmaf_rn(M[rr][0], M[c][0], __fmul_rn(M[rr][1], M[c][1])))
    cov[0] = PXB_SIG(0, 0);
    cov[1] = PXB_SIG(0, 1);
    cov[2] = PXB_SIG(0, 2);
    cov[3] = PXB_SIG(1, 1);
    cov[4] = PXB_SIG(1, 2);
    cov[5] = PXB_SIG(2, 2);
#undef PXB_SIG
}

// T = J*W (2x3 useful part), ewa_project.cu:41-50.  Tk0 = T[k][0], Tk1 = T[k][1].
struct EwaT {
    float T00, T10, T20, T01, T11, T21;
    float J00, J11, J20, J21;
};
__device__ __forceinline__ EwaT ewa_T(const Cam& c, const float3 t) {
    EwaT o;
    const float rz = rcp_approx(t.z);
    const float rz2 = rcp_approx(__fmul_rn(t.z, t.z));
    o.J00 = __fmul_rn(c.fx, rz);
    o.J11 = __fmul_rn(c.fy, rz);
    o.J20 = __fmul_rn(__fmul_rn(c.fx, -t.x), rz2);
    o.J21 = __fmul_rn(__fmul_rn(c.fy, -t.y), rz2);
    // T[k][0] = fma(J20, e[8+k], fma(J00, e[k], 0*e[4+k]))
    o.T00 = __fmaf_rn(c.e[8], o.J20, __fmaf_rn(c.e[0], o.J00, __fmul_rn(0.f, c.e[4])));
    o.T10 = __fmaf_rn(c.e[9], o.J20, __fmaf_rn(c.e[1], o.J00, __fmul_rn(0.f, c.e[5])));
    o.T20 = __fmaf_rn(c.e[10], o.J20, __fmaf_rn(c.e[2], o.J00, __fmul_rn(0.f, c.e[6])));
    // T[k][1] = fma(J21, e[8+k], fma(0, e[k], J11*e[4+k]))
    o.T01 = __fmaf_rn(c.e[8], o.J21, __fmaf_rn(0.f, c.e[0], __fmul_rn(c.e[4], o.J11)));
    o.T11 = __fmaf_rn(c.e[9], o.J21, __fmaf_rn(0.f, c.e[1], __fmul_rn(c.e[5], o.J11)));
    o.T21 = __fmaf_rn(c.e[10], o.J21, __fmaf_rn(0.f, c.e[2], __fmul_rn(c.e[6], o.J11)));
    return o;
}

// cov2D = T V T^T + 0.3 I  (ewa_project.cu:52-58); returns (a,b,c).
// A = T*V: columns 0,1 contract as fma(T2,.,fma(T0,.,T1*.)), column 2 as
// fma(T2,.,fma(T1,.,T0*.)) -- that asymmetry is what ptxas emitted for the
// reference and is kept deliberately.
__device__ __forceinline__ void ewa_cov2d(const EwaT& o, const float v[6], float& a, float& b, float& c) {
    const float A0x = __fmaf_rn(o.T20, v[2], __fmaf_rn(o.T00, v[0], __fmul_rn(o.T10, v[1])));
    const float A0y = __fmaf_rn(o.T21, v[2], __fmaf_rn(o.T01, v[0], __fmul_rn(o.T11, v[1])));
    const float A1x = __fmaf_rn(o.T20, v[4], __fmaf_rn(o.T00, v[1], __fmul_rn(o.T10, v[3])));
    const float A1y = __fmaf_rn(o.T21, v[4], __fmaf_rn(o.T01, v[1], __fmul_rn(o.T11, v[3])));
    const float A2x = __fmaf_rn(o.T20, v[5], __fmaf_rn(o.T10, v[4], __fmul_rn(o.T00, v[2])));
    const float A2y = __fmaf_rn(o.T21, v[5], __fmaf_rn(o.T11, v[4], __fmul_rn(o.T01, v[2])));
    const float c00 = __fmaf_rn(o.T20, A2x, __fmaf_rn(o.T00, A0x, __fmul_rn(o.T10, A1x)));
    const float c11 = __fmaf_rn(o.T21, A2y, __fmaf_rn(o.T01, A0y, __fmul_rn(o.T11, A1y)));
    const float c01 = __fmaf_rn(o.T20, A2y, __fmaf_rn(o.T00, A0y, __fmul_rn(o.T10, A1y)));
    a = __fadd_rn(c00, 0.3f);
    c = __fadd_rn(c11, 0.3f);
    b = c01;
}

// Tile rectangle (msplat/msplat/include/utils.h:17-37): float ops in the
// reference's order, truncation toward zero, clamp to [0, grid].
__device__ __forceinline__ void tile_rect(float u, float v, int radius, int gx, int gy, int& x0, int& y0,
                                          int& x1, int& y1) {
    const float rf = (float)radius;
    x0 = min(max(0, __float2int_rz(__fmul_rn(__fadd_rn(u, -rf), 0.0625f))), gx);
    y0 = min(max(0, __float2int_rz(__fmul_rn(__fadd_rn(v, -rf), 0.0625f))), gy);
    x1 = min(max(0, __float2int_rz(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(u, rf), 16.0f), -1.0f), 0.0625f))), gx);
    y1 = min(max(0, __float2int_rz(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(v, rf), 16.0f), -1.0f), 0.0625f))), gy);
}

// Tile rectangle of the fused render path: the reference rectangle intersected with the tiles that
// the alpha >= 1/255 ellipse can reach.  A pixel passes the reference's skip rules only if
// op*exp(-q/2) >= 1/255, q = A dx^2 + 2B dx dy + C dy^2, i.e. q <= 2 ln(255 op); the bounding box of
// that ellipse has half extents sqrt(k C/det), sqrt(k A/det).  k is inflated by 1 % + 0.02 and the box
// by half a pixel, which dwarfs the MUFU approximation error: no tile holding a pixel the exact test
// would blend is dropped, so images and gradients are unchanged while the intersection count falls
// by the ratio of the two boxes.  Tile c holds pixel centres 16c .. 16c+15.
__device__ __forceinline__ void tight_tile_rect(float u, float v, int radius, float A, float B, float C, float op,
                                                int gx, int gy, int& x0, int& y0, int& x1, int& y1) {
    tile_rect(u, v, radius, gx, gy, x0, y0, x1, y1);
    if (op < 1.0f / 255.0f) { x1 = x0; y1 = y0; return; }  // alpha <= op < 1/255 at every pixel
    const float det = A * C - B * B;
    if (!(det > 0.f) || !(A > 0.f) || !(C > 0.f)) return;
    const float k = 2.02f * __logf(255.0f * op) + 0.02f;
    const float inv = 1.0f / det;
    const float hx = sqrtf(k * C * inv) + 0.51f, hy = sqrtf(k * A * inv) + 0.51f;
    if (!(hx == hx) || !(hy == hy) || hx > 1e6f || hy > 1e6f) return;
    const float lim = 1e7f;
    const int cx0 = (int)ceilf(fminf(fmaxf((u - hx - 15.f) * 0.0625f, -lim), lim));
    const int cx1 = (int)floorf(fminf(fmaxf((u + hx) * 0.0625f, -lim), lim)) + 1;
    const int cy0 = (int)ceilf(fminf(fmaxf((v - hy - 15.f) * 0.0625f, -lim), lim));
    const int cy1 = (int)floorf(fminf(fmaxf((v + hy) * 0.0625f, -lim), lim)) + 1;
    x0 = max(x0, cx0); x1 = max(min(x1, cx1), x0);
    y0 = max(y0, cy0); y1 = max(min(y1, cy1), y0);
}

// radius / conic / tiles from (a,b,c) (ewa_project.cu:60-82).  Returns false if
// the Gaussian is dropped (det == 0 or empty rect): outputs stay zero.
__device__ __forceinline__ bool ewa_finish(float a, float b, float c, float u, float v, int gx, int gy,
                                           float& det, int& radius, int& tiles, float conic[3]) {
    det = __fmaf_rn(a, c, -__fmul_rn(b, b));
    if (det == 0.0f) return false;
    const float mid = __fmul_rn(__fadd_rn(a, c), 0.5f);
    const float s = sqrt_approx(fmax_nn(__fmaf_rn(mid, mid, -det), 0.1f));
    const float lam = fmax_nn(__fadd_rn(mid, s), __fadd_rn(mid, -s));
    radius = __float2int_ru(__fmul_rn(sqrt_approx(lam), 3.0f));
    int x0, y0, x1, y1;
    tile_rect(u, v, radius, gx, gy, x0, y0, x1, y1);
    tiles = (x1 - x0) * (y1 - y0);
    if (tiles == 0) return false;
    const float rd = rcp_approx(det);
    conic[0] = __fmul_rn(c, rd);
    conic[1] = __fmul_rn(b, -rd);
    conic[2] = __fmul_rn(a, rd);
    return true;
}

// Gaussian weight exponent (alpha_blending.cu:76-80), SASS order:
// q = fma(dx, dx*A, dy*(dy*C)); power = fma(q, -0.5, -(dy*(dx*B))).
__device__ __forceinline__ float blend_power(float dx, float dy, float A, float B, float C) {
    const float q = __fmaf_rn(dx, __fmul_rn(dx, A), __fmul_rn(dy, __fmul_rn(dy, C)));
    return __fmaf_rn(q, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, B)));
}
__device__ __forceinline__ float blend_G(float power) { return ex2_approx(__fmul_rn(power, 1.4426950216293335f)); }

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

#define PXB_CUDA_OK(call)                  \
    do {                                   \
        cudaError_t _e = (call);           \
        if (_e != cudaSuccess) return (int)_e; \
    } while (0)

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// The render path is a chain of ~25 short kernels per view.  Every kernel of the chain starts with
// pdl_wait() (griddepcontrol.wait: returns once the preceding kernel of the stream has completed and its
// writes are visible) and is launched through launch_k() with programmatic stream serialization, so its
// launch latency and CTA scheduling overlap the tail of its predecessor instead of following it.  A
// kernel launched this way MUST call pdl_wait() before its first global-memory access.  No kernel
// triggers early (griddepcontrol.launch_dependents is not used): correctness never depends on PDL, and
// PXB_PDL=0 in the environment launches everything with plain stream ordering (A/B measurements).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();  // pipeline.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// library-internal (not part of the C ABI), binning.cu: the stages of pxb_bin_prepare / pxb_sort_gaussian
// as the whole-view entry point (pipeline.cu) drives them.
//   bin_clear        zero the control block of the P-level workspace (before the fused forward counts into it)
//   bin_vis_counters the per-1024-Gaussian visible counters inside that block
//   bin_prepare      counted != 0: those counters are already filled; *total_dev is left to the key emission
//   sort_gaussian    publish != 0: the key emission stores the intersection count into *total_dev and into the
//                    pinned, device-mapped host word total_host (nullable); rect: the fused forward's tile
//                    rectangles {x0 | y0 << 16, w | h << 16} (nullable: recomputed from uv / radius)
int fused_forward(int P, int sh_degree, const float* pos, const float* scales, const float* quats, const float* opacity,
                  const float* shs, const float* shs_rest, const float* extra, int n_extra, int with_depth, const float* intr,
                  const float* extr,
                  const float* cam_center, int W, int H, float nearest, float extent, int S, int tight, float* rec,
                  float* depth, int* radius, int* tiles, int* rect /*int2[P], nullable*/, unsigned int* vis_cnt /*nullable*/,
                  void* stream);
int bin_clear(int P, void* ws_p, size_t ws_p_bytes, void* stream);
unsigned int* bin_vis_counters(int P, void* ws_p);
int bin_prepare(int P, const float* depth, const int* radius, const int* tiles, int* total_dev, int counted, void* ws_p,
                size_t ws_p_bytes, void* stream);
int sort_gaussian(int P, long long N, int* total_dev, int publish, int* total_host, const float* uv, int uv_stride,
                  int tight, const int* rect, const float* depth, const int* radius, const int* tiles, int W, int H,
                  int* idx_sorted, int* tile_range, long long* keys_sorted_out, void* ws_p, size_t ws_p_bytes, void* ws_n,
                  size_t ws_n_bytes, void* stream);

}  // namespace pxb
