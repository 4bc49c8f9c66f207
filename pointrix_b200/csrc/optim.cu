// The step right after the render backward, ONE launch for the whole Gaussian table:
//   * Adam on every parameter group (the reference drives torch.optim.Adam through BaseOptimizer.update_model,
//     pointrix/optimizer/optimizer.py:128-140, six groups with their own learning rates and eps = 1e-15,
//     examples/gaussian_splatting/configs/nerf.yaml:49-69), and
//   * the densification statistics of DensificationController.preprocess (pointrix/controller/gs.py:259-333):
//     grad_accum += || ndc.grad * (W/2, H/2) ||, acc_steps += 1, max_radii = max(max_radii, radii), all three
//     only where the Gaussian was visible (radii > 0) in the batch.
// The reference spends ~6 elementwise kernels per group (multi-tensor Adam) plus ~10 boolean-mask scatter
// kernels on these P-sized streams; here every float of the table is read once and written once.
// A group's gradient may live inside a wider row of another tensor (features / features_rest are columns
// 0..2 / 3..47 of the fused backward's dL/dshs[P,16,3] row), so no gradient is ever re-packed.
#include <math.h>

#include "common.cuh"
#include "pointrix_b200.h"

namespace pxb {

constexpr int kOptThreads = 256;
constexpr int kOptItems = 4;  // elements per thread, strided by the CTA so every access is coalesced
constexpr int kOptChunk = kOptThreads * kOptItems;

struct OptGroup {
    float* p;
    const float* g;
    float* m;
    float* v;
    long long n;       // rows * width elements
    int width;         // floats per row of this group
    int p_stride;      // floats per row of the parameter / moment tensors, first column of the group in them
    int p_off;
    int g_stride;      // the same for the gradient tensor
    int g_off;
    float step_size;      // lr / (1 - beta1^t), t = this group's own step count (torch.optim keeps it per parameter)
    float inv_bc2_sqrt;   // 1 / sqrt(1 - beta2^t)
    int cta_begin;        // first CTA of this group
};
struct OptArgs {
    OptGroup grp[PXB_MAX_ADAM_GROUPS];
    int n_groups;
    float beta2, omb1, omb2, eps;  // omb = 1 - beta, rounded from double
    // densification statistics (P == 0: none)
    int P;
    int stats_cta_begin;
    const float* ndc_grad;  // [P,2], summed over the views of the batch
    const int* radii;       // [P], max over the views
    float sx, sy;           // W/2, H/2 when normalize_grad, else 1
    float* grad_accum;      // [P]
    float* acc_steps;       // [P]
    float* max_radii;       // [P]
};

__global__ void __launch_bounds__(kOptThreads) adam_densify_kernel(const OptArgs a) {
    pdl_wait();
    const int cta = blockIdx.x;
    if (a.P > 0 && cta >= a.stats_cta_begin) {
        // DensificationController.preprocess, gs.py:316-333 (selected_points = visibility = radii > 0)
        const long long base = (long long)(cta - a.stats_cta_begin) * kOptChunk;
#pragma unroll
        for (int k = 0; k < kOptItems; k++) {
            const long long i = base + k * kOptThreads + threadIdx.x;
            if (i < a.P) {
                const int r = a.radii[i];
                if (r > 0) {
                    const float2 g = reinterpret_cast<const float2*>(a.ndc_grad)[i];
                    const float gx = g.x * a.sx, gy = g.y * a.sy;  // accumulate_viewspace_grad, gs.py:280-282
                    a.grad_accum[i] += __fsqrt_rn(__fmaf_rn(gx, gx, gy * gy));
                    a.acc_steps[i] += 1.0f;
                    a.max_radii[i] = fmaxf(a.max_radii[i], (float)r);
                }
            }
        }
        return;
    }
    int gi = 0;
#pragma unroll
    for (int k = 1; k < PXB_MAX_ADAM_GROUPS; k++)
        if (k < a.n_groups && cta >= a.grp[k].cta_begin) gi = k;
    const OptGroup& G = a.grp[gi];
    const long long base = (long long)(cta - G.cta_begin) * kOptChunk;
    const bool dense = (G.g_stride == G.width && G.p_stride == G.width);
    float g[kOptItems], m[kOptItems], v[kOptItems], p[kOptItems];
    long long pe[kOptItems];
#pragma unroll
    for (int k = 0; k < kOptItems; k++) {
        const long long e = base + k * kOptThreads + threadIdx.x;
        if (e < G.n) {
            long long ge = e;
            pe[k] = e;
            if (!dense) {
                const long long r = e / G.width;
                const long long c = e - r * G.width;
                ge = r * G.g_stride + G.g_off + c;
                pe[k] = r * G.p_stride + G.p_off + c;
            }
            g[k] = G.g[ge];
            m[k] = G.m[pe[k]];
            v[k] = G.v[pe[k]];
            p[k] = G.p[pe[k]];
        }
    }
#pragma unroll
    for (int k = 0; k < kOptItems; k++) {
        const long long e = base + k * kOptThreads + threadIdx.x;
        if (e < G.n) {
            // torch.optim.Adam (single tensor, no amsgrad / weight decay):
            //   exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
            //   denom = exp_avg_sq.sqrt() / sqrt(1 - beta2^t) + eps;  param.addcdiv_(exp_avg, denom, value = -lr / (1 - beta1^t))
            const float mk = __fmaf_rn(a.omb1, g[k] - m[k], m[k]);
            const float vk = __fmaf_rn(a.omb2 * g[k], g[k], v[k] * a.beta2);
            const float denom = __fmaf_rn(__fsqrt_rn(vk), G.inv_bc2_sqrt, a.eps);
            G.m[pe[k]] = mk;
            G.v[pe[k]] = vk;
            G.p[pe[k]] = __fmaf_rn(-G.step_size, __fdiv_rn(mk, denom), p[k]);
        }
    }
}

}  // namespace pxb

using namespace pxb;

extern "C" int pxb_adam_densify_step(const pxb_adam_group* groups, int n_groups, double beta1, double beta2, double eps,
                                     int P, const float* ndc_grad, const int* radii, float sx, float sy,
                                     float* grad_accum, float* acc_steps, float* max_radii, void* stream) {
    if (n_groups < 0 || n_groups > PXB_MAX_ADAM_GROUPS || (n_groups > 0 && groups == nullptr)) return PXB_ERR_BAD_ARG;
    if (P > 0 && (!ndc_grad || !radii || !grad_accum || !acc_steps || !max_radii)) return PXB_ERR_BAD_ARG;
    if (P > 0 && (((uintptr_t)ndc_grad) & 7)) return PXB_ERR_ALIGN;
    OptArgs a = {};
    a.n_groups = n_groups;
    a.beta2 = (float)beta2;
    a.omb1 = (float)(1.0 - beta1);
    a.omb2 = (float)(1.0 - beta2);
    a.eps = (float)eps;
    long long cta = 0;
    for (int k = 0; k < n_groups; k++) {
        const pxb_adam_group& s = groups[k];
        if (s.step < 1 || s.rows < 0 || s.width <= 0 || s.grad_stride < s.grad_offset + s.width || s.grad_offset < 0 ||
            s.param_stride < s.param_offset + s.width || s.param_offset < 0)
            return PXB_ERR_BAD_ARG;
        if (s.rows > 0 && (!s.param || !s.grad || !s.exp_avg || !s.exp_avg_sq)) return PXB_ERR_BAD_ARG;
        OptGroup& G = a.grp[k];
        G.p = s.param; G.g = s.grad; G.m = s.exp_avg; G.v = s.exp_avg_sq;
        G.n = s.rows * s.width;
        G.width = s.width; G.p_stride = s.param_stride; G.p_off = s.param_offset;
        G.g_stride = s.grad_stride; G.g_off = s.grad_offset;
        const double bc1 = 1.0 - pow(beta1, (double)s.step), bc2 = 1.0 - pow(beta2, (double)s.step);
        G.step_size = (float)(s.lr / bc1);
        G.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
        G.cta_begin = (int)cta;
        cta += (G.n + kOptChunk - 1) / kOptChunk;
    }
    a.P = P > 0 ? P : 0;
    a.stats_cta_begin = (int)cta;
    if (P > 0) {
        a.ndc_grad = ndc_grad; a.radii = radii; a.sx = sx; a.sy = sy;
        a.grad_accum = grad_accum; a.acc_steps = acc_steps; a.max_radii = max_radii;
        cta += ((long long)P + kOptChunk - 1) / kOptChunk;
    }
    if (cta == 0) return 0;
    if (cta > 0x7fffffffll) return PXB_ERR_UNSUPPORTED;
    return (int)launch_k(adam_densify_kernel, dim3((unsigned)cta), dim3(kOptThreads), 0, (cudaStream_t)stream, a);
}
