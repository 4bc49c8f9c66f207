// Camera model in front of the render path: CameraModel.extrinsic_matrices / camera_centers
// (pointrix/model/camera/camera_model.py:92-175) for one view -- normalize(qrot) (w first) -> rotation matrix
// (unitquat_to_rotmat, pointrix/utils/pose.py:40-83) -> E = [R | t; 0 0 0 1], centre = -R^T t -- and the
// gradient back into (qrot, tvec).  The reference spends ~12 tiny torch kernels forward and as many backward on
// these seven numbers per view; here it is one single-warp kernel each way, chained to the render kernels by PDL.
#include "common.cuh"
#include "pointrix_b200.h"

namespace pxb {

__device__ __forceinline__ void quat_rotmat(float w, float x, float y, float z, float R[9]) {
    R[0] = x * x - y * y - z * z + w * w; R[1] = 2.f * (x * y - z * w);         R[2] = 2.f * (x * z + y * w);
    R[3] = 2.f * (x * y + z * w);         R[4] = -x * x + y * y - z * z + w * w; R[5] = 2.f * (y * z - x * w);
    R[6] = 2.f * (x * z - y * w);         R[7] = 2.f * (y * z + x * w);         R[8] = -x * x - y * y + z * z + w * w;
}

__global__ void camera_fwd_kernel(const float* __restrict__ qrot, const float* __restrict__ tvec, float* __restrict__ E,
                                  float* __restrict__ center) {
    pdl_wait();
    if (threadIdx.x != 0) return;
    const float n = __fsqrt_rn(qrot[0] * qrot[0] + qrot[1] * qrot[1] + qrot[2] * qrot[2] + qrot[3] * qrot[3]);
    const float inv = __fdiv_rn(1.0f, fmaxf(n, 1e-12f));  // F.normalize(dim = -1)
    float R[9];
    quat_rotmat(qrot[0] * inv, qrot[1] * inv, qrot[2] * inv, qrot[3] * inv, R);
    const float t[3] = {tvec[0], tvec[1], tvec[2]};
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) E[4 * r + c] = R[3 * r + c];
        E[4 * r + 3] = t[r];
    }
    E[12] = 0.f; E[13] = 0.f; E[14] = 0.f; E[15] = 1.f;
    for (int i = 0; i < 3; i++) center[i] = -(R[i] * t[0] + R[3 + i] * t[1] + R[6 + i] * t[2]);  // -R^T t
}

__global__ void camera_bwd_kernel(const float* __restrict__ qrot, const float* __restrict__ tvec,
                                  const float* __restrict__ dE /*[16] or null*/, const float* __restrict__ dcenter /*[3] or null*/,
                                  float* __restrict__ dq, float* __restrict__ dt) {
    pdl_wait();
    if (threadIdx.x != 0) return;
    const float n = __fsqrt_rn(qrot[0] * qrot[0] + qrot[1] * qrot[1] + qrot[2] * qrot[2] + qrot[3] * qrot[3]);
    const float inv = __fdiv_rn(1.0f, fmaxf(n, 1e-12f));
    const float w = qrot[0] * inv, x = qrot[1] * inv, y = qrot[2] * inv, z = qrot[3] * inv;
    float R[9];
    quat_rotmat(w, x, y, z, R);
    const float t[3] = {tvec[0], tvec[1], tvec[2]};
    float dR[9], g_t[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < 9; k++) dR[k] = 0.f;
    if (dE != nullptr)
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 3; c++) dR[3 * r + c] = dE[4 * r + c];
            g_t[r] = dE[4 * r + 3];
        }
    if (dcenter != nullptr)
        for (int j = 0; j < 3; j++) {      // centre_i = -sum_j R[j][i] t[j]
            for (int i = 0; i < 3; i++) {
                dR[3 * j + i] -= t[j] * dcenter[i];
                g_t[j] -= R[3 * j + i] * dcenter[i];
            }
        }
    const float d00 = dR[0], d01 = dR[1], d02 = dR[2], d10 = dR[3], d11 = dR[4], d12 = dR[5], d20 = dR[6], d21 = dR[7], d22 = dR[8];
    float g[4];
    g[0] = 2.f * (w * (d00 + d11 + d22) + z * (d10 - d01) + y * (d02 - d20) + x * (d21 - d12));
    g[1] = 2.f * (x * (d00 - d11 - d22) + y * (d10 + d01) + z * (d20 + d02) + w * (d21 - d12));
    g[2] = 2.f * (y * (-d00 + d11 - d22) + x * (d10 + d01) + z * (d21 + d12) + w * (d02 - d20));
    g[3] = 2.f * (z * (-d00 - d11 + d22) + x * (d20 + d02) + y * (d21 + d12) + w * (d10 - d01));
    const float dot = w * g[0] + x * g[1] + y * g[2] + z * g[3];
    const float q[4] = {w, x, y, z};
    for (int k = 0; k < 4; k++) dq[k] = (g[k] - q[k] * dot) * inv;  // normalize'
    for (int k = 0; k < 3; k++) dt[k] = g_t[k];
}

}  // namespace pxb

using namespace pxb;

extern "C" int pxb_camera_forward(const float* qrot, const float* tvec, float* extrinsic, float* center, void* stream) {
    if (!qrot || !tvec || !extrinsic || !center) return PXB_ERR_BAD_ARG;
    return (int)launch_k(camera_fwd_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, qrot, tvec, extrinsic, center);
}

extern "C" int pxb_camera_backward(const float* qrot, const float* tvec, const float* d_extrinsic, const float* d_center,
                                   float* d_qrot, float* d_tvec, void* stream) {
    if (!qrot || !tvec || !d_qrot || !d_tvec) return PXB_ERR_BAD_ARG;
    return (int)launch_k(camera_bwd_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, qrot, tvec, d_extrinsic, d_center,
                         d_qrot, d_tvec);
}
