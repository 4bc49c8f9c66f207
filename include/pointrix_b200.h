/*
 * pointrix_b200 -- C ABI of the B200-native msplat render path.
 *
 * Drop-in boundary: these entry points are what a binding for the reference's
 * 12 `msplat._C` functions (msplat/msplat/src/ext.cpp:14-25; C++ signatures in
 * msplat/msplat/include/<op>.h) would call, expressed with plain device pointers,
 * sizes and a CUDA stream -- no torch types.  Every function returns 0 on
 * success, a cudaError_t value (>0) on a CUDA failure or a PXB_ERR_* code (<0)
 * on a bad argument; nothing throws across the boundary.
 *
 * Ownership: the caller allocates every buffer (device memory unless stated);
 * the library never allocates, frees or retains pointers.  Calls are
 * asynchronous on `stream` (a cudaStream_t passed as void*).  Thread-safe as
 * long as concurrent calls use distinct streams and workspaces.
 *
 * Layout conventions (all row-major, contiguous, fp32 unless stated):
 *   xyz[P,3] scales[P,3] uquats[P,4] (w,x,y,z) opacity[P,1] intr[4]=(fx,fy,cx,cy)
 *   extr[3,4] uv[P,2] depth[P,1] cov3d[P,6] conic[P,3] radius[P] i32 tiles[P] i32
 *   visible[P] u8 (bool) or NULL = all visible.
 *   Packed blend record rec[P,S]  = {u,v,A,B,C,opacity,f_0..f_{c-1},0-pad}, S = pxb_record_stride(c)
 *   Packed gradient    grec[P,S]  = {du,dv,dA,dB,dC,dopacity,df_0..}; must be zeroed by the caller.
 *   Both must be 16-byte aligned.
 */
#ifndef POINTRIX_B200_H
#define POINTRIX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PXB_ERR_BAD_ARG (-1)
#define PXB_ERR_UNSUPPORTED (-2)
#define PXB_ERR_WORKSPACE (-3)
#define PXB_ERR_ALIGN (-4)
#define PXB_MAX_CHANNELS_PER_PASS 26

/* One-time per-process initialisation (constant tables).  Idempotent. */
int pxb_init(void);

/* ---- projection: replaces projectPointsForward/Backward
 *      (msplat/msplat/include/project_point.h, src/project_point.cu:147-227) ---- */
int pxb_project_point_forward(int P, const float* xyz, const float* intr, const float* extr, int W, int H,
                              float nearest, float extent, float* uv, float* depth, void* stream);
/* dL_dintr[4] / dL_dextr[12] may be NULL (camera does not require grad); when
 * given they must be zeroed by the caller and are accumulated atomically. */
int pxb_project_point_backward(int P, const float* xyz, const float* intr, const float* extr, const float* depth,
                               const float* dL_duv, const float* dL_ddepth, float* dL_dxyz, float* dL_dintr,
                               float* dL_dextr, void* stream);

/* ---- 3D covariance: replaces computeCov3DForward/Backward
 *      (include/compute_cov3d.h, src/compute_cov3d.cu:149-199) ---- */
int pxb_compute_cov3d_forward(int P, const float* scales, const float* uquats, const uint8_t* visible, float* cov3d,
                              void* stream);
int pxb_compute_cov3d_backward(int P, const float* scales, const float* uquats, const uint8_t* visible,
                               const float* dL_dcov3d, float* dL_dscales, float* dL_duquats, void* stream);

/* ---- EWA projection: replaces EWAProjectForward/Backward
 *      (include/ewa_project.h, src/ewa_project.cu:254-344) ---- */
int pxb_ewa_project_forward(int P, const float* xyz, const float* cov3d, const float* intr, const float* extr,
                            const float* uv, int W, int H, const uint8_t* visible, float* conic, int* radius,
                            int* tiles, void* stream);
int pxb_ewa_project_backward(int P, const float* xyz, const float* cov3d, const float* intr, const float* extr,
                             const int* radius, const float* dL_dconic, float* dL_dxyz, float* dL_dcov3d,
                             float* dL_dintr, float* dL_dextr, void* stream);

/* ---- spherical harmonics: replaces computeSHForward/Backward
 *      (include/compute_sh.h, src/compute_sh.cu:1696-1753); shs[P,C,D], D in {1,4,...,121} ---- */
int pxb_compute_sh_forward(int P, int C, int D, const float* shs, const float* dirs, const uint8_t* visible,
                           float* value, void* stream);
int pxb_compute_sh_backward(int P, int C, int D, const float* shs, const float* dirs, const uint8_t* visible,
                            const float* dL_dval, float* dL_dshs, float* dL_ddirs, void* stream);

/* ---- tile binning: replaces torch.cumsum + computeGaussianKey + torch.sort + torch.gather +
 *      computeTileGaussianRange (msplat/msplat/sort_gaussian.py:42-52, include/sort_gaussian.h,
 *      src/sort_gaussian.cu:74-142).  Two calls sharing a P-sized workspace:
 *      pxb_bin_prepare depth-orders the Gaussians and counts the intersections (*total_dev, device int);
 *      pxb_sort_gaussian emits (tile, id) pairs in that order, radix-sorts them by tile and extracts
 *      the ranges.  Output order == the reference's stable sort of tile<<32|depth keys. ---- */
size_t pxb_bin_prepare_workspace_bytes(int P);
size_t pxb_bin_sort_workspace_bytes(long long N_cap, int W, int H);
int pxb_bin_prepare(int P, const float* depth, const int* radius, const int* tiles, int* total_dev, void* ws_p,
                    size_t ws_p_bytes, void* stream);
/* N: the exact intersection count when total_dev == NULL (the caller read *total_dev back), otherwise a
 * capacity: grids/buffers are sized for N, the kernels process min(*total_dev, N) entries and the caller
 * verifies *total_dev <= N afterwards (re-running with more capacity if not) -- no host round trip.
 * uv may be strided (uv_stride floats between Gaussians) so the packed record can be passed.
 * tight != 0 (fused render path only; uv must then be the packed record): emit only the tiles of the
 * reference rectangle that the alpha >= 1/255 ellipse can reach -- `tiles` must hold those counts
 * (pxb_fused_forward with tight != 0).  Images and gradients are unchanged; idx_sorted / tile_range
 * are then NOT the reference's arrays (use tight = 0 for those).
 * idx_sorted[N] i32, tile_range[tiles,2] i32; keys_sorted_out[N] i64 optional (NULL to skip; exact-N mode only). */
int pxb_sort_gaussian(int P, long long N, const int* total_dev, const float* uv, int uv_stride, int tight,
                      const float* depth, const int* radius, const int* tiles, int W, int H, int* idx_sorted, int* tile_range,
                      long long* keys_sorted_out, void* ws_p, size_t ws_p_bytes, void* ws_n, size_t ws_n_bytes,
                      void* stream);

/* ---- alpha blending: replaces alphaBlendingForward/Backward
 *      (include/alpha_blending.h, src/alpha_blending.cu:248-572) ---- */
int pxb_record_stride(int C); /* S for C <= PXB_MAX_CHANNELS_PER_PASS channels, else -1 */
int pxb_pack_records(int P, const float* uv, const float* conic, const float* opacity, const float* feature, int C,
                     int c0, int cn, int S, float* rec, void* stream);
int pxb_unpack_grads(int P, const float* grec, int S, int C, int c0, int cn, int accumulate, float* dL_duv,
                     float* dL_dconic, float* dL_dopacity, float* dL_dfeature, void* stream);
/* out[C,H,W], final_T[H,W], ncontrib[H,W] i32 */
int pxb_blend_forward(const float* rec, int S, int C, const int* idx_sorted, const int* tile_range, float bg, int W,
                      int H, float* final_T, int* ncontrib, float* out, void* stream);
int pxb_blend_backward(const float* rec, int S, int C, const int* idx_sorted, const int* tile_range, float bg, int W,
                       int H, const float* final_T, const int* ncontrib, const float* dL_dout, float* grec,
                       void* stream);
/* Measurement aid (no reference counterpart): counters_dev = device pointer to >= 2 uint64 words, or NULL to
 * switch counting off.  While set, slot 0 / slot 1 accumulate the (8x4 pixel block, Gaussian) candidates the
 * warps of blend forward / backward launches executed (32 pixel-Gaussian pair tests each). */
int pxb_blend_counters(unsigned long long* counters_dev);

/* ---- fused per-Gaussian stages of MsplatRender.render_iter
 *      (pointrix/model/renderer/msplat.py:94-139 forward; its autograd graph backward).
 *      shs[P,16,3] (the point cloud's native layout), sh_degree in 0..3,
 *      features = rgb(3) [+ depth] [+ extra[P,n_extra]].
 *      d_cam[19] = dintr[4], dextr[12], dcamera_center[3] (zeroed by caller) or NULL.
 *      RAW mode (shs_rest != NULL): the inputs are the point cloud's RAW parameters --
 *      scales = log-scales, quats = un-normalised quaternions, opacity = logits, shs = features[P,1,3],
 *      shs_rest = features_rest[P,15,3] -- and exp / normalize / sigmoid / cat
 *      (pointrix/model/point_cloud/gaussian_points.py:70-86) happen inside the kernels; the backward then
 *      needs opacity_raw and writes the gradients OF THE RAW TENSORS (d_shs = d_features[P,3],
 *      d_shs_rest = d_features_rest[P,45]).  shs_rest == NULL: post-activation inputs, as render_iter gets them. ---- */
int pxb_fused_forward(int P, int sh_degree, const float* pos, const float* scales, const float* quats,
                      const float* opacity, const float* shs, const float* shs_rest, const float* extra, int n_extra,
                      int with_depth, const float* intr, const float* extr, const float* cam_center, int W, int H,
                      float nearest, float extent, int S, int tight, float* rec, float* depth, int* radius, int* tiles,
                      void* stream);
int pxb_fused_backward(int P, int sh_degree, const float* pos, const float* scales, const float* quats,
                       const float* opacity_raw, const float* shs, const float* shs_rest, int n_extra, int with_depth,
                       const float* intr, const float* extr, const float* cam_center, int W, int H, int S,
                       const float* depth, const int* radius, const float* grec, float* d_pos, float* d_scales,
                       float* d_quats, float* d_opacity, float* d_shs, float* d_shs_rest, float* d_rgb, float* d_extra,
                       float* d_ndc, float* d_cam, void* stream);

/* ---- one view of MsplatRender.render_iter behind ONE call each way
 *      (pointrix/model/renderer/msplat.py:94-151 and its autograd graph): pxb_fused_forward (tight
 *      tile counts) -> pxb_bin_prepare -> pxb_sort_gaussian (capacity mode) -> pxb_blend_forward, queued
 *      back to back on `stream`.  N_cap: intersection capacity idx_sorted[N_cap] and the workspace are
 *      sized for.  total_host: pinned, device-accessible host int; the caller stores -1 before the
 *      call and reads it after the call returns (spinning while it is -1): a value > N_cap means the
 *      lists were truncated and the call must be repeated with a larger capacity, otherwise the
 *      outputs are exact.  stage_events: NULL, or 5 cudaEvent_t recorded before the first and after each
 *      of the four stages (per-stage timing; a NULL entry is skipped).  ws: pxb_render_workspace_bytes(P, N_cap, W, H) bytes,
 *      256-byte aligned, scratch (nothing in it is needed by the backward).
 *      Saved for the backward: rec[P,S], depth[P], radius[P], idx_sorted, tile_range, final_T, ncontrib. ---- */
size_t pxb_render_workspace_bytes(int P, long long N_cap, int W, int H);
int pxb_render_forward(int P, int sh_degree, const float* pos, const float* scales, const float* quats,
                       const float* opacity, const float* shs, const float* shs_rest, const float* extra, int n_extra,
                       int with_depth, const float* intr, const float* extr, const float* cam_center, int W, int H, float nearest,
                       float extent, float bg, int S, long long N_cap, float* rec, float* depth, int* radius,
                       int* idx_sorted, int* tile_range, float* final_T, int* ncontrib, float* out, int* total_host,
                       void* ws, size_t ws_bytes, void* const* stage_events, void* stream);
/* grec[P,S]: scratch (zeroed inside); d_cam[19] or NULL (zeroed inside); stage_events: NULL or 3 events
 * (before blend backward, between, after the per-Gaussian backward).
 * d_rgb: NULL, or [P,3] receiving the clamp-gated dL/drgb INSTEAD of d_shs (which may then be NULL): the
 * factored form of the SH gradient, d_shs = basis(dir) (x) d_rgb, that pxb_sh_grad_gather sums over the
 * views of a data-parallel step. */
int pxb_render_backward(int P, int sh_degree, const float* pos, const float* scales, const float* quats,
                        const float* opacity_raw, const float* shs, const float* shs_rest, int n_extra, int with_depth,
                        const float* intr, const float* extr, const float* cam_center, int W, int H, float bg, int S,
                        const float* rec, const float* depth, const int* radius, const int* idx_sorted,
                        const int* tile_range, const float* final_T, const int* ncontrib, const float* dL_dout,
                        float* grec, float* d_pos, float* d_scales, float* d_quats, float* d_opacity, float* d_shs,
                        float* d_shs_rest, float* d_rgb, float* d_extra, float* d_ndc, float* d_cam,
                        void* const* stage_events, void* stream);

/* ---- data-parallel gradient exchange (all-reduce(SUM) of the parameter gradients and
 *      ndc.grad, all-reduce(MAX) of radii; the reference's equivalent is batch_size = world on one GPU,
 *      pointrix/model/loss.py:27-46, pointrix/controller/gs.py:274-278, msplat.py:211-212).
 *      In-switch (NVLS) all-reduce of a symmetric buffer [n_f32 floats | n_i32 int32] through its
 *      multicast address mc_ptr: rank r reduces (float SUM, int MAX) and re-broadcasts its 1/world slice.
 *      n_f32 % (4*world) == 0, n_i32 % world == 0.  The caller runs a cross-rank barrier before (all
 *      replicas written) and after (all slices landed); the kernel itself never waits on a peer. ---- */
int pxb_nvls_allreduce(void* mc_ptr, long long n_f32, long long n_i32, int rank, int world, void* stream);
/* Same contract over plain peer mappings (peer_ptrs: HOST array of `world` device pointers, the replicas in
 * rank order; world <= 16): rank r sums slice r over all replicas and stores it into every replica.
 * Preferred for world = 2 and when the group has no multicast support. */
int pxb_p2p_allreduce(const void* const* peer_ptrs, long long n_f32, long long n_i32, int rank, int world,
                      void* stream);
/* SH gradient of a data-parallel step without exchanging it: every rank published, in its symmetric replica
 * (peer_ptrs: HOST array of `world` replica base pointers in rank order), the clamp-gated dL/drgb [P,3] of its
 * view at float offset rgb_offset and its camera centre (3 floats) at float offset cam_offset
 * (pxb_render_backward with d_rgb).  Writes d_shs[P,16,3] = sum over ranks q of basis(dir(pos, centre_q)) (x)
 * d_rgb_q (zero above sh_degree), reading the peers over NVLink inside the kernel; identical bits on all ranks.
 * The caller brackets it with the same cross-rank barriers as the all-reduce. */
int pxb_sh_grad_gather(const void* const* peer_ptrs, long long rgb_offset, long long cam_offset, int world, int P,
                       int sh_degree, const float* pos, float* d_shs, void* stream);

/* ---- photometric loss at the render boundary: replaces l1_loss / l2_loss / ssim
 *      (pointrix/model/loss.py:27-67, 73-117) and their autograd, as called by BaseModel.get_loss_dict
 *      (pointrix/model/base_model.py:113-120).  Images are [B,C,H,W] fp32 contiguous; window 11, sigma 1.5,
 *      zero padding 5 (loss.py:100-108).  ws: pxb_loss_workspace_bytes(B,C,H,W) bytes of scratch.
 *      l1_mean[B], ssim_mean[B]: per-image means over C*H*W (deterministic fixed-order reduction).
 *      dmaps: NULL (no backward will follow) or [3,B,C,H,W] receiving dSSIM/dE[x], dSSIM/dE[xx], dSSIM/dE[xy].
 *      pxb_l1_ssim_loss_forward: the training loss in the same pass, loss3 = {(1-l)*L1 + l*(1-SSIM), L1, 1-SSIM}
 *      with means over the whole batch (base_model.py:117-120).
 *      Backward: d_pred = w_ssim[b] * dSSIM_sum_b/dpred + w_l1[b] * sign(pred-gt) with
 *      w_x[b] = s_x * g_x[b*g_stride]: g_* are DEVICE arrays (NULL = 0; g_stride 0 broadcasts one scalar), s_*
 *      host scales, so the upstream gradient never has to reach the host. ---- */
size_t pxb_loss_workspace_bytes(int B, int C, int H, int W);
int pxb_l1_ssim_forward(int B, int C, int H, int W, const float* pred, const float* gt, float* dmaps, float* l1_mean,
                        float* ssim_mean, void* ws, size_t ws_bytes, void* stream);
int pxb_l1_ssim_loss_forward(int B, int C, int H, int W, const float* pred, const float* gt, float lambda_ssim,
                             float* dmaps, float* loss3, void* ws, size_t ws_bytes, void* stream);
int pxb_l1_ssim_backward(int B, int C, int H, int W, const float* pred, const float* gt, const float* dmaps,
                         const float* g_l1, const float* g_ssim, int g_stride, float s_l1, float s_ssim, float* d_pred,
                         void* stream);
/* mode 1 = |pred-gt| (l1_loss), 2 = (pred-gt)^2 (l2_loss); n = elements per image.  map_out: NULL or [B*n]
 * (return_mean=False); mean_out[B].  ws: at least 4*B*min(ceil(n/256), 1184) bytes of scratch.  Backward:
 * d_pred = (w[b] + g_map[i]) * d loss/d pred, either of w (device [B]) / g_map (device [B*n]) may be NULL. */
int pxb_pixel_loss_forward(int mode, int B, long long n, const float* pred, const float* gt, float* map_out,
                           float* mean_out, void* ws, size_t ws_bytes, void* stream);
int pxb_pixel_loss_backward(int mode, int B, long long n, const float* pred, const float* gt, const float* w,
                            const float* g_map, float* d_pred, void* stream);

/* ---- camera model of one view: replaces CameraModel.extrinsic_matrices / camera_centers
 *      (pointrix/model/camera/camera_model.py:92-175; unitquat_to_rotmat, pointrix/utils/pose.py:40-83).
 *      qrot[4] (w first, any norm: normalised inside), tvec[3] -> extrinsic[4,4] = [R | t; 0 0 0 1], center[3] = -R^T t.
 *      Backward: d_extrinsic[4,4] and / or d_center[3] (either may be NULL) -> d_qrot[4], d_tvec[3]. ---- */
int pxb_camera_forward(const float* qrot, const float* tvec, float* extrinsic, float* center, void* stream);
int pxb_camera_backward(const float* qrot, const float* tvec, const float* d_extrinsic, const float* d_center,
                        float* d_qrot, float* d_tvec, void* stream);

/* ---- optimizer step + densification statistics on the Gaussian table: replaces
 *      torch.optim.Adam.step as driven by BaseOptimizer.update_model (pointrix/optimizer/optimizer.py:128-140;
 *      six parameter groups, eps 1e-15, examples/gaussian_splatting/configs/nerf.yaml:49-69) and the boolean-mask
 *      updates of DensificationController.preprocess (pointrix/controller/gs.py:259-333) -- ONE launch.
 *      A group is a column block of `width` floats per row: columns [param_offset, param_offset + width) of the
 *      [rows, param_stride] tensors param / exp_avg / exp_avg_sq and columns [grad_offset, ..) of a
 *      [rows, grad_stride] gradient.  Dense tensors: stride = width, offset = 0.  The reference's features[P,1,3]
 *      / features_rest[P,15,3] groups take their gradients from columns 0..2 / 3..47 of the fused backward's
 *      dL/dshs[P,48] rows (grad_stride 48); a model that keeps ONE shs[P,16,3] leaf trains the same two column
 *      blocks of it with the two learning rates (param_stride 48 as well).
 *      step: the group's own 1-based Adam step count (bias correction; torch.optim keeps it per parameter and
 *      does not advance it for a parameter without a gradient).  `groups` is a HOST array.
 *      Statistics (P > 0): ndc_grad[P,2] = sum over the batch's views of ndc.grad, radii[P] = max over views;
 *      where radii > 0: grad_accum += || ndc_grad * (sx, sy) ||, acc_steps += 1, max_radii = max(max_radii, radii)
 *      (sx, sy = W/2, H/2 with normalize_grad, gs.py:280-282).  P = 0: Adam only; n_groups = 0: statistics only. ---- */
#define PXB_MAX_ADAM_GROUPS 8
typedef struct pxb_adam_group {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    long long rows;
    int width;
    int param_stride;
    int param_offset;
    int grad_stride;
    int grad_offset;
    int step;
    double lr;
} pxb_adam_group;
int pxb_adam_densify_step(const pxb_adam_group* groups, int n_groups, double beta1, double beta2, double eps,
                          int P, const float* ndc_grad, const int* radii, float sx, float sy, float* grad_accum,
                          float* acc_steps, float* max_radii, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* POINTRIX_B200_H */
