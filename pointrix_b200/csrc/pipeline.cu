// Whole-view entry points: every stage of MsplatRender.render_iter (pointrix/model/renderer/msplat.py:94-151)
// behind ONE call forward and ONE call backward, so that the host side of a view is two foreign calls
// and the kernels of a view are queued back to back from native code (the small binning kernels are
// 3-10 us each: queued from Python, one call per stage, the GPU waits for the host between them).
//
// The intersection count reaches the host without a memcpy in the stream: the scan kernel stores it
// into a pinned, device-mapped host word the caller polls while the blend kernel is already queued.
#include <stdlib.h>

#include "common.cuh"
#include "pointrix_b200.h"

using namespace pxb;

// PXB_PDL=0 switches programmatic dependent launch off (common.cuh)
bool pxb::pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("PXB_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

namespace {

struct WsRender {
    int* tiles; int* rect; int* total_dev; void* ws_p; size_t ws_p_bytes; void* ws_n; size_t ws_n_bytes;
    size_t total;
};

inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

WsRender carve(void* ws, int P, long long N_cap, int W, int H) {
    WsRender b;
    unsigned char* base = (unsigned char*)ws;
    size_t o = 0;
    b.total_dev = (int*)(base + o); o += 256;
    b.tiles = (int*)(base + o); o += up256((size_t)(P > 0 ? P : 1) * 4);
    b.rect = (int*)(base + o); o += up256((size_t)(P > 0 ? P : 1) * 8);  // int2 per Gaussian
    b.ws_p_bytes = pxb_bin_prepare_workspace_bytes(P);
    b.ws_p = base + o; o += up256(b.ws_p_bytes);
    b.ws_n_bytes = pxb_bin_sort_workspace_bytes(N_cap, W, H);
    b.ws_n = base + o; o += up256(b.ws_n_bytes);
    b.total = o;
    return b;
}

inline int mark(void* const* ev, int i, cudaStream_t s) {
    return ev && ev[i] ? (int)cudaEventRecord((cudaEvent_t)ev[i], s) : 0;  // NULL entry: stage not timed
}

}  // namespace

extern "C" {

size_t pxb_render_workspace_bytes(int P, long long N_cap, int W, int H) {
    return carve(nullptr, P, N_cap > 0 ? N_cap : 1, W, H).total;
}

int pxb_render_forward(int P, int sh_degree, const float* pos, const float* scales, const float* quats,
                       const float* opacity, const float* shs, const float* shs_rest, const float* extra, int n_extra,
                       int with_depth,
                       const float* intr, const float* extr, const float* cam_center, int W, int H, float nearest,
                       float extent, float bg, int S, long long N_cap, float* rec, float* depth, int* radius,
                       int* idx_sorted, int* tile_range, float* final_T, int* ncontrib, float* out, int* total_host,
                       void* ws, size_t ws_bytes, void* const* stage_events, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const int C = 3 + (with_depth ? 1 : 0) + n_extra;
    if (P <= 0 || N_cap <= 0 || W <= 0 || H <= 0) return PXB_ERR_BAD_ARG;
    const WsRender b = carve(ws, P, N_cap, W, H);
    if (ws_bytes < b.total) return PXB_ERR_WORKSPACE;
    int rc;
    // the binning counters the fused forward accumulates into (visible Gaussians per 1024 ids) are zeroed first
    if ((rc = bin_clear(P, b.ws_p, b.ws_p_bytes, stream))) return rc;
    if ((rc = mark(stage_events, 0, s))) return rc;
    rc = fused_forward(P, sh_degree, pos, scales, quats, opacity, shs, shs_rest, extra, n_extra, with_depth, intr, extr,
                       cam_center, W, H, nearest, extent, S, /*tight=*/1, rec, depth, radius, b.tiles, b.rect,
                       bin_vis_counters(P, b.ws_p), stream);
    if (rc) return rc;
    if ((rc = mark(stage_events, 1, s))) return rc;
    rc = bin_prepare(P, depth, radius, b.tiles, b.total_dev, /*counted=*/1, b.ws_p, b.ws_p_bytes, stream);
    if (rc) return rc;
    if ((rc = mark(stage_events, 2, s))) return rc;
    // the key emission publishes the intersection count (device word + the pinned host word the caller polls)
    rc = sort_gaussian(P, N_cap, b.total_dev, /*publish=*/1, total_host, rec, S, /*tight=*/1, b.rect, depth, radius,
                       b.tiles, W, H, idx_sorted, tile_range, nullptr, b.ws_p, b.ws_p_bytes, b.ws_n, b.ws_n_bytes, stream);
    if (rc) return rc;
    if ((rc = mark(stage_events, 3, s))) return rc;
    rc = pxb_blend_forward(rec, S, C, idx_sorted, tile_range, bg, W, H, final_T, ncontrib, out, stream);
    if (rc) return rc;
    return mark(stage_events, 4, s);
}

int pxb_render_backward(int P, int sh_degree, const float* pos, const float* scales, const float* quats,
                        const float* opacity_raw, const float* shs, const float* shs_rest, int n_extra, int with_depth,
                        const float* intr, const float* extr,
                        const float* cam_center, int W, int H, float bg, int S, const float* rec, const float* depth,
                        const int* radius, const int* idx_sorted, const int* tile_range, const float* final_T,
                        const int* ncontrib, const float* dL_dout, float* grec, float* d_pos, float* d_scales,
                        float* d_quats, float* d_opacity, float* d_shs, float* d_shs_rest, float* d_rgb, float* d_extra,
                        float* d_ndc, float* d_cam, void* const* stage_events, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const int C = 3 + (with_depth ? 1 : 0) + n_extra;
    if (P <= 0 || W <= 0 || H <= 0) return PXB_ERR_BAD_ARG;
    int rc;
    PXB_CUDA_OK(cudaMemsetAsync(grec, 0, (size_t)P * S * sizeof(float), s));
    if (d_cam) PXB_CUDA_OK(cudaMemsetAsync(d_cam, 0, 19 * sizeof(float), s));
    if ((rc = mark(stage_events, 0, s))) return rc;
    rc = pxb_blend_backward(rec, S, C, idx_sorted, tile_range, bg, W, H, final_T, ncontrib, dL_dout, grec, stream);
    if (rc) return rc;
    if ((rc = mark(stage_events, 1, s))) return rc;
    rc = pxb_fused_backward(P, sh_degree, pos, scales, quats, opacity_raw, shs, shs_rest, n_extra, with_depth, intr, extr,
                            cam_center, W, H, S, depth, radius, grec, d_pos, d_scales, d_quats, d_opacity, d_shs, d_shs_rest,
                            d_rgb, d_extra, d_ndc, d_cam, stream);
    if (rc) return rc;
    return mark(stage_events, 2, s);
}

}  // extern "C"
