// Tile binning (reference: msplat/msplat/sort_gaussian.py:42-52 and
// msplat/msplat/src/sort_gaussian.cu:17-71).
//
// The reference sorts all N tile intersections by the 64-bit key tile<<32|depth (torch.sort, 8 radix
// passes over N pairs + a gather).  The same order -- ascending (tile, depth bits), ties in
// emission order = ascending Gaussian id -- is produced here with far less traffic:
//   1. stable LSD radix sort of the P Gaussians by depth bits (4 passes over P, P << N);
//   2. inclusive prefix sum of tiles-touched in that order (single pass, decoupled look-back);
//   3. key emission in depth order: (tile id, Gaussian id) pairs, warp-cooperative so every store
//      is coalesced and big Gaussians do not serialise one thread;
//   4. stable radix sort of the N pairs by TILE ID ONLY (ceil(log2 tiles)/8 = 2 passes at 1080p):
//      stability carries the (depth, id) order of step 1 into every tile;
//   5. per-tile [first,last) ranges from the sorted tile ids.
// All radix passes are onesweep style (one histogram pass, then one read + one write of the pairs
// per 8-bit digit with decoupled look-back).  Integer work throughout: idx_sorted and tile_range
// are bit-identical to the reference's; the 64-bit sorted keys can be rebuilt for inspection.
#include "common.cuh"
#include "pointrix_b200.h"

namespace pxb {

// ---------------------------------------------------------------------------
// single-pass inclusive scan (decoupled look-back, warp-parallel), int32, with an optional gather
// ---------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;
#define kFlagAgg (1ull << 32)
#define kFlagPrefix (2ull << 32)

// out[i] = sum_{j<=i} cnt(order[j]),  cnt(g) = radius[g] > 0 ? tiles[g] : 0
__global__ void __launch_bounds__(kScanThreads)
scan_kernel(int P, const int* __restrict__ tiles, const int* __restrict__ radius, const unsigned int* __restrict__ order,
            int* __restrict__ out, int* __restrict__ total, unsigned long long* status, unsigned int* ticket) {
    __shared__ int s_warp[kScanThreads / 32];
    __shared__ int s_tile, s_excl;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = s_tile;
    const int base = tile * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        v[k] = 0;
        if (base + k < P) {
            const int g = order ? (int)order[base + k] : base + k;
            v[k] = (radius == nullptr || radius[g] > 0) ? tiles[g] : 0;
        }
        sum += v[k];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int warp_off = 0, block_sum = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) {
        const int s = s_warp[w];
        if (w < warp) warp_off += s;
        block_sum += s;
    }
    if (warp == 0) {
        // warp-parallel decoupled look-back: 32 predecessors per probe
        volatile unsigned long long* st = status;
        if (lane == 0) st[tile] = (tile == 0 ? kFlagPrefix : kFlagAgg) | (unsigned int)block_sum;
        int excl = 0;
        int t_base = tile - 1;
        while (t_base >= 0) {
            const int idx = t_base - lane;
            const unsigned long long sv = (idx >= 0) ? st[idx] : kFlagPrefix;
            const unsigned int flag = (unsigned int)(sv >> 32);
            const unsigned int zero = __ballot_sync(0xffffffffu, flag == 0);
            const unsigned int pref = __ballot_sync(0xffffffffu, flag == 2);
            const int first = pref ? (__ffs(pref) - 1) : 32;  // nearest tile holding an inclusive prefix
            const unsigned int need = (first < 32) ? ((2u << first) - 1u) : 0xffffffffu;
            if (zero & need) continue;  // some needed predecessor not published yet
            int v_ = (lane <= first) ? (int)(unsigned int)sv : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v_ += __shfl_xor_sync(0xffffffffu, v_, o);
            excl += v_;
            if (first < 32) break;
            t_base -= 32;
        }
        if (lane == 0) {
            if (tile > 0) st[tile] = kFlagPrefix | (unsigned int)(excl + block_sum);
            s_excl = excl;
            if ((tile + 1) * kScanTile >= P) *total = excl + block_sum;
        }
    }
    __syncthreads();
    int run = s_excl + warp_off + (incl - sum);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        run += v[k];
        if (base + k < P) out[base + k] = run;
    }
}

// depth bits + identity permutation: the input of the Gaussian-order sort
__global__ void init_depth_keys_kernel(int P, const float* __restrict__ depth, unsigned int* __restrict__ keys,
                                       unsigned int* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    keys[i] = __float_as_uint(depth[i]);
    vals[i] = (unsigned int)i;
}

// ---------------------------------------------------------------------------
// key emission: sorted position i holds Gaussian g = order[i]; its tile ids go to
// [offs_incl[i] - cnt, offs_incl[i])
// ---------------------------------------------------------------------------
constexpr int kEmitThreads = 256;

__global__ void __launch_bounds__(kEmitThreads)
emit_keys_kernel(int P, const float* __restrict__ uv, int uv_stride, const int* __restrict__ radius,
                 const int* __restrict__ tiles, const unsigned int* __restrict__ order,
                 const int* __restrict__ offs_incl, int gx, int gy, long long N_cap, const int* __restrict__ n_dev,
                 unsigned int* __restrict__ keys, unsigned int* __restrict__ vals) {
    const long long N = n_dev ? min((long long)*n_dev, N_cap) : N_cap;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int x0 = 0, y0 = 0, w = 1, cnt = 0, off = 0, g = 0;
    if (i < P) {
        g = (int)order[i];
        const int r = radius[g];
        if (r > 0) {
            int x1, y1;
            tile_rect(uv[(size_t)g * uv_stride], uv[(size_t)g * uv_stride + 1], r, gx, gy, x0, y0, x1, y1);
            w = max(x1 - x0, 1);
            cnt = tiles[g];  // == w * (y1 - y0) (ewa_project.cu:81); the scan used the same count
            off = offs_incl[i] - cnt;
        }
    }
    // warp-cooperative expansion: slot s of the warp's cnt-sum belongs to the lane whose inclusive
    // prefix first exceeds s
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - cnt;
    for (int s = lane; s < ((total + 31) & ~31); s += 32) {
        int lo = 0;  // binary search over the 32 inclusive prefixes (held one per lane)
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int probe = __shfl_sync(0xffffffffu, incl, lo + step - 1);
            if (probe <= s) lo += step;
        }
        const int src = min(lo, 31);
        const int e_src = __shfl_sync(0xffffffffu, excl, src);
        const int x0s = __shfl_sync(0xffffffffu, x0, src);
        const int y0s = __shfl_sync(0xffffffffu, y0, src);
        const int ws = __shfl_sync(0xffffffffu, w, src);
        const int offs = __shfl_sync(0xffffffffu, off, src);
        const int gs = __shfl_sync(0xffffffffu, g, src);
        if (s < total) {
            const int local = s - e_src;
            const int row = local / ws, col = local - row * ws;
            const long long pos = (long long)offs + local;
            if (pos >= 0 && pos < N) {
                keys[pos] = (unsigned int)((y0s + row) * gx + (x0s + col));
                vals[pos] = (unsigned int)gs;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// onesweep LSD radix sort, 8-bit digits, 32-bit keys + 32-bit values
// ---------------------------------------------------------------------------
constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;  // 4096 pairs per CTA
constexpr int kRadix = 256;
constexpr int kMaxPasses = 4;
#define kStAgg (1u << 30)
#define kStPrefix (2u << 30)
#define kStMask ((1u << 30) - 1u)

__global__ void __launch_bounds__(kRsThreads)
rs_histogram_kernel(const unsigned int* __restrict__ keys, int N_cap, const int* __restrict__ n_dev, int passes,
                    unsigned int* __restrict__ hist) {
    __shared__ unsigned int h[kMaxPasses][kRadix];
    const int N = n_dev ? min(*n_dev, N_cap) : N_cap;
    for (int k = threadIdx.x; k < kMaxPasses * kRadix; k += blockDim.x) (&h[0][0])[k] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const unsigned int k = keys[i];
        for (int p = 0; p < passes; p++) atomicAdd(&h[p][(k >> (8 * p)) & 255u], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < passes * kRadix; k += blockDim.x) {
        const unsigned int c = (&h[0][0])[k];
        if (c) atomicAdd(hist + k, c);
    }
}

// exclusive scan over the 256 bins of every pass (one block per pass)
__global__ void rs_scan_hist_kernel(unsigned int* __restrict__ hist) {
    __shared__ unsigned int s_warp[kRadix / 32];
    unsigned int* h = hist + blockIdx.x * kRadix;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int c = h[threadIdx.x];
    unsigned int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned int off = 0;
    for (int w = 0; w < warp; w++) off += s_warp[w];
    h[threadIdx.x] = off + incl - c;
}

struct RsSmem {
    unsigned int keys[kRsTile];
    unsigned int vals[kRsTile];
    unsigned int warp_hist[kRsThreads / 32][kRadix];
    unsigned int digit_start[kRadix];
    long long gbase[kRadix];
    unsigned int warp_sums[kRadix / 32];
    int tile;
};

__global__ void __launch_bounds__(kRsThreads)
rs_onesweep_kernel(const unsigned int* __restrict__ keys_in, const unsigned int* __restrict__ vals_in,
                   unsigned int* __restrict__ keys_out, unsigned int* __restrict__ vals_out, int N_cap,
                   const int* __restrict__ n_dev, int shift, const unsigned int* __restrict__ bin_base /*[256] exclusive*/,
                   unsigned int* status, unsigned int* ticket) {
    const int N = n_dev ? min(*n_dev, N_cap) : N_cap;
    __shared__ RsSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sm.tile = (int)atomicAdd(ticket, 1u);
    for (int k = tid; k < (kRsThreads / 32) * kRadix; k += kRsThreads) (&sm.warp_hist[0][0])[k] = 0;
    __syncthreads();
    const int tile = sm.tile;
    if ((long long)tile * kRsTile >= N) return;  // capacity-sized grid: surplus CTAs leave
    const long long tile_base = (long long)tile * kRsTile;
    const int tile_n = (int)min((long long)kRsTile, (long long)N - tile_base);

    // warp-striped load: item k of lane l in warp w sits at w*32*IPT + k*32 + l
    unsigned int key[kRsItems];
    unsigned int val[kRsItems];
    unsigned int rank[kRsItems];
    const int wbase = warp * 32 * kRsItems + lane;
#pragma unroll
    for (int k = 0; k < kRsItems; k++) {
        const int p = wbase + k * 32;
        if (p < tile_n) {
            key[k] = keys_in[tile_base + p];
            val[k] = vals_in[tile_base + p];
        } else {
            key[k] = ~0u;
            val[k] = 0;
        }
    }
    // warp-level stable ranking; counters private to the warp.  The peer masks of all items are
    // computed first (independent, they pipeline); only the short leader read-modify-write of the
    // warp-private counter is serial per item.
    const unsigned int lt_mask = (1u << lane) - 1u;
    unsigned int* wh = sm.warp_hist[warp];
    unsigned int peers[kRsItems];
#pragma unroll
    for (int k = 0; k < kRsItems; k++) {
        const bool valid = (wbase + k * 32) < tile_n;
        const unsigned int d = (key[k] >> shift) & 255u;
        const unsigned int vmask = __ballot_sync(0xffffffffu, valid);
        const unsigned int m = __match_any_sync(0xffffffffu, d);  // all lanes participate
        peers[k] = valid ? (m & vmask) : 0u;
    }
#pragma unroll
    for (int k = 0; k < kRsItems; k++) {
        const unsigned int d = (key[k] >> shift) & 255u;
        unsigned int old = 0;
        const int leader = peers[k] ? (__ffs(peers[k]) - 1) : 0;
        if (peers[k] && lane == leader) {
            old = wh[d];
            wh[d] = old + __popc(peers[k]);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[k] = old + __popc(peers[k] & lt_mask);
        __syncwarp();
    }
    __syncthreads();
    // per-digit: exclusive offsets across warps, tile total, publish for look-back
    unsigned int count = 0;
    {
#pragma unroll
        for (int w = 0; w < kRsThreads / 32; w++) {
            const unsigned int c = sm.warp_hist[w][tid];
            sm.warp_hist[w][tid] = count;
            count += c;
        }
        volatile unsigned int* st = status + (size_t)tile * kRadix + tid;
        *st = (tile == 0 ? kStPrefix : kStAgg) | count;
    }
    // exclusive scan of tile totals over digits -> start of each digit inside the sorted tile
    {
        unsigned int incl = count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) sm.warp_sums[warp] = incl;
        __syncthreads();
        unsigned int off = 0;
        for (int w = 0; w < warp; w++) off += sm.warp_sums[w];
        sm.digit_start[tid] = off + incl - count;
    }
    // decoupled look-back: digits are independent, one thread per digit.  Eight predecessors are
    // probed per round trip (independent loads) so that the first wave, where every resident CTA
    // still holds only its aggregate, is walked 8 tiles per L2 latency instead of one.
    {
        unsigned int excl = 0;
        if (tile > 0) {
            const volatile unsigned int* st = status + tid;
            int t = tile - 1;
            bool done_lb = false;
            while (!done_lb) {
                unsigned int sv[8];
#pragma unroll
                for (int u = 0; u < 8; u++) sv[u] = (t - u >= 0) ? st[(size_t)(t - u) * kRadix] : kStPrefix;
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    if (done_lb) break;
                    const unsigned int flag = sv[u] >> 30;
                    if (flag == 0) break;  // not published yet: re-probe from this tile
                    excl += sv[u] & kStMask;
                    t--;
                    if (flag == 2) done_lb = true;
                }
            }
            *((volatile unsigned int*)(status + (size_t)tile * kRadix + tid)) = kStPrefix | (excl + count);
        }
        sm.gbase[tid] = (long long)bin_base[tid] + (long long)excl - (long long)sm.digit_start[tid];
    }
    __syncthreads();
    // scatter into tile-local sorted order
#pragma unroll
    for (int k = 0; k < kRsItems; k++) {
        if ((wbase + k * 32) < tile_n) {
            const unsigned int d = (key[k] >> shift) & 255u;
            const unsigned int pos = sm.digit_start[d] + sm.warp_hist[warp][d] + rank[k];
            sm.keys[pos] = key[k];
            sm.vals[pos] = val[k];
        }
    }
    __syncthreads();
    // coalesced write-out: consecutive sorted positions of one digit are consecutive in global memory
#pragma unroll
    for (int k = 0; k < kRsItems; k++) {
        const int p = k * kRsThreads + tid;
        if (p < tile_n) {
            const unsigned int kk = sm.keys[p];
            const unsigned int d = (kk >> shift) & 255u;
            const long long dst = sm.gbase[d] + p;
            keys_out[dst] = kk;
            vals_out[dst] = sm.vals[p];
        }
    }
}

// ---------------------------------------------------------------------------
// tile ranges (sort_gaussian.cu:45-71) from the sorted tile ids
// ---------------------------------------------------------------------------
__global__ void tile_range_kernel(int N_cap, const int* __restrict__ n_dev, const unsigned int* __restrict__ tile_sorted,
                                  int num_tiles, int2* __restrict__ tile_range) {
    const int N = n_dev ? min(*n_dev, N_cap) : N_cap;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const unsigned int cur = tile_sorted[i];
    if (cur >= (unsigned)num_tiles) return;
    if (i == 0) tile_range[cur].x = 0;
    else {
        const unsigned int prev = tile_sorted[i - 1];
        if (prev != cur) {
            tile_range[cur].x = i;
            if (prev < (unsigned)num_tiles) tile_range[prev].y = i;
        }
    }
    if (i == N - 1) tile_range[cur].y = N;
}

// the reference's sorted int64 keys, rebuilt for inspection / tests
__global__ void rebuild_keys_kernel(int N, const unsigned int* __restrict__ tile_sorted, const int* __restrict__ idx_sorted,
                                    const float* __restrict__ depth, long long* __restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    keys[i] = ((long long)tile_sorted[i] << 32) | (long long)(int)__float_as_uint(depth[idx_sorted[i]]);
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int ceil_log2(int x) {
    int b = 0;
    while ((1ll << b) < x) b++;
    return b;
}
static inline int tile_passes(int num_tiles) { return max(1, (ceil_log2(num_tiles) + 7) / 8); }

// P-sized workspace: survives from pxb_bin_prepare to pxb_sort_gaussian
struct WsP {
    unsigned char* ctl; size_t ctl_bytes;          // zeroed per call
    unsigned long long* scan_status; unsigned int* scan_ticket;
    unsigned int* hist; unsigned int* rs_ticket; unsigned int* rs_status;
    unsigned int* keys[2]; unsigned int* vals[2];  // depth-sort ping-pong; result in keys[0]/vals[0] (4 passes)
    int* offsets;
    size_t total;
};
static WsP carve_p(void* ws, int P) {
    WsP b;
    const size_t scan_tiles = (size_t)(P + kScanTile - 1) / kScanTile + 1;
    const size_t rs_tiles = (size_t)(P + kRsTile - 1) / kRsTile + 1;
    size_t o = 0;
    unsigned char* base = (unsigned char*)ws;
    b.ctl = base;
    b.scan_status = (unsigned long long*)(base + o); o += align_up(scan_tiles * 8, 256);
    b.scan_ticket = (unsigned int*)(base + o); o += 256;
    b.hist = (unsigned int*)(base + o); o += align_up((size_t)4 * kRadix * 4, 256);
    b.rs_ticket = (unsigned int*)(base + o); o += 256;
    b.rs_status = (unsigned int*)(base + o); o += align_up((size_t)4 * rs_tiles * kRadix * 4, 256);
    b.ctl_bytes = o;
    for (int k = 0; k < 2; k++) {
        b.keys[k] = (unsigned int*)(base + o); o += align_up((size_t)P * 4, 256);
        b.vals[k] = (unsigned int*)(base + o); o += align_up((size_t)P * 4, 256);
    }
    b.offsets = (int*)(base + o); o += align_up((size_t)P * 4, 256);
    b.total = o;
    return b;
}
// N-sized workspace of pxb_sort_gaussian
struct WsN {
    unsigned char* ctl; size_t ctl_bytes;
    unsigned int* hist; unsigned int* rs_ticket; unsigned int* rs_status;
    unsigned int* keys[2]; unsigned int* vals_tmp;
    size_t total;
};
static WsN carve_n(void* ws, long long Ncap, int num_tiles) {
    WsN b;
    const int passes = tile_passes(num_tiles);
    const size_t rs_tiles = (size_t)((Ncap + kRsTile - 1) / kRsTile) + 1;
    size_t o = 0;
    unsigned char* base = (unsigned char*)ws;
    b.ctl = base;
    b.hist = (unsigned int*)(base + o); o += align_up((size_t)kMaxPasses * kRadix * 4, 256);
    b.rs_ticket = (unsigned int*)(base + o); o += 256;
    b.rs_status = (unsigned int*)(base + o); o += align_up((size_t)passes * rs_tiles * kRadix * 4, 256);
    b.ctl_bytes = o;
    for (int k = 0; k < 2; k++) { b.keys[k] = (unsigned int*)(base + o); o += align_up((size_t)Ncap * 4, 256); }
    b.vals_tmp = (unsigned int*)(base + o); o += align_up((size_t)Ncap * 4, 256);
    b.total = o;
    return b;
}

// one stable LSD radix sort over `passes` 8-bit digits starting at bit 0; result lands in (k[passes&1], v[passes&1])
static int radix_sort_u32(unsigned int* k[2], unsigned int* v[2], int N_cap, const int* n_dev, int passes,
                          unsigned int* hist, unsigned int* status, unsigned int* ticket, cudaStream_t s) {
    const int rs_tiles = (N_cap + kRsTile - 1) / kRsTile;
    const int hist_blocks = (int)min((long long)(148 * 8), ((long long)N_cap + kRsThreads * 8 - 1) / (kRsThreads * 8));
    rs_histogram_kernel<<<hist_blocks, kRsThreads, 0, s>>>(k[0], N_cap, n_dev, passes, hist);
    rs_scan_hist_kernel<<<passes, kRadix, 0, s>>>(hist);
    for (int p = 0; p < passes; p++) {
        rs_onesweep_kernel<<<rs_tiles, kRsThreads, 0, s>>>(k[p & 1], v[p & 1], k[(p + 1) & 1], v[(p + 1) & 1], N_cap, n_dev,
                                                          8 * p, hist + p * kRadix,
                                                          status + (size_t)p * (rs_tiles + 1) * kRadix, ticket + p);
    }
    return (int)cudaGetLastError();
}

}  // namespace pxb

using namespace pxb;

extern "C" {

size_t pxb_bin_prepare_workspace_bytes(int P) { return carve_p(nullptr, P > 0 ? P : 1).total; }

size_t pxb_bin_sort_workspace_bytes(long long N_cap, int W, int H) {
    const int gx = (W + PXB_TILE - 1) / PXB_TILE, gy = (H + PXB_TILE - 1) / PXB_TILE;
    return carve_n(nullptr, N_cap > 0 ? N_cap : 1, gx * gy).total;
}

// Depth-order the Gaussians and prefix-sum their tile counts in that order.
// *total_dev = number of intersections N.  ws_p must stay untouched until pxb_sort_gaussian.
int pxb_bin_prepare(int P, const float* depth, const int* radius, const int* tiles, int* total_dev, void* ws_p,
                    size_t ws_p_bytes, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P <= 0) return (int)cudaMemsetAsync(total_dev, 0, sizeof(int), s);
    WsP b = carve_p(ws_p, P);
    if (ws_p_bytes < b.total) return PXB_ERR_WORKSPACE;
    PXB_CUDA_OK(cudaMemsetAsync(b.ctl, 0, b.ctl_bytes, s));
    init_depth_keys_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, depth, b.keys[0], b.vals[0]);
    int rc = radix_sort_u32(b.keys, b.vals, P, nullptr, 4, b.hist, b.rs_status, b.rs_ticket, s);  // -> keys[0]/vals[0]
    if (rc) return rc;
    scan_kernel<<<(P + kScanTile - 1) / kScanTile, kScanThreads, 0, s>>>(P, tiles, radius, b.vals[0], b.offsets, total_dev,
                                                                        b.scan_status, b.scan_ticket);
    return (int)cudaGetLastError();
}

// keys + tile sort + ranges.  N: exact count when total_dev == NULL, else a capacity: the kernels
// then process min(*total_dev, N) intersections and the caller verifies *total_dev <= N afterwards.
int pxb_sort_gaussian(int P, long long N, const int* total_dev, const float* uv, int uv_stride, const float* depth,
                      const int* radius, const int* tiles, int W, int H, int* idx_sorted, int* tile_range,
                      long long* keys_sorted_out /*nullable, [N]*/, void* ws_p, size_t ws_p_bytes, void* ws_n,
                      size_t ws_n_bytes, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const int gx = (W + PXB_TILE - 1) / PXB_TILE, gy = (H + PXB_TILE - 1) / PXB_TILE;
    const int num_tiles = gx * gy;
    PXB_CUDA_OK(cudaMemsetAsync(tile_range, 0, (size_t)num_tiles * 2 * sizeof(int), s));
    if (P <= 0 || N <= 0) return 0;
    if (N > 0x3fffffffll) return PXB_ERR_BAD_ARG;
    WsP bp = carve_p(ws_p, P);
    WsN bn = carve_n(ws_n, N, num_tiles);
    if (ws_p_bytes < bp.total || ws_n_bytes < bn.total) return PXB_ERR_WORKSPACE;
    const int passes = tile_passes(num_tiles);
    PXB_CUDA_OK(cudaMemsetAsync(bn.ctl, 0, bn.ctl_bytes, s));
    // ping-pong so that the last pass writes the values straight into idx_sorted
    unsigned int* k[2] = {bn.keys[0], bn.keys[1]};
    unsigned int* v[2];
    v[passes & 1] = (unsigned int*)idx_sorted;
    v[(passes + 1) & 1] = bn.vals_tmp;
    emit_keys_kernel<<<(P + kEmitThreads - 1) / kEmitThreads, kEmitThreads, 0, s>>>(
        P, uv, uv_stride, radius, tiles, bp.vals[0], bp.offsets, gx, gy, N, total_dev, k[0], v[0]);
    int rc = radix_sort_u32(k, v, (int)N, total_dev, passes, bn.hist, bn.rs_status, bn.rs_ticket, s);
    if (rc) return rc;
    const unsigned int* tile_sorted = k[passes & 1];
    tile_range_kernel<<<(int)((N + 255) / 256), 256, 0, s>>>((int)N, total_dev, tile_sorted, num_tiles, (int2*)tile_range);
    if (keys_sorted_out && total_dev == nullptr)
        rebuild_keys_kernel<<<(int)((N + 255) / 256), 256, 0, s>>>((int)N, tile_sorted, idx_sorted, depth, keys_sorted_out);
    return (int)cudaGetLastError();
}

}  // extern "C"
