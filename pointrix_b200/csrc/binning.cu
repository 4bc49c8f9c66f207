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
// Radix passes are chain-free (per-tile digit histogram, scan over tiles, scatter; digits of <= 8 bits,
// the significant bits split evenly over the passes).  Integer work throughout: idx_sorted and tile_range
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
            int* __restrict__ out, int* __restrict__ total, int* __restrict__ total_host, unsigned long long* status,
            unsigned int* ticket) {
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
            if ((tile + 1) * kScanTile >= P) {
                *total = excl + block_sum;
                if (total_host) {  // pinned host word the caller polls (no memcpy in the stream)
                    *reinterpret_cast<volatile int*>(total_host) = excl + block_sum;
                    __threadfence_system();
                }
            }
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
                 const int* __restrict__ offs_incl, int gx, int gy, int tight, long long N_cap,
                 const int* __restrict__ n_dev, unsigned int* __restrict__ keys, unsigned int* __restrict__ vals) {
    const long long N = n_dev ? min((long long)*n_dev, N_cap) : N_cap;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int x0 = 0, y0 = 0, w = 1, cnt = 0, off = 0, g = 0;
    if (i < P) {
        g = (int)order[i];
        const int r = radius[g];
        if (r > 0) {
            int x1, y1;
            const float* rg = uv + (size_t)g * uv_stride;
            if (tight)  // uv is the packed blend record {u,v,A,B,C,op,...}: the fused path's rectangle
                tight_tile_rect(rg[0], rg[1], r, rg[2], rg[3], rg[4], rg[5], gx, gy, x0, y0, x1, y1);
            else
                tile_rect(rg[0], rg[1], r, gx, gy, x0, y0, x1, y1);
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
// LSD radix sort, digits of <= 8 bits, 32-bit keys + 32-bit values.
// Each pass is three chain-free kernels:
//   rs_tile_hist  : digit histogram of every 4096-pair tile            -> counts[digit][tile]
//   rs_tile_scan  : exclusive scan over tiles, per digit (+ digit totals)
//   rs_scatter    : stable in-CTA ranking, shared-memory exchange, coalesced write-out at
//                   base(digit) + counts_excl[digit][tile]
// A single-pass ("onesweep") variant with decoupled look-back was measured first: with ~600 resident
// CTAs every wave restarts the look-back chain (tile t needs the aggregates of all unfinished
// predecessors), which cost 30+ us per wave -- 5x the data movement time.  Re-reading the keys once
// per pass (4 B/pair) removes every inter-CTA dependency.
// ---------------------------------------------------------------------------
constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;  // 4096 pairs per CTA
constexpr int kRadix = 256;
constexpr int kMaxPasses = 4;

// digit layout of one radix sort: pass p ranks bits [shift[p], shift[p]+bits[p]) (bits <= 8)
struct RsDigits {
    int passes;
    int shift[kMaxPasses];
    int bits[kMaxPasses];
};
// `total_bits` significant key bits split into ceil(total/8) passes of near-equal width
static RsDigits make_digits(int total_bits) {
    RsDigits d;
    d.passes = max(1, (total_bits + 7) / 8);
    int done = 0;
    for (int p = 0; p < kMaxPasses; p++) {
        const int left = d.passes - p;
        const int b = p < d.passes ? (total_bits - done + left - 1) / left : 0;
        d.shift[p] = done;
        d.bits[p] = max(b, p < d.passes ? 1 : 0);
        done += b;
    }
    return d;
}

template <int NBITS>
__global__ void __launch_bounds__(kRsThreads)
rs_tile_hist_kernel(const unsigned int* __restrict__ keys, int N_cap, const int* __restrict__ n_dev, int shift, int T,
                    unsigned int* __restrict__ counts /*[2^NBITS][T]*/) {
    constexpr int NB = 1 << NBITS;
    constexpr int NW = kRsThreads / 32;
    __shared__ unsigned int h[NW][NB];  // warp-private: 8x less contention on skewed digits
    const int N = n_dev ? min(*n_dev, N_cap) : N_cap;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int k = tid; k < NW * NB; k += kRsThreads) (&h[0][0])[k] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * kRsTile;
    unsigned int key[kRsItems];
#pragma unroll
    for (int k = 0; k < kRsItems; k++) {
        const long long p = base + k * kRsThreads + tid;
        key[k] = p < N ? keys[p] : 0u;
    }
#pragma unroll
    for (int k = 0; k < kRsItems; k++)
        if (base + k * kRsThreads + tid < N) atomicAdd(&h[warp][(key[k] >> shift) & (NB - 1)], 1u);
    __syncthreads();
    for (int k = tid; k < NB; k += kRsThreads) {
        unsigned int c = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) c += h[w][k];
        counts[(size_t)k * T + blockIdx.x] = c;
    }
}

// one CTA per digit: exclusive scan of counts[digit][0..T) in place, totals[digit] = row sum
__global__ void __launch_bounds__(1024) rs_tile_scan_kernel(unsigned int* __restrict__ counts, int T,
                                                            unsigned int* __restrict__ totals) {
    __shared__ unsigned int s_warp[32];
    __shared__ unsigned int s_total;
    unsigned int* row = counts + (size_t)blockIdx.x * T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned int carry = 0;
    for (int base = 0; base < T; base += 1024) {
        const int i = base + tid;
        const unsigned int c = i < T ? row[i] : 0u;
        unsigned int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const unsigned int w = s_warp[lane];
            unsigned int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int n = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += n;
            }
            s_warp[lane] = wi - w;  // exclusive warp offsets
            if (lane == 31) s_total = wi;
        }
        __syncthreads();
        if (i < T) row[i] = carry + s_warp[warp] + incl - c;
        carry += s_total;
        __syncthreads();
    }
    if (tid == 0) totals[blockIdx.x] = carry;
}

struct RsSmem {
    uint2 pairs[kRsTile];  // (key, value) exchange buffer; its first 16 KB double as the match masks
    unsigned int warp_hist[kRsThreads / 32][kRadix];
    unsigned int digit_start[kRadix];
    int gbase[kRadix];
    unsigned int warp_sums[kRadix / 32];
};

// FULL: the tile holds exactly kRsTile pairs (every tile but the last) -- no bounds predicates
template <int NBITS, bool FULL>
__device__ __forceinline__ void rs_scatter_tile(RsSmem& sm, const unsigned int* __restrict__ keys_in,
                                                const unsigned int* __restrict__ vals_in,
                                                unsigned int* __restrict__ keys_out, unsigned int* __restrict__ vals_out,
                                                long long tile_base, int tile_n, int shift, int T, int tile,
                                                const unsigned int* __restrict__ counts_excl,
                                                const unsigned int* __restrict__ totals) {
    constexpr int NB = 1 << NBITS;
    constexpr unsigned int kDigitMask = NB - 1u;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // warp-striped load: item k of lane l in warp w sits at w*32*IPT + k*32 + l
    unsigned int key[kRsItems];
    unsigned int val[kRsItems];
    const int wbase = warp * 32 * kRsItems + lane;
#pragma unroll
    for (int k = 0; k < kRsItems; k++) {
        const int p = wbase + k * 32;
        if (FULL || p < tile_n) {
            key[k] = keys_in[tile_base + p];
            val[k] = vals_in[tile_base + p];
        } else {
            key[k] = ~0u;
            val[k] = 0;
        }
    }
    // this tile's global write bases (independent of the ranking below: issue the loads early)
    const unsigned int my_total = tid < NB ? totals[tid] : 0u;
    const unsigned int my_excl = tid < NB ? counts_excl[(size_t)tid * T + tile] : 0u;

    // Warp-level stable ranking; counters private to the warp.  The lanes holding the same digit
    // ("peers") are found by OR-ing lane bits into a warp-private shared-memory word per digit
    // (measured on B200 for 11 M keys: +8 us, against +22 us for NBITS ballots and +64 us for
    // MATCH.ANY); two mask buffers alternate so that clearing overlaps the next item.  Every lane
    // reads the digit's running count, the first peer adds the group size: no shuffle, no branch.
    const unsigned int lt_mask = (1u << lane) - 1u;
    unsigned int* wh = sm.warp_hist[warp];
    unsigned int* mk = reinterpret_cast<unsigned int*>(sm.pairs) + warp * (2 * kRadix);
    for (int k = lane; k < 2 * NB; k += 32) mk[(k / NB) * kRadix + (k % NB)] = 0;
    __syncwarp();
    unsigned int rank2[kRsItems / 2];  // two 16-bit ranks per register (rank < kRsTile = 4096)
#pragma unroll
    for (int k = 0; k < kRsItems; k++) {
        const bool valid = FULL || (wbase + k * 32) < tile_n;
        const unsigned int d = (key[k] >> shift) & kDigitMask;
        unsigned int* m = mk + (k & 1) * kRadix;
        if (valid) atomicOr(&m[d], 1u << lane);
        __syncwarp();
        const unsigned int peers = m[d];
        const unsigned int old = wh[d];
        // the other buffer was last read one item ago: clear it before the next item ORs into it
        if (k > 0) mk[((k - 1) & 1) * kRadix + ((key[k - 1] >> shift) & kDigitMask)] = 0;
        const unsigned int below = __popc(peers & lt_mask);
        const unsigned int r = old + below;
        if (k & 1) rank2[k >> 1] |= r << 16; else rank2[k >> 1] = r;
        __syncwarp();
        if (valid && below == 0) wh[d] = old + __popc(peers);
    }
    __syncthreads();
    // per digit (thread = digit): exclusive offsets across warps, start of the digit inside the
    // sorted tile, start of the digit in the output (exclusive scan of the totals)
    {
        unsigned int count = 0;
#pragma unroll
        for (int w = 0; w < kRsThreads / 32; w++) {
            const unsigned int c = sm.warp_hist[w][tid];
            sm.warp_hist[w][tid] = count;
            count += c;
        }
        unsigned int incl = count, tincl = my_total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int n = __shfl_up_sync(0xffffffffu, incl, o);
            const unsigned int tn = __shfl_up_sync(0xffffffffu, tincl, o);
            if (lane >= o) { incl += n; tincl += tn; }
        }
        if (lane == 31) { sm.warp_sums[warp] = incl; sm.digit_start[warp] = tincl; }  // digit_start reused as scratch
        __syncthreads();
        unsigned int off = 0, toff = 0;
        for (int w = 0; w < warp; w++) { off += sm.warp_sums[w]; toff += sm.digit_start[w]; }
        __syncthreads();
        const unsigned int dstart = off + incl - count;
        sm.digit_start[tid] = dstart;
        sm.gbase[tid] = (int)(toff + tincl - my_total + my_excl) - (int)dstart;
    }
    __syncthreads();
    // scatter into tile-local sorted order
#pragma unroll
    for (int k = 0; k < kRsItems; k++) {
        if (FULL || (wbase + k * 32) < tile_n) {
            const unsigned int d = (key[k] >> shift) & kDigitMask;
            const unsigned int pos = sm.digit_start[d] + sm.warp_hist[warp][d] + ((rank2[k >> 1] >> (16 * (k & 1))) & 0xffffu);
            sm.pairs[pos] = make_uint2(key[k], val[k]);
        }
    }
    __syncthreads();
    // coalesced write-out: consecutive sorted positions of one digit are consecutive in global memory
#pragma unroll
    for (int k = 0; k < kRsItems; k++) {
        const int p = k * kRsThreads + tid;
        if (FULL || p < tile_n) {
            const uint2 kv = sm.pairs[p];
            const int dst = sm.gbase[(kv.x >> shift) & kDigitMask] + p;
            keys_out[dst] = kv.x;
            vals_out[dst] = kv.y;
        }
    }
}

template <int NBITS>
__global__ void __launch_bounds__(kRsThreads, 3)
rs_scatter_kernel(const unsigned int* __restrict__ keys_in, const unsigned int* __restrict__ vals_in,
                  unsigned int* __restrict__ keys_out, unsigned int* __restrict__ vals_out, int N_cap,
                  const int* __restrict__ n_dev, int shift, int T, const unsigned int* __restrict__ counts_excl,
                  const unsigned int* __restrict__ totals) {
    const int N = n_dev ? min(*n_dev, N_cap) : N_cap;
    __shared__ RsSmem sm;
    const int tile = blockIdx.x;
    const long long tile_base = (long long)tile * kRsTile;
    if (tile_base >= N) return;  // capacity-sized grid: surplus CTAs leave
    for (int k = threadIdx.x; k < (kRsThreads / 32) * kRadix; k += kRsThreads) (&sm.warp_hist[0][0])[k] = 0;
    __syncthreads();
    const int tile_n = (int)min((long long)kRsTile, (long long)N - tile_base);
    if (tile_n == kRsTile)
        rs_scatter_tile<NBITS, true>(sm, keys_in, vals_in, keys_out, vals_out, tile_base, tile_n, shift, T, tile, counts_excl, totals);
    else
        rs_scatter_tile<NBITS, false>(sm, keys_in, vals_in, keys_out, vals_out, tile_base, tile_n, shift, T, tile, counts_excl, totals);
}

// ---------------------------------------------------------------------------
// tile ranges (sort_gaussian.cu:45-71) from the sorted tile ids
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_range_kernel(int N_cap, const int* __restrict__ n_dev, const unsigned int* __restrict__ tile_sorted,
                  int num_tiles, int2* __restrict__ tile_range) {
    const int N = n_dev ? min(*n_dev, N_cap) : N_cap;
    // four consecutive entries per thread (one 16-byte load) plus the entry before them
    const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= N) return;
    unsigned int v[5];
    v[0] = i0 > 0 ? tile_sorted[i0 - 1] : 0xffffffffu;
    if (i0 + 4 <= N) {
        const uint4 q = *reinterpret_cast<const uint4*>(tile_sorted + i0);
        v[1] = q.x; v[2] = q.y; v[3] = q.z; v[4] = q.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) v[1 + k] = (i0 + k < N) ? tile_sorted[i0 + k] : 0xffffffffu;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = i0 + k;
        if (i >= N) break;
        const unsigned int cur = v[1 + k], prev = v[k];
        if (cur >= (unsigned)num_tiles) continue;
        if (i == 0) tile_range[cur].x = 0;
        else if (prev != cur) {
            tile_range[cur].x = i;
            if (prev < (unsigned)num_tiles) tile_range[prev].y = i;
        }
        if (i == N - 1) tile_range[cur].y = N;
    }
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
static inline int tile_bits(int num_tiles) { return max(1, ceil_log2(num_tiles)); }
static inline int tile_passes(int num_tiles) { return (tile_bits(num_tiles) + 7) / 8; }

// P-sized workspace: survives from pxb_bin_prepare to pxb_sort_gaussian
struct WsP {
    unsigned char* ctl; size_t ctl_bytes;          // zeroed per call
    unsigned long long* scan_status; unsigned int* scan_ticket;
    unsigned int* totals; unsigned int* counts;
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
    b.ctl_bytes = o;
    b.totals = (unsigned int*)(base + o); o += align_up((size_t)kRadix * 4, 256);
    b.counts = (unsigned int*)(base + o); o += align_up(rs_tiles * kRadix * 4, 256);
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
    unsigned int* totals; unsigned int* counts;
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
    b.ctl_bytes = 0;
    b.totals = (unsigned int*)(base + o); o += align_up((size_t)kRadix * 4, 256);
    b.counts = (unsigned int*)(base + o); o += align_up(rs_tiles * kRadix * 4, 256);
    for (int k = 0; k < 2; k++) { b.keys[k] = (unsigned int*)(base + o); o += align_up((size_t)Ncap * 4, 256); }
    b.vals_tmp = (unsigned int*)(base + o); o += align_up((size_t)Ncap * 4, 256);
    b.total = o;
    return b;
}

template <int NBITS>
static void launch_pass(int T, const unsigned int* ki, const unsigned int* vi, unsigned int* ko, unsigned int* vo, int N_cap,
                        const int* n_dev, int shift, unsigned int* counts, unsigned int* totals, cudaStream_t s) {
    static bool attr = false;
    if (!attr) {  // shared memory, not L1, is what the scatter kernel lives on
        cudaFuncSetAttribute(rs_scatter_kernel<NBITS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr = true;
    }
    rs_tile_hist_kernel<NBITS><<<T, kRsThreads, 0, s>>>(ki, N_cap, n_dev, shift, T, counts);
    rs_tile_scan_kernel<<<1 << NBITS, 1024, 0, s>>>(counts, T, totals);
    rs_scatter_kernel<NBITS><<<T, kRsThreads, 0, s>>>(ki, vi, ko, vo, N_cap, n_dev, shift, T, counts, totals);
}

// one stable LSD radix sort over the low `total_bits` key bits; result lands in (k[passes&1], v[passes&1]).
// counts: kRadix * T words of scratch, totals: kRadix words (neither needs initialising)
static int radix_sort_u32(unsigned int* k[2], unsigned int* v[2], int N_cap, const int* n_dev, int total_bits,
                          unsigned int* counts, unsigned int* totals, cudaStream_t s) {
    const RsDigits dg = make_digits(total_bits);
    const int T = (N_cap + kRsTile - 1) / kRsTile;
    for (int p = 0; p < dg.passes; p++) {
        const unsigned int *ki = k[p & 1], *vi = v[p & 1];
        unsigned int *ko = k[(p + 1) & 1], *vo = v[(p + 1) & 1];
        switch (dg.bits[p]) {
#define PXB_PASS(B) case B: launch_pass<B>(T, ki, vi, ko, vo, N_cap, n_dev, dg.shift[p], counts, totals, s); break;
            PXB_PASS(1) PXB_PASS(2) PXB_PASS(3) PXB_PASS(4) PXB_PASS(5) PXB_PASS(6) PXB_PASS(7) PXB_PASS(8)
#undef PXB_PASS
        }
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
    return pxb::bin_prepare(P, depth, radius, tiles, total_dev, nullptr, ws_p, ws_p_bytes, stream);
}

}  // extern "C"

// total_host: optional pinned (device-mapped) host word that also receives the count
int pxb::bin_prepare(int P, const float* depth, const int* radius, const int* tiles, int* total_dev, int* total_host,
                     void* ws_p, size_t ws_p_bytes, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P <= 0) return (int)cudaMemsetAsync(total_dev, 0, sizeof(int), s);
    WsP b = carve_p(ws_p, P);
    if (ws_p_bytes < b.total) return PXB_ERR_WORKSPACE;
    PXB_CUDA_OK(cudaMemsetAsync(b.ctl, 0, b.ctl_bytes, s));
    init_depth_keys_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, depth, b.keys[0], b.vals[0]);
    int rc = radix_sort_u32(b.keys, b.vals, P, nullptr, 32, b.counts, b.totals, s);  // -> keys[0]/vals[0]
    if (rc) return rc;
    scan_kernel<<<(P + kScanTile - 1) / kScanTile, kScanThreads, 0, s>>>(P, tiles, radius, b.vals[0], b.offsets, total_dev,
                                                                        total_host, b.scan_status, b.scan_ticket);
    return (int)cudaGetLastError();
}

extern "C" {

// keys + tile sort + ranges.  N: exact count when total_dev == NULL, else a capacity: the kernels
// then process min(*total_dev, N) intersections and the caller verifies *total_dev <= N afterwards.
int pxb_sort_gaussian(int P, long long N, const int* total_dev, const float* uv, int uv_stride, int tight,
                      const float* depth, const int* radius, const int* tiles, int W, int H, int* idx_sorted, int* tile_range,
                      long long* keys_sorted_out /*nullable, [N]*/, void* ws_p, size_t ws_p_bytes, void* ws_n,
                      size_t ws_n_bytes, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const int gx = (W + PXB_TILE - 1) / PXB_TILE, gy = (H + PXB_TILE - 1) / PXB_TILE;
    const int num_tiles = gx * gy;
    PXB_CUDA_OK(cudaMemsetAsync(tile_range, 0, (size_t)num_tiles * 2 * sizeof(int), s));
    if (P <= 0 || N <= 0) return 0;
    if (N > 0x3fffffffll || (tight && uv_stride < 6)) return PXB_ERR_BAD_ARG;
    WsP bp = carve_p(ws_p, P);
    WsN bn = carve_n(ws_n, N, num_tiles);
    if (ws_p_bytes < bp.total || ws_n_bytes < bn.total) return PXB_ERR_WORKSPACE;
    const int passes = tile_passes(num_tiles);
    // ping-pong so that the last pass writes the values straight into idx_sorted
    unsigned int* k[2] = {bn.keys[0], bn.keys[1]};
    unsigned int* v[2];
    v[passes & 1] = (unsigned int*)idx_sorted;
    v[(passes + 1) & 1] = bn.vals_tmp;
    emit_keys_kernel<<<(P + kEmitThreads - 1) / kEmitThreads, kEmitThreads, 0, s>>>(
        P, uv, uv_stride, radius, tiles, bp.vals[0], bp.offsets, gx, gy, tight, N, total_dev, k[0], v[0]);
    int rc = radix_sort_u32(k, v, (int)N, total_dev, tile_bits(num_tiles), bn.counts, bn.totals, s);
    if (rc) return rc;
    const unsigned int* tile_sorted = k[passes & 1];
    tile_range_kernel<<<(int)((N + 1023) / 1024), 256, 0, s>>>((int)N, total_dev, tile_sorted, num_tiles, (int2*)tile_range);
    if (keys_sorted_out && total_dev == nullptr)
        rebuild_keys_kernel<<<(int)((N + 255) / 256), 256, 0, s>>>((int)N, tile_sorted, idx_sorted, depth, keys_sorted_out);
    return (int)cudaGetLastError();
}

}  // extern "C"
