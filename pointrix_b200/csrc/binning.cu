// Tile binning (reference: msplat/msplat/sort_gaussian.py:42-52 and
// msplat/msplat/src/sort_gaussian.cu:17-71).
//
// The reference sorts all N tile intersections by the 64-bit key tile<<32|depth (torch.sort, 8 radix
// passes over N pairs + a gather).  The same order -- ascending (tile, depth bits), ties in
// emission order = ascending Gaussian id -- is produced here with far less traffic:
//   1. order-preserving compaction of the VISIBLE Gaussians (radius > 0 and at least one tile), then a
//      stable LSD radix sort of those M <= P (depth bits, id) pairs (4 passes over M, M << N).  The
//      digit histogram of pass k+1 is accumulated by the scatter of pass k (one RED per pair), that of
//      pass 0 by the compaction, so a pass is two kernels (scan over chunks, scatter) instead of three;
//   2. the last pass also sums the tiles-touched of every 1024 sorted Gaussians; the key emission
//      turns those chunk sums into its write offsets itself (no separate prefix-sum kernel, no offsets
//      array) and publishes the intersection count;
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

constexpr int kRadix = 256;

__device__ __forceinline__ unsigned int warp_sum_u32(unsigned int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------
// order-preserving compaction of the visible Gaussians (+ digit histogram of radix pass 0)
//   visible(g) = radius[g] > 0 && tiles[g] > 0      (tiles: the count the emission will use)
// Two kernels without any inter-CTA dependency: per-chunk counts (skipped when the fused forward has
// already accumulated them), then every CTA sums the counts of the chunks before it (<= P/1024 words).
// ---------------------------------------------------------------------------
constexpr int kCompThreads = 256;
constexpr int kCompChunk = 1024;  // Gaussians per CTA, 4 consecutive ids per thread

__global__ void __launch_bounds__(kCompThreads)
count_visible_kernel(int P, const int* __restrict__ radius, const int* __restrict__ tiles,
                     unsigned int* __restrict__ vis_cnt) {
    pdl_wait();
    __shared__ unsigned int s_w[kCompThreads / 32];
    const int i0 = blockIdx.x * kCompChunk + threadIdx.x * 4;
    unsigned int c = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (i0 + k < P) c += (radius[i0 + k] > 0 && tiles[i0 + k] > 0) ? 1u : 0u;
    c = warp_sum_u32(c);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = 0;
#pragma unroll
        for (int w = 0; w < kCompThreads / 32; w++) t += s_w[w];
        vis_cnt[blockIdx.x] = t;
    }
}

// keys[j] = depth bits, vals[j] = id of the j-th visible Gaussian (ascending id); *m_dev = their number;
// counts0[digit][chunk] += 1 for the lowest radix digit (chunk = j >> log_chunk, row pitch T)
__global__ void __launch_bounds__(kCompThreads)
compact_kernel(int P, const float* __restrict__ depth, const int* __restrict__ radius, const int* __restrict__ tiles,
               const unsigned int* __restrict__ vis_cnt, unsigned int* __restrict__ keys, unsigned int* __restrict__ vals,
               int* __restrict__ m_dev, unsigned int* __restrict__ counts0, int T, int log_chunk) {
    pdl_wait();
    __shared__ unsigned int s_w[kCompThreads / 32], s_v[kCompThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned int part = 0;
    for (int c = tid; c < (int)blockIdx.x; c += kCompThreads) part += vis_cnt[c];
    part = warp_sum_u32(part);
    const int i0 = blockIdx.x * kCompChunk + tid * 4;
    bool vis[4];
    unsigned int n = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        vis[k] = (i0 + k < P) && radius[i0 + k] > 0 && tiles[i0 + k] > 0;
        n += vis[k] ? 1u : 0u;
    }
    unsigned int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += x;
    }
    if (lane == 0) s_w[warp] = part;
    if (lane == 31) s_v[warp] = incl;
    __syncthreads();
    unsigned int base = 0, woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kCompThreads / 32; w++) {
        base += s_w[w];
        if (w < warp) woff += s_v[w];
        total += s_v[w];
    }
    unsigned int pos = base + woff + incl - n;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (vis[k]) {
            const unsigned int key = __float_as_uint(depth[i0 + k]);
            keys[pos] = key;
            vals[pos] = (unsigned int)(i0 + k);
            atomicAdd(&counts0[(size_t)(key & (kRadix - 1)) * T + (pos >> log_chunk)], 1u);
            pos++;
        }
    if (blockIdx.x == gridDim.x - 1 && tid == 0) *m_dev = (int)(base + total);
}

// *total = sum of the chunk sums (the intersection count N); operator path only -- the fused path lets the
// key emission publish it
__global__ void __launch_bounds__(256)
sum_chunks_kernel(const int* __restrict__ m_dev, const int* __restrict__ chunk_tiles, int* __restrict__ total) {
    pdl_wait();
    __shared__ unsigned int s_w[8];
    const int nchunks = (*m_dev + kCompChunk - 1) / kCompChunk;
    unsigned int v = 0;
    for (int c = threadIdx.x; c < nchunks; c += 256) v += (unsigned int)chunk_tiles[c];
    v = warp_sum_u32(v);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = 0;
        for (int w = 0; w < 8; w++) t += s_w[w];
        *total = (int)t;
    }
}

// ---------------------------------------------------------------------------
// key emission: sorted position i (< M) holds Gaussian g = order[i]; its cnt(g) tile ids go to
// [off, off + cnt) with off = sum of cnt over the sorted positions before i.  A CTA owns 1024 sorted
// positions: the sums of the chunks before it (accumulated by the last depth pass) give its base, a
// CTA-wide scan per round of 256 the rest.  Block 0 publishes N = sum of all chunks.
// ---------------------------------------------------------------------------
constexpr int kEmitThreads = 256;

__global__ void __launch_bounds__(kEmitThreads)
emit_keys_kernel(const int* __restrict__ m_dev, const float* __restrict__ uv, int uv_stride, const int* __restrict__ radius,
                 const int* __restrict__ tiles, const int2* __restrict__ rect, const unsigned int* __restrict__ order,
                 const int* __restrict__ chunk_tiles, int gx, int gy, int tight, long long N_cap, int* __restrict__ n_dev,
                 int* __restrict__ n_host, unsigned int* __restrict__ keys, unsigned int* __restrict__ vals) {
    pdl_wait();
    __shared__ unsigned int s_a[kEmitThreads / 32], s_b[kEmitThreads / 32];
    const int M = *m_dev;
    const int nchunks = (M + kCompChunk - 1) / kCompChunk;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if ((int)blockIdx.x >= nchunks && blockIdx.x != 0) return;
    unsigned int before = 0, all = 0;
    for (int c = tid; c < nchunks; c += kEmitThreads) {
        const unsigned int v = (unsigned int)chunk_tiles[c];
        all += v;
        if (c < (int)blockIdx.x) before += v;
    }
    before = warp_sum_u32(before);
    all = warp_sum_u32(all);
    if (lane == 0) { s_a[warp] = before; s_b[warp] = all; }
    __syncthreads();
    unsigned int run = 0, total_all = 0;
#pragma unroll
    for (int w = 0; w < kEmitThreads / 32; w++) { run += s_a[w]; total_all += s_b[w]; }
    if (blockIdx.x == 0 && tid == 0) {
        *n_dev = (int)total_all;
        if (n_host) {  // pinned host word the caller polls (no memcpy in the stream)
            *reinterpret_cast<volatile int*>(n_host) = (int)total_all;
            __threadfence_system();
        }
    }
    if ((int)blockIdx.x >= nchunks) return;
    __syncthreads();
    for (int round = 0; round < kCompChunk / kEmitThreads; round++) {
        const int i = blockIdx.x * kCompChunk + round * kEmitThreads + tid;
        int x0 = 0, y0 = 0, w = 1, cnt = 0, g = 0;
        if (i < M) {
            g = (int)order[i];
            if (rect != nullptr) {  // the fused forward's own rectangle: {x0 | y0 << 16, w | h << 16}
                const int2 r = rect[g];
                x0 = r.x & 0xffff; y0 = r.x >> 16;
                w = max(r.y & 0xffff, 1);
                cnt = (r.y & 0xffff) * (r.y >> 16);
            } else {
                int x1, y1;
                const float* rg = uv + (size_t)g * uv_stride;
                if (tight)  // uv is the packed blend record {u,v,A,B,C,op,...}
                    tight_tile_rect(rg[0], rg[1], radius[g], rg[2], rg[3], rg[4], rg[5], gx, gy, x0, y0, x1, y1);
                else
                    tile_rect(rg[0], rg[1], radius[g], gx, gy, x0, y0, x1, y1);
                w = max(x1 - x0, 1);
                cnt = tiles[g];  // == w * (y1 - y0) (ewa_project.cu:81); the chunk sums used the same count
            }
        }
        // warp scan of the counts, then the warps' offsets inside this round
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 31) s_a[warp] = (unsigned int)incl;
        __syncthreads();
        unsigned int woff = 0, round_total = 0;
#pragma unroll
        for (int k = 0; k < kEmitThreads / 32; k++) {
            if (k < warp) woff += s_a[k];
            round_total += s_a[k];
        }
        const int excl = incl - cnt;
        const long long wbase = (long long)run + woff;  // first slot of this warp's Gaussians
        // warp-cooperative expansion: slot s of the warp's cnt-sum belongs to the lane whose inclusive
        // prefix first exceeds s
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
            const int gs = __shfl_sync(0xffffffffu, g, src);
            if (s < total) {
                const int local = s - e_src;
                const int row = local / ws, col = local - row * ws;
                const long long pos = wbase + s;
                if (pos < N_cap) {
                    keys[pos] = (unsigned int)((y0s + row) * gx + (x0s + col));
                    vals[pos] = (unsigned int)gs;
                }
            }
        }
        run += round_total;
        __syncthreads();  // s_a is rewritten by the next round
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
constexpr int kMaxPasses = 4;
// pairs per thread: 16 (4096 per CTA) for the N-level tile sort; 8 (2048 per CTA) for the P-level depth
// sort, whose <= P/2048 CTAs are all resident at once -- smaller chunks halve the serial ranking chain
constexpr int kItemsN = 16;
constexpr int kItemsP = 8;

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

template <int NBITS, int ITEMS>
__global__ void __launch_bounds__(kRsThreads)
rs_tile_hist_kernel(const unsigned int* __restrict__ keys, int N_cap, const int* __restrict__ n_dev, int shift, int T,
                    unsigned int* __restrict__ counts /*[2^NBITS][T]*/) {
    pdl_wait();
    constexpr int NB = 1 << NBITS;
    constexpr int NW = kRsThreads / 32;
    constexpr int TILE = kRsThreads * ITEMS;
    __shared__ unsigned int h[NW][NB];  // warp-private: 8x less contention on skewed digits
    const int N = n_dev ? min(*n_dev, N_cap) : N_cap;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int k = tid; k < NW * NB; k += kRsThreads) (&h[0][0])[k] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * TILE;
    unsigned int key[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const long long p = base + k * kRsThreads + tid;
        key[k] = p < N ? keys[p] : 0u;
    }
#pragma unroll
    for (int k = 0; k < ITEMS; k++)
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
    pdl_wait();
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

// what a scatter pass accumulates for its successor while it writes out (one RED per pair)
enum RsNext {
    kNextNone = 0,
    kNextHist = 1,   // digit histogram of the NEXT pass: next_counts[digit][dst / TILE] += 1
    kNextTiles = 2,  // last depth pass: chunk_tiles[dst / 1024] += tiles[value]
};
struct RsAux {
    unsigned int* next_counts;  // kNextHist
    int next_shift;
    const int* tiles;           // kNextTiles
    int* chunk_tiles;
};

template <int ITEMS>
struct RsSmem {
    uint2 pairs[kRsThreads * ITEMS];  // (key, value) exchange buffer; its first 16 KB double as the match masks
    unsigned int warp_hist[kRsThreads / 32][kRadix];
    unsigned int digit_start[kRadix];
    int gbase[kRadix];
    unsigned int warp_sums[kRadix / 32];
};

// FULL: the tile holds exactly TILE pairs (every tile but the last) -- no bounds predicates
template <int NBITS, int ITEMS, int NEXT, bool FULL>
__device__ __forceinline__ void rs_scatter_tile(RsSmem<ITEMS>& sm, const unsigned int* __restrict__ keys_in,
                                                const unsigned int* __restrict__ vals_in,
                                                unsigned int* __restrict__ keys_out, unsigned int* __restrict__ vals_out,
                                                long long tile_base, int tile_n, int shift, int T, int tile,
                                                const unsigned int* __restrict__ counts_excl,
                                                const unsigned int* __restrict__ totals, const RsAux& aux) {
    constexpr int NB = 1 << NBITS;
    constexpr unsigned int kDigitMask = NB - 1u;
    constexpr int LOG_TILE = (ITEMS == 16) ? 12 : 11;
    static_assert(kRsThreads * ITEMS == (1 << LOG_TILE), "tile size");
    static_assert(sizeof(uint2) * kRsThreads * ITEMS >= (kRsThreads / 32) * 2 * kRadix * 4, "mask buffers alias pairs[]");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // warp-striped load: item k of lane l in warp w sits at w*32*IPT + k*32 + l
    unsigned int key[ITEMS];
    unsigned int val[ITEMS];
    const int wbase = warp * 32 * ITEMS + lane;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
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
    unsigned int rank2[ITEMS / 2];  // two 16-bit ranks per register (rank < TILE <= 4096)
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
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
    for (int k = 0; k < ITEMS; k++) {
        if (FULL || (wbase + k * 32) < tile_n) {
            const unsigned int d = (key[k] >> shift) & kDigitMask;
            const unsigned int pos = sm.digit_start[d] + sm.warp_hist[warp][d] + ((rank2[k >> 1] >> (16 * (k & 1))) & 0xffffu);
            sm.pairs[pos] = make_uint2(key[k], val[k]);
        }
    }
    __syncthreads();
    // coalesced write-out: consecutive sorted positions of one digit are consecutive in global memory
    // The successor's counters are bumped with warp-aggregated REDs: consecutive sorted positions share their
    // output chunk, and the high depth digits take only a handful of values, so a warp's 32 updates mostly
    // hit one or two words -- issued one by one they serialise in the L2 atomic unit (measured: 148 us for
    // the 675 K updates of one pass, against 14 us for the same pass when the digits are spread).
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const int p = k * kRsThreads + tid;
        const bool live = FULL || p < tile_n;
        uint2 kv = make_uint2(0u, 0u);
        int dst = 0;
        if (live) {
            kv = sm.pairs[p];
            dst = sm.gbase[(kv.x >> shift) & kDigitMask] + p;
            keys_out[dst] = kv.x;
            vals_out[dst] = kv.y;
        }
        if (NEXT == kNextHist) {
            const unsigned int slot = live ? ((kv.x >> aux.next_shift) & (kRadix - 1)) * (unsigned int)T + (unsigned int)(dst >> LOG_TILE)
                                           : 0xffffffffu;
            const unsigned int peers = __match_any_sync(0xffffffffu, slot);
            if (live && (peers & lt_mask) == 0u) atomicAdd(&aux.next_counts[slot], (unsigned int)__popc(peers));
        }
        if (NEXT == kNextTiles) {
            const int chunk = live ? (dst >> 10) : -1;
            const int cnt = live ? aux.tiles[kv.y] : 0;
            const unsigned int peers = __match_any_sync(0xffffffffu, chunk);
            const int sum = __reduce_add_sync(peers, cnt);
            if (live && (peers & lt_mask) == 0u) atomicAdd(&aux.chunk_tiles[chunk], sum);
        }
    }
}

template <int NBITS, int ITEMS, int NEXT>
__global__ void __launch_bounds__(kRsThreads, 3)
rs_scatter_kernel(const unsigned int* __restrict__ keys_in, const unsigned int* __restrict__ vals_in,
                  unsigned int* __restrict__ keys_out, unsigned int* __restrict__ vals_out, int N_cap,
                  const int* __restrict__ n_dev, int shift, int T, const unsigned int* __restrict__ counts_excl,
                  const unsigned int* __restrict__ totals, RsAux aux) {
    pdl_wait();
    constexpr int TILE = kRsThreads * ITEMS;
    const int N = n_dev ? min(*n_dev, N_cap) : N_cap;
    __shared__ RsSmem<ITEMS> sm;
    const int tile = blockIdx.x;
    const long long tile_base = (long long)tile * TILE;
    if (tile_base >= N) return;  // capacity-sized grid: surplus CTAs leave
    for (int k = threadIdx.x; k < (kRsThreads / 32) * kRadix; k += kRsThreads) (&sm.warp_hist[0][0])[k] = 0;
    __syncthreads();
    const int tile_n = (int)min((long long)TILE, (long long)N - tile_base);
    if (tile_n == TILE)
        rs_scatter_tile<NBITS, ITEMS, NEXT, true>(sm, keys_in, vals_in, keys_out, vals_out, tile_base, tile_n, shift, T, tile, counts_excl, totals, aux);
    else
        rs_scatter_tile<NBITS, ITEMS, NEXT, false>(sm, keys_in, vals_in, keys_out, vals_out, tile_base, tile_n, shift, T, tile, counts_excl, totals, aux);
}

// ---------------------------------------------------------------------------
// tile ranges (sort_gaussian.cu:45-71) from the sorted tile ids
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_range_kernel(int N_cap, const int* __restrict__ n_dev, const unsigned int* __restrict__ tile_sorted,
                  int num_tiles, int2* __restrict__ tile_range) {
    pdl_wait();
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
    unsigned char* ctl; size_t ctl_bytes;            // zeroed per call (one memset)
    unsigned int* vis_cnt;                           // visible Gaussians per 1024-id chunk
    int* chunk_tiles;                                // tiles touched per 1024 depth-sorted Gaussians
    unsigned int* counts[kMaxPasses];                // [256][T] per depth pass (accumulated by REDs)
    int* m_dev;                                      // number of visible Gaussians
    unsigned int* totals;                            // [256], rewritten by every scan
    unsigned int* keys[2]; unsigned int* vals[2];    // depth-sort ping-pong; result in keys[0]/vals[0] (4 passes)
    int T;                                           // chunks of kItemsP * 256 pairs
    size_t total;
};
static WsP carve_p(void* ws, int P) {
    WsP b;
    const size_t chunks = (size_t)(P + kCompChunk - 1) / kCompChunk + 1;
    b.T = (P + kRsThreads * kItemsP - 1) / (kRsThreads * kItemsP);
    size_t o = 0;
    unsigned char* base = (unsigned char*)ws;
    b.ctl = base;
    b.m_dev = (int*)(base + o); o += 256;
    b.vis_cnt = (unsigned int*)(base + o); o += align_up(chunks * 4, 256);
    b.chunk_tiles = (int*)(base + o); o += align_up(chunks * 4, 256);
    for (int p = 0; p < kMaxPasses; p++) { b.counts[p] = (unsigned int*)(base + o); o += align_up((size_t)kRadix * b.T * 4, 256); }
    b.ctl_bytes = o;
    b.totals = (unsigned int*)(base + o); o += align_up((size_t)kRadix * 4, 256);
    for (int k = 0; k < 2; k++) {
        b.keys[k] = (unsigned int*)(base + o); o += align_up((size_t)P * 4, 256);
        b.vals[k] = (unsigned int*)(base + o); o += align_up((size_t)P * 4, 256);
    }
    b.total = o;
    return b;
}
// N-sized workspace of pxb_sort_gaussian
struct WsN {
    unsigned int* totals; unsigned int* counts;
    unsigned int* keys[2]; unsigned int* vals_tmp;
    size_t total;
};
static WsN carve_n(void* ws, long long Ncap, int num_tiles) {
    WsN b;
    const size_t rs_tiles = (size_t)((Ncap + kRsThreads * kItemsN - 1) / (kRsThreads * kItemsN)) + 1;
    size_t o = 0;
    unsigned char* base = (unsigned char*)ws;
    b.totals = (unsigned int*)(base + o); o += align_up((size_t)kRadix * 4, 256);
    b.counts = (unsigned int*)(base + o); o += align_up(rs_tiles * kRadix * 4, 256);
    for (int k = 0; k < 2; k++) { b.keys[k] = (unsigned int*)(base + o); o += align_up((size_t)Ncap * 4, 256); }
    b.vals_tmp = (unsigned int*)(base + o); o += align_up((size_t)Ncap * 4, 256);
    b.total = o;
    return b;
}

template <int NBITS, int ITEMS, int NEXT>
static cudaError_t launch_scatter(int T, const unsigned int* ki, const unsigned int* vi, unsigned int* ko, unsigned int* vo,
                                  int N_cap, const int* n_dev, int shift, unsigned int* counts, unsigned int* totals,
                                  const RsAux& aux, cudaStream_t s) {
    static bool attr = false;
    if (!attr) {  // shared memory, not L1, is what the scatter kernel lives on
        cudaFuncSetAttribute(rs_scatter_kernel<NBITS, ITEMS, NEXT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr = true;
    }
    return launch_k(rs_scatter_kernel<NBITS, ITEMS, NEXT>, dim3(T), dim3(kRsThreads), 0, s, ki, vi, ko, vo, N_cap, n_dev, shift, T,
                    (const unsigned int*)counts, (const unsigned int*)totals, aux);
}

// N-level: stable LSD radix sort over the low `total_bits` key bits, three chain-free kernels per pass
// (histogram, scan over tiles, scatter); result lands in (k[passes&1], v[passes&1]).
// counts: kRadix * T words of scratch, totals: kRadix words (neither needs initialising)
static int radix_sort_n(unsigned int* k[2], unsigned int* v[2], int N_cap, const int* n_dev, int total_bits,
                        unsigned int* counts, unsigned int* totals, cudaStream_t s) {
    const RsDigits dg = make_digits(total_bits);
    const int T = (N_cap + kRsThreads * kItemsN - 1) / (kRsThreads * kItemsN);
    const RsAux none = {nullptr, 0, nullptr, nullptr};
    for (int p = 0; p < dg.passes; p++) {
        const unsigned int *ki = k[p & 1], *vi = v[p & 1];
        unsigned int *ko = k[(p + 1) & 1], *vo = v[(p + 1) & 1];
        switch (dg.bits[p]) {
#define PXB_PASS(B)                                                                                                      \
    case B:                                                                                                              \
        PXB_CUDA_OK(launch_k(rs_tile_hist_kernel<B, kItemsN>, dim3(T), dim3(kRsThreads), 0, s, ki, N_cap, n_dev, dg.shift[p], T, counts)); \
        PXB_CUDA_OK(launch_k(rs_tile_scan_kernel, dim3(1 << B), dim3(1024), 0, s, counts, T, totals));                  \
        PXB_CUDA_OK((launch_scatter<B, kItemsN, kNextNone>(T, ki, vi, ko, vo, N_cap, n_dev, dg.shift[p], counts, totals, none, s))); \
        break;
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

// Depth-order the visible Gaussians and count the intersections.
// *total_dev = number of intersections N.  ws_p must stay untouched until pxb_sort_gaussian.
int pxb_bin_prepare(int P, const float* depth, const int* radius, const int* tiles, int* total_dev, void* ws_p,
                    size_t ws_p_bytes, void* stream) {
    return pxb::bin_prepare(P, depth, radius, tiles, total_dev, /*counted=*/0, ws_p, ws_p_bytes, stream);
}

}  // extern "C"

// the control block of the P-level workspace (counters accumulated by atomics): zero it before the fused
// forward (which counts the visible Gaussians itself) / at the start of bin_prepare
int pxb::bin_clear(int P, void* ws_p, size_t ws_p_bytes, void* stream) {
    WsP b = carve_p(ws_p, P > 0 ? P : 1);
    if (ws_p_bytes < b.total) return PXB_ERR_WORKSPACE;
    return (int)cudaMemsetAsync(b.ctl, 0, b.ctl_bytes, (cudaStream_t)stream);
}
unsigned int* pxb::bin_vis_counters(int P, void* ws_p) { return carve_p(ws_p, P > 0 ? P : 1).vis_cnt; }

// counted != 0: the control block was cleared by bin_clear and the per-chunk visible counts are already
// accumulated (fused forward); *total_dev is then left to the key emission (pxb::sort_gaussian publishes it)
int pxb::bin_prepare(int P, const float* depth, const int* radius, const int* tiles, int* total_dev, int counted,
                     void* ws_p, size_t ws_p_bytes, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P <= 0) return (int)cudaMemsetAsync(total_dev, 0, sizeof(int), s);
    WsP b = carve_p(ws_p, P);
    if (ws_p_bytes < b.total) return PXB_ERR_WORKSPACE;
    const int nchunk = (P + kCompChunk - 1) / kCompChunk;
    if (!counted) {
        PXB_CUDA_OK(cudaMemsetAsync(b.ctl, 0, b.ctl_bytes, s));
        PXB_CUDA_OK(launch_k(count_visible_kernel, dim3(nchunk), dim3(kCompThreads), 0, s, P, radius, tiles, b.vis_cnt));
    }
    constexpr int LOG_P = 11;  // kItemsP * 256 = 2048 pairs per chunk
    PXB_CUDA_OK(launch_k(compact_kernel, dim3(nchunk), dim3(kCompThreads), 0, s, P, depth, radius, tiles,
                         (const unsigned int*)b.vis_cnt, b.keys[0], b.vals[0], b.m_dev, b.counts[0], b.T, LOG_P));
    // 4 passes x 8 bits over the M visible pairs; pass p's scatter accumulates pass p+1's histogram, the last
    // one the per-chunk tile sums the key emission starts from
    for (int p = 0; p < 4; p++) {
        const unsigned int *ki = b.keys[p & 1], *vi = b.vals[p & 1];
        unsigned int *ko = b.keys[(p + 1) & 1], *vo = b.vals[(p + 1) & 1];
        PXB_CUDA_OK(launch_k(rs_tile_scan_kernel, dim3(kRadix), dim3(1024), 0, s, b.counts[p], b.T, b.totals));
        if (p < 3) {
            const RsAux aux = {b.counts[p + 1], 8 * (p + 1), nullptr, nullptr};
            PXB_CUDA_OK((launch_scatter<8, kItemsP, kNextHist>(b.T, ki, vi, ko, vo, P, b.m_dev, 8 * p, b.counts[p], b.totals, aux, s)));
        } else {
            const RsAux aux = {nullptr, 0, tiles, b.chunk_tiles};
            PXB_CUDA_OK((launch_scatter<8, kItemsP, kNextTiles>(b.T, ki, vi, ko, vo, P, b.m_dev, 8 * p, b.counts[p], b.totals, aux, s)));
        }
    }
    if (!counted)
        PXB_CUDA_OK(launch_k(sum_chunks_kernel, dim3(1), dim3(256), 0, s, (const int*)b.m_dev, (const int*)b.chunk_tiles, total_dev));
    return (int)cudaGetLastError();
}

// keys + tile sort + ranges.  N: exact count when total_dev == NULL, else a capacity: the kernels
// then process min(*total_dev, N) intersections and the caller verifies *total_dev <= N afterwards.
// publish != 0: the key emission writes the count to *total_dev (and to the pinned word total_host).
int pxb::sort_gaussian(int P, long long N, int* total_dev, int publish, int* total_host, const float* uv, int uv_stride,
                       int tight, const int* rect, const float* depth, const int* radius, const int* tiles, int W, int H,
                       int* idx_sorted, int* tile_range, long long* keys_sorted_out, void* ws_p, size_t ws_p_bytes,
                       void* ws_n, size_t ws_n_bytes, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const int gx = (W + PXB_TILE - 1) / PXB_TILE, gy = (H + PXB_TILE - 1) / PXB_TILE;
    const int num_tiles = gx * gy;
    PXB_CUDA_OK(cudaMemsetAsync(tile_range, 0, (size_t)num_tiles * 2 * sizeof(int), s));
    if (P <= 0 || N <= 0) return 0;
    if (N > 0x3fffffffll || (tight && rect == nullptr && uv_stride < 6) || gx > 0xffff || gy > 0x7fff) return PXB_ERR_BAD_ARG;
    WsP bp = carve_p(ws_p, P);
    WsN bn = carve_n(ws_n, N, num_tiles);
    if (ws_p_bytes < bp.total || ws_n_bytes < bn.total) return PXB_ERR_WORKSPACE;
    const int passes = tile_passes(num_tiles);
    // ping-pong so that the last pass writes the values straight into idx_sorted
    unsigned int* k[2] = {bn.keys[0], bn.keys[1]};
    unsigned int* v[2];
    v[passes & 1] = (unsigned int*)idx_sorted;
    v[(passes + 1) & 1] = bn.vals_tmp;
    int* n_sink = publish ? total_dev : (int*)(bp.m_dev + 1);  // scratch word of the control block otherwise
    PXB_CUDA_OK(launch_k(emit_keys_kernel, dim3((P + kCompChunk - 1) / kCompChunk), dim3(kEmitThreads), 0, s,
                         (const int*)bp.m_dev, uv, uv_stride, radius, tiles, (const int2*)rect, (const unsigned int*)bp.vals[0],
                         (const int*)bp.chunk_tiles, gx, gy, tight, N, n_sink, publish ? total_host : (int*)nullptr, k[0], v[0]));
    int rc = radix_sort_n(k, v, (int)N, total_dev, tile_bits(num_tiles), bn.counts, bn.totals, s);
    if (rc) return rc;
    const unsigned int* tile_sorted = k[passes & 1];
    PXB_CUDA_OK(launch_k(tile_range_kernel, dim3((int)((N + 1023) / 1024)), dim3(256), 0, s, (int)N, (const int*)total_dev,
                         tile_sorted, num_tiles, (int2*)tile_range));
    if (keys_sorted_out && total_dev == nullptr)
        rebuild_keys_kernel<<<(int)((N + 255) / 256), 256, 0, s>>>((int)N, tile_sorted, idx_sorted, depth, keys_sorted_out);
    return (int)cudaGetLastError();
}

extern "C" int pxb_sort_gaussian(int P, long long N, const int* total_dev, const float* uv, int uv_stride, int tight,
                                 const float* depth, const int* radius, const int* tiles, int W, int H, int* idx_sorted,
                                 int* tile_range, long long* keys_sorted_out, void* ws_p, size_t ws_p_bytes, void* ws_n,
                                 size_t ws_n_bytes, void* stream) {
    return pxb::sort_gaussian(P, N, const_cast<int*>(total_dev), /*publish=*/0, nullptr, uv, uv_stride, tight, nullptr, depth,
                              radius, tiles, W, H, idx_sorted, tile_range, keys_sorted_out, ws_p, ws_p_bytes, ws_n, ws_n_bytes,
                              stream);
}
