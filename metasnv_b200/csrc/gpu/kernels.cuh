// kernels.cuh -- sm_100a kernels of the pileup + call path (device side of include/msnv.h).
//
// Data flow for one shard (all samples of one genome bin):
//   index_kernel     (tile, sample) pairs that have reads  -> ordered work items (ballot compaction)
//   pileup_kernel    per item: stage the reads' position-aligned segments (TMA bulk copies) ->
//                    mate-overlap quality correction (SURVEY.md Annex A.2) in shared memory ->
//                    four positions per thread scattered into byte-lane counters
//                    -> packed A/C/G/T/N counts, 10 B per sample-position      [dominant kernel]
//   call_kernel      per tile: reduce over samples, snpCall thresholds (call_vC.cpp:545-601)
//   compact_kernel   ordered stream compaction of called positions (warp ballot + block scan)
//   gather_kernel    per hit: per-sample coverage / allele counts for the host formatter
//
// The reference computes the same quantities one text character at a time
// (call_vC.cpp:503-535 over the columns rendered by `samtools mpileup`).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../../include/msnv.h"
#include "overlap_rule.h"

namespace msnv_gpu {

constexpr int TILE = MSNV_TILE;                 // positions per tile
// CTA shapes of the pileup kernel (threads, reads staged per chunk <= threads - 1: one thread per read in the
// per-read steps; at most 255 so that 8-bit per-chunk counters cannot overflow). A CTA's life is a sequence of
// short dependent steps, so many small CTAs per SM keep the SM busy better than few large ones; deep data gets
// the larger shape (fewer chunks per tile).
constexpr int CHUNK_Q4_MAX = 4096;              // upper bound of the quads staged per chunk (chosen per launch)
constexpr int CHUNK_Q4_MIN = MSNV_MAX_READ_BASES / 4 + 2 * MSNV_MAX_READ_SEGMENTS;   // a single read always fits
constexpr int CHUNK_SEGS = 256;                 // aligned segments per chunk (<= 256: one byte tags a quad)

static_assert(MSNV_MAX_READ_SEGMENTS <= CHUNK_SEGS, "one read's segments must fit a chunk");

struct SampleDev {
    const int32_t*  pos;
    const uint32_t* seg_off;
    const uint32_t* q4_off;
    const int32_t*  mate;
    const int32_t*  seg_pos;
    const uint16_t* seg_len;
    const uint8_t*  seq2;
    const uint8_t*  qual;
    uint32_t        n_reads, max_span;
};

struct Item { uint32_t sample, tile, r_lo, r_hi; };   // reads [r_lo, r_hi) of `sample` may overlap `tile`

// Per-position population result of call_kernel.
struct CallParamsDev { int32_t min_cov; int32_t thr; double frac; };

// ------------------------------------------------------------------------------------------------
// small PTX helpers (mbarrier + 1-D TMA bulk copy), see blackwell_cuda_programming.md G3/G15
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// global -> shared bulk copy (TMA, SASS UBLKCP); dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ uint32_t lower_bound_i32(const int32_t* __restrict__ a, uint32_t n, int64_t key)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// lower_bound_i32 when the answer is expected near `guess`: gallop away from the guess in doubling steps,
// then bisect the bracket. A good guess costs ~8 probes that share one or two cache lines instead of
// log2(n) scattered ones; a bad one costs 2*log2(error).
__device__ __forceinline__ uint32_t lower_bound_near_i32(const int32_t* __restrict__ a, uint32_t n, int64_t key, uint32_t guess)
{
    uint32_t lo, hi, cur = guess < n ? guess : n, step = 16;
    if (cur < n && (int64_t)__ldg(a + cur) < key) {            // the answer lies above the guess
        lo = cur + 1;
        for (;;) {
            const uint32_t c = cur + step;
            if (c >= n) { hi = n; break; }
            if ((int64_t)__ldg(a + c) >= key) { hi = c; break; }
            lo = c + 1; cur = c; step <<= 1;
        }
    } else {                                                   // at or below it
        hi = cur;
        for (;;) {
            if (cur < step) { lo = 0; break; }
            const uint32_t c = cur - step;
            if ((int64_t)__ldg(a + c) < key) { lo = c + 1; break; }
            hi = c; cur = c; step <<= 1;
        }
    }
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Block-wide exclusive rank of `flag` among the block's threads (thread order) and the block total.
// blockDim.x must be a multiple of 32 and at most 1024.
__device__ __forceinline__ uint32_t block_rank(bool flag, uint32_t* s_warp /*[33]*/, uint32_t& total)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    uint32_t bal = __ballot_sync(0xffffffffu, flag);
    uint32_t r = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
        uint32_t v = lane < nwarp ? s_warp[lane] : 0;
        uint32_t incl = v;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        s_warp[lane] = incl - v;
        if (lane == 31) s_warp[32] = incl;
    }
    __syncthreads();
    total = s_warp[32];
    uint32_t base = s_warp[warp];
    __syncthreads();
    return base + r;
}

// ------------------------------------------------------------------------------------------------
// index: which (tile, sample) pairs have reads. Pairs are enumerated tile-major so that the
// compacted item list is grouped by tile (call_kernel reads one contiguous run of slots per tile).
// Two passes over the same predicate: COUNT writes per-block totals, EMIT writes the items at the
// scanned offsets (ordered stream compaction by warp ballot).
// ------------------------------------------------------------------------------------------------
// Occupancy bitmap for sparse shards (most (tile, sample) pairs have no reads at all, e.g. a sample that
// carries 10 % of the genomes of its bin): one bit per (sample, tile), set for the tile a read starts in.
// index_kernel then skips the two binary searches of every pair whose tile and the tiles a read could
// reach it from are all clear.
__global__ void __launch_bounds__(256) mark_kernel(const SampleDev* __restrict__ samples, uint32_t words_per_sample,
                                                   uint32_t* __restrict__ bitmap)
{
    const SampleDev sd = samples[blockIdx.y];
    uint32_t* bm = bitmap + (size_t)blockIdx.y * words_per_sample;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < sd.n_reads; i += gridDim.x * blockDim.x) {
        const uint32_t t = (uint32_t)__ldg(sd.pos + i) / TILE;
        // consecutive reads mostly share a tile: one atomic per (warp, tile)
        const uint32_t peers = __match_any_sync(__activemask(), t);
        if ((threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1)) atomicOr(bm + (t >> 5), 1u << (t & 31));
    }
}

template <bool EMIT>
__global__ void __launch_bounds__(256) index_kernel(const SampleDev* __restrict__ samples, uint32_t n_samples,
                                                    uint32_t n_tiles, uint32_t* __restrict__ block_sums,
                                                    Item* __restrict__ items, uint32_t* __restrict__ tile_begin,
                                                    uint2* __restrict__ range_cache /* [tiles*samples] or null: COUNT stores, EMIT reloads */,
                                                    const uint32_t* __restrict__ bitmap /* mark_kernel's, or null */, uint32_t words_per_sample)
{
    __shared__ uint32_t s_warp[33];
    const uint64_t pair = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    const uint64_t n_pairs = (uint64_t)n_tiles * n_samples;
    bool active = false;
    uint32_t s = 0, t = 0, r_lo = 0, r_hi = 0;
    if (pair < n_pairs) {
        t = (uint32_t)(pair / n_samples);
        s = (uint32_t)(pair - (uint64_t)t * n_samples);
        if (EMIT && range_cache) {
            const uint2 c = range_cache[pair];
            r_lo = c.x; r_hi = c.y;
        } else {
            const uint32_t n = samples[s].n_reads;
            bool maybe = n != 0;
            if (maybe && bitmap) {                       // any read starting in a tile that can reach tile t?
                const uint32_t back = (samples[s].max_span + TILE - 2) / TILE;      // tiles a read can reach back from
                const uint32_t* bm = bitmap + (size_t)s * words_per_sample;
                maybe = false;
                for (uint32_t u = t >= back ? t - back : 0; u <= t && !maybe; ++u) maybe = (__ldg(bm + (u >> 5)) >> (u & 31)) & 1u;
            }
            if (maybe) {
                const int32_t* pos = samples[s].pos;
                const int64_t t0 = (int64_t)t * TILE, k_lo = t0 - (int64_t)samples[s].max_span + 1;
                if (bitmap) {
                    // sparse shard: the sample covers a few contigs of many, its read density is anything but even
                    r_lo = lower_bound_i32(pos, n, k_lo);
                    r_hi = lower_bound_near_i32(pos, n, t0 + TILE, r_lo + 32);          // ... but the end is close to the start
                } else {
                    // reads are spread fairly evenly over what the sample covers: interpolate, then search near the guess
                    const int64_t first = __ldg(pos), extent = (int64_t)__ldg(pos + n - 1) - first + 1;
                    const int64_t g = k_lo <= first ? 0 : (k_lo - first) * (int64_t)n / extent;
                    r_lo = lower_bound_near_i32(pos, n, k_lo, (uint32_t)(g < (int64_t)n ? g : (int64_t)n));
                    const int64_t w = ((int64_t)TILE + samples[s].max_span) * (int64_t)n / extent;
                    r_hi = lower_bound_near_i32(pos, n, t0 + TILE, r_lo + (uint32_t)(w < (int64_t)n ? w : (int64_t)n));
                }
            }
            if (!EMIT && range_cache) range_cache[pair] = make_uint2(r_lo, r_hi);
        }
        active = r_hi > r_lo;
    }
    uint32_t total;
    uint32_t rank = block_rank(active, s_warp, total);
    if (!EMIT) {
        if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
    } else {
        const uint32_t slot = block_sums[blockIdx.x] + rank;     // block_sums holds exclusive offsets now
        if (active) items[slot] = Item{s, t, r_lo, r_hi};
        if (pair < n_pairs && s == 0) tile_begin[t] = slot;
    }
}

// Work items of dense count tiles handed in by the host (classic text mode): every sample on every tile.
__global__ void dense_items_kernel(uint32_t n_samples, uint32_t n_tiles, Item* __restrict__ items, uint32_t* __restrict__ tile_begin)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n = (uint64_t)n_samples * n_tiles;
    if (i < n) {
        const uint32_t t = (uint32_t)(i / n_samples), s = (uint32_t)(i - (uint64_t)t * n_samples);
        items[i] = Item{s, t, 0u, 0u};
        if (s == 0) tile_begin[t] = (uint32_t)i;
    }
    if (i == n) tile_begin[n_tiles] = (uint32_t)n;
}

// Exclusive scan of `v[0..n)` in place by one CTA; total to *total.
__global__ void __launch_bounds__(1024) scan_kernel(uint32_t* __restrict__ v, uint32_t n, uint32_t* __restrict__ total)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t x = i < n ? v[i] : 0;
        uint32_t incl = x;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane], wi = w;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += o;
            }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        if (i < n) v[i] = carry + s_warp[warp] + incl - x;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[warp] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}

// ------------------------------------------------------------------------------------------------
// helpers of the pileup kernel
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// pileup: one CTA (128 or 256 threads, see the variant table in msnv_gpu.cu) per work item
// (sample, tile of TILE positions).
// Reads arrive as position-aligned segments (include/msnv.h): a staged quad holds four consecutive
// positions starting at a multiple of four, so a quad is either on the tile or off it, and the four
// bases of a quad are handled with byte-lane (SWAR) arithmetic, never one at a time.
// Per chunk of reads (<= CHUNK_READS reads, chunk_q4 quads, CHUNK_SEGS segments):
//   1. offsets and mates of the chunk's reads -> shared memory; __syncthreads_count sizes the chunk
//   2. two TMA bulk copies (2-bit bases, qualities) are issued; while they fly,
//   3. one thread per read loads its segment records, derives for every segment where its quads
//      lie in the staging buffer and on the tile, and tags every quad with (the low byte of) its
//      tile-relative index, four tags per store; after the copies have landed the same threads zero
//      the qualities of the few quads that lie off the tile
//   4. mate-overlap quality correction in shared memory (overlap_rule.h), eight lanes per pair, a
//      quad of both mates per lane and step (mates staged in another chunk are read, pristine,
//      from global memory)
//   5. flat scatter: thread g takes the g-th staged quad: quality test and base decoding for the
//      four positions at once, then ONE shared-memory atomic per base letter. Plane X of the
//      counters holds, per quad of the tile, a word with one byte lane per position; the plane
//      offsets are immediates. Padding bytes and quads off the tile carry quality 0 and fail the
//      threshold like any poor base: the loop has no bounds test and no table look-up.
//   6. thread t folds the byte lanes of its quads into 16-bit lanes held in registers and clears them
// The reads in HBM are never modified.
// What bounds the kernel is the length of this chain of short dependent steps, not HBM: CTAs are
// kept small and the staging buffers are sized at launch from the mean work per item so that 7-8
// CTAs per SM overlap each other's barriers and latencies.
//
// Shared memory (dynamic), regions 16-byte aligned:
//   s_meta   3 x META_STRIDE u32      q4_off | seg_off | mate of the chunk's reads
//   s_seg    CHUNK_SEGS x 16 bytes    {first position, length, byte index of its first quality,
//                                      tile-relative index of its first quad}
//   s_cnt    5 x TILE bytes           planes A, C, G, T, non-ACGT: one byte per position
//   s_seq    chunk_q4 + 32 bytes      2-bit bases           (TMA destination)
//   s_qual   4*chunk_q4 + 32 bytes    qualities             (TMA destination)
//   s_g2s    chunk_q4 bytes           tile-relative index of every staged quad (low byte)
// ------------------------------------------------------------------------------------------------
constexpr int TILE_QUADS = TILE / 4;

__host__ __device__ constexpr size_t pileup_smem_bytes(int chunk_reads, uint32_t chunk_q4)
{
    return (size_t)(3 * ((chunk_reads + 4) / 4 * 4) * 4 + CHUNK_SEGS * 16 + ((chunk_reads + 4) / 4 * 4) * 2 + 64 + 5 * TILE) +
           (chunk_q4 + 32) + (4 * (size_t)chunk_q4 + 32) + (size_t)chunk_q4 + 32;
}

// explicit shared-window accesses (32-bit addresses): the compiler otherwise rebuilds generic
// pointers from the CTA's shared base in every iteration of the scatter loop
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
template <int OFF>
__device__ __forceinline__ void red_shared_add(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0+%2], %1;" :: "r"(a), "r"(v), "n"(OFF) : "memory"); }

template <int THREADS, int CHUNK_READS, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
pileup_kernel(const SampleDev* __restrict__ samples, const Item* __restrict__ items, uint32_t n_items, uint32_t chunk_q4,
              uint64_t* __restrict__ acgt /*[n_items][TILE]*/, uint16_t* __restrict__ ncnt /*[n_items][TILE]*/,
              int* __restrict__ err_flag)
{
    static_assert(TILE_QUADS == 256 && CHUNK_READS < THREADS && TILE_QUADS % THREADS == 0, "one thread per staged read; each thread folds whole quads");
    constexpr int QUADS_PER_THREAD = TILE_QUADS / THREADS;
    constexpr int META_STRIDE = (CHUNK_READS + 4) / 4 * 4;
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* s_q4   = (uint32_t*)smem;
    uint32_t* s_sgo  = s_q4 + META_STRIDE;
    int32_t*  s_mate = (int32_t*)(s_sgo + META_STRIDE);
    uint4*    s_seg  = (uint4*)(s_mate + META_STRIDE);
    uint16_t* s_pairs = (uint16_t*)(s_seg + CHUNK_SEGS);
    uint64_t* s_bar  = (uint64_t*)(s_pairs + META_STRIDE);
    uint32_t* s_misc = (uint32_t*)(s_bar + 1);        // [0] number of overlap tasks of the chunk
    uint32_t* s_cnt  = s_misc + 14;                   // 5 planes of TILE_QUADS words
    uint8_t*  s_seq  = (uint8_t*)(s_cnt + 5 * TILE_QUADS);
    uint8_t*  s_qual = s_seq + chunk_q4 + 32;
    uint8_t*  s_g2s  = s_qual + 4 * chunk_q4 + 32;                // 4-byte aligned

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const Item it = items[blockIdx.x];
    const int32_t p0 = (int32_t)(it.tile * TILE);
    const SampleDev* __restrict__ sd = samples + it.sample;

    if (tid == 0) { mbar_init(s_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int k = tid; k < 5 * TILE_QUADS; k += THREADS) s_cnt[k] = 0;

    // per letter and quad: 16-bit lanes, [0] = positions 0 and 2 of the quad, [1] = positions 1 and 3
    uint32_t acc[QUADS_PER_THREAD][5][2];
    #pragma unroll
    for (int k = 0; k < QUADS_PER_THREAD; ++k)
        #pragma unroll
        for (int c = 0; c < 5; ++c) { acc[k][c][0] = 0; acc[k][c][1] = 0; }
    uint32_t parity = 0;

    for (uint32_t c0 = it.r_lo; c0 < it.r_hi;) {
        // ---- 1. metadata of up to CHUNK_READS reads (+1 for the end offsets); the chunk takes the
        // longest prefix within the quad and segment budgets (prefix sums: the predicate is monotone)
        uint32_t n = it.r_hi - c0; if (n > CHUNK_READS) n = CHUNK_READS;
        bool fits = false;
        if (tid <= n) {
            const uint32_t* q4p = sd->q4_off + c0; const uint32_t* sgp = sd->seg_off + c0;
            const uint32_t q = __ldg(q4p + tid), g = __ldg(sgp + tid);
            s_q4[tid] = q; s_sgo[tid] = g;
            fits = tid >= 1 && q - __ldg(q4p) <= chunk_q4 && g - __ldg(sgp) <= CHUNK_SEGS;
            if (tid < n) s_mate[tid] = __ldg(sd->mate + c0 + tid);
        }
        if (tid == 0) s_misc[0] = 0;
        const uint32_t m = (uint32_t)__syncthreads_count(fits);
        if (m == 0) {                       // a single read over the documented limits: host validation failed
            if (tid == 0) atomicExch(err_flag, 1);
            break;
        }
        const uint32_t q4_0 = s_q4[0], sg_0 = s_sgo[0], nq4 = s_q4[m] - q4_0;

        // ---- 2. stage bases and qualities: two bulk copies from 16-byte aligned addresses at or
        // below the first byte needed; d_* is the offset of that byte in the buffer
        const uint8_t* g_seq = sd->seq2 + q4_0;
        const uint8_t* g_qual = sd->qual + (size_t)q4_0 * 4;
        const uint32_t d_seq = (uint32_t)((uintptr_t)g_seq & 15), d_qual = (uint32_t)((uintptr_t)g_qual & 15);
        if (tid == 0) {
            const uint32_t b_seq = (d_seq + nq4 + 15) & ~15u, b_qual = (d_qual + nq4 * 4 + 15) & ~15u;
            fence_proxy_async();            // earlier generic-proxy accesses to these buffers are ordered before the copies
            mbar_expect_tx(s_bar, b_seq + b_qual);
            if (b_seq)  tma_load_1d(s_seq, g_seq - d_seq, b_seq, s_bar);
            if (b_qual) tma_load_1d(s_qual, g_qual - d_qual, b_qual, s_bar);
        }

        // ---- 3. segment records while the copies are in flight: one thread per read
        if (tid < m) {
            const uint32_t k0 = s_sgo[tid] - sg_0, k1 = s_sgo[tid + 1] - sg_0;
            uint32_t q = s_q4[tid] - q4_0;                        // first staged quad of the next segment
            const uint32_t q_end = s_q4[tid + 1] - q4_0;
            for (uint32_t k = k0; k < k1; ++k) {
                const int32_t p = __ldg(sd->seg_pos + sg_0 + k);
                const uint32_t len = __ldg(sd->seg_len + sg_0 + k);
                const uint32_t a = (uint32_t)p & 3u, nq = (a + len + 3u) >> 2;
                const int32_t jw = (p - (int32_t)a - p0) >> 2;    // tile-relative index of the segment's first quad (may be off the tile)
                // (a segment whose quads run past the read's - inconsistent input, flagged below - gets length 0,
                // which keeps it out of the overlap step and of the zeroing in step 4a)
                s_seg[k] = make_uint4((uint32_t)p, q + nq <= q_end ? len : 0u, d_qual + q * 4u + a, (uint32_t)jw);
                uint32_t e = q + nq; if (e > q_end) e = q_end;
                {   // tag every quad with the low byte of its tile-relative index (TILE_QUADS == 256: a quad ON the tile
                    // is tagged exactly; quads off the tile get their qualities zeroed in step 4 and may carry any tag):
                    // bytes up to a word boundary, whole words of four consecutive tags, trailing bytes
                    uint32_t g = q, t = (uint32_t)jw;
                    for (; (g & 3u) && g < e; ++g, ++t) s_g2s[g] = (uint8_t)t;
                    for (; g + 4u <= e; g += 4u, t += 4u) {
                        const uint32_t t0 = t & 0xffu;
                        if (t0 <= 252u) *reinterpret_cast<uint32_t*>(s_g2s + g) = t0 * 0x01010101u + 0x03020100u;
                        else { s_g2s[g] = (uint8_t)t; s_g2s[g + 1] = (uint8_t)(t + 1); s_g2s[g + 2] = (uint8_t)(t + 2); s_g2s[g + 3] = (uint8_t)(t + 3); }
                    }
                    for (; g < e; ++g, ++t) s_g2s[g] = (uint8_t)t;
                }
                q += nq;
            }
            if (q != q_end) {                                     // offsets and segments disagree: refuse, stay in bounds
                atomicExch(err_flag, 2);
                for (uint32_t g = s_q4[tid] - q4_0; g < q_end; ++g) s_g2s[g] = 0;
            }
            const int32_t mt = s_mate[tid];
            if (mt >= 0) {                                        // overlap task: once per pair when both mates are here
                const bool mate_here = (uint32_t)mt >= c0 && (uint32_t)mt < c0 + m;
                if (!mate_here || c0 + tid < (uint32_t)mt) s_pairs[atomicAdd(&s_misc[0], 1u)] = (uint16_t)tid;
            }
        }
        if (tid == 0) mbar_wait(s_bar, parity);
        parity ^= 1;
        __syncthreads();                    // segments written, copies landed (thread 0 observed the barrier)

        // ---- 4a. quads off the tile (front of reads that start before it, tail of reads that run past it) must not
        // count: their qualities are zeroed, which the scatter's threshold test then rejects like any poor base
        if (tid < m) {
            const uint32_t k0 = s_sgo[tid] - sg_0, k1 = s_sgo[tid + 1] - sg_0;
            for (uint32_t k = k0; k < k1; ++k) {
                const uint4 sg = s_seg[k];
                if (!sg.y) continue;
                const int32_t jw = (int32_t)sg.w;
                const uint32_t a = sg.x & 3u;
                const int32_t nq = (int32_t)((a + sg.y + 3u) >> 2);
                if (jw >= 0 && jw + nq <= TILE_QUADS) continue;                    // wholly on the tile: the common case
                uint32_t* qw = reinterpret_cast<uint32_t*>(s_qual + (sg.z - a));    // the segment's first quad
                const int32_t lead = min(max(-jw, 0), nq), tail = min(max(TILE_QUADS - jw, 0), nq);
                for (int32_t i = 0; i < lead; ++i) qw[i] = 0;
                for (int32_t i = tail; i < nq; ++i) qw[i] = 0;
            }
        }

        // ---- 4. mate-overlap quality correction, restricted to this tile's positions (other tiles
        // are counted by other CTAs). Pairs with both mates in the chunk: both are rewritten from
        // pristine values. Mates outside the chunk (only when a tile needs several chunks): this read
        // alone is rewritten, the mate's pristine data come from global memory.
        const uint32_t n_tasks = s_misc[0];
        if (n_tasks) {
            // eight lanes per task (four tasks per warp): the work per pair is a few dozen shared-memory
            // bytes -- a whole warp per pair wastes ~200 instructions on set-up, one thread per pair
            // serialises ~50 dependent read-modify-writes
            const uint32_t l8 = lane & 7u;
            for (uint32_t t = tid >> 3; t < n_tasks; t += THREADS / 8) {
                const uint32_t i = s_pairs[t];
                const int32_t mt = s_mate[i];
                const bool self_is_a = c0 + i < (uint32_t)mt;
                const bool mate_here = (uint32_t)mt >= c0 && (uint32_t)mt < c0 + m;
                const uint32_t sa0 = s_sgo[i] - sg_0, sa1 = s_sgo[i + 1] - sg_0;
                if (mate_here) {
                    const uint32_t j = (uint32_t)mt - c0;
                    const uint32_t sb0 = s_sgo[j] - sg_0, sb1 = s_sgo[j + 1] - sg_0;
                    for (uint32_t ka = sa0; ka < sa1; ++ka) {
                        const uint4 A = s_seg[ka];
                        for (uint32_t kb = sb0; kb < sb1; ++kb) {
                            const uint4 B = s_seg[kb];
                            const int32_t lo = max(max((int32_t)A.x, (int32_t)B.x), p0);
                            const int32_t hi = min(min((int32_t)(A.x + A.y), (int32_t)(B.x + B.y)), p0 + TILE);
                            if (lo >= hi) continue;
                            // four positions per lane and step: both mates are stored position-aligned, so the
                            // quads of the common range line up word for word
                            const uint32_t za0 = A.z - (A.x & 3u) - 4u * (A.x >> 2), zb0 = B.z - (B.x & 3u) - 4u * (B.x >> 2);
                            for (int32_t P = (lo >> 2) + (int32_t)l8; P < ((hi + 3) >> 2); P += 8) {
                                const uint32_t za = za0 + 4u * (uint32_t)P, zb = zb0 + 4u * (uint32_t)P;   // byte index of the quad in s_qual
                                const uint32_t va = *reinterpret_cast<const uint32_t*>(s_qual + za), vb = *reinterpret_cast<const uint32_t*>(s_qual + zb);
                                const uint32_t xa = msnv_spread_bases(s_seq[d_seq + ((za - d_qual) >> 2)]);
                                const uint32_t xb = msnv_spread_bases(s_seq[d_seq + ((zb - d_qual) >> 2)]);
                                uint32_t na, nb;
                                msnv_overlap_rule4(va, vb, xa, xb, msnv_quad_mask(P << 2, lo, hi), na, nb);
                                *reinterpret_cast<uint32_t*>(s_qual + za) = na; *reinterpret_cast<uint32_t*>(s_qual + zb) = nb;
                            }
                        }
                    }
                } else {
                    // the mate is staged in another chunk: read its segments and pristine bytes from global memory
                    const uint32_t ms0 = __ldg(sd->seg_off + mt), ms1 = __ldg(sd->seg_off + mt + 1);
                    uint32_t mq = __ldg(sd->q4_off + mt);                  // first quad of the mate's next segment
                    for (uint32_t ks = ms0; ks < ms1; ++ks) {
                        const int32_t bx = __ldg(sd->seg_pos + ks);
                        const uint32_t bl = __ldg(sd->seg_len + ks), ba0 = (uint32_t)bx & 3u;
                        const uint32_t* mqual4 = reinterpret_cast<const uint32_t*>(sd->qual) + mq;   // the segment's first quad
                        const uint8_t* mseq = sd->seq2 + mq;
                        for (uint32_t ka = sa0; ka < sa1; ++ka) {
                            const uint4 A = s_seg[ka];
                            const int32_t lo = max(max((int32_t)A.x, bx), p0);
                            const int32_t hi = min(min((int32_t)(A.x + A.y), bx + (int32_t)bl), p0 + TILE);
                            if (lo >= hi) continue;
                            const uint32_t zs0 = A.z - (A.x & 3u) - 4u * (A.x >> 2);
                            for (int32_t P = (lo >> 2) + (int32_t)l8; P < ((hi + 3) >> 2); P += 8) {
                                const uint32_t zs = zs0 + 4u * (uint32_t)P, im = (uint32_t)(P - (bx >> 2));   // the mate's quad index
                                const uint32_t vs = *reinterpret_cast<const uint32_t*>(s_qual + zs), vm = __ldg(mqual4 + im);
                                const uint32_t xs = msnv_spread_bases(s_seq[d_seq + ((zs - d_qual) >> 2)]), xm = msnv_spread_bases(__ldg(mseq + im));
                                const uint32_t msk = msnv_quad_mask(P << 2, lo, hi);
                                uint32_t na, nb;
                                if (self_is_a) msnv_overlap_rule4(vs, vm, xs, xm, msk, na, nb); else msnv_overlap_rule4(vm, vs, xm, xs, msk, nb, na);
                                *reinterpret_cast<uint32_t*>(s_qual + zs) = na;
                            }
                        }
                        mq += (ba0 + bl + 3u) >> 2;
                    }
                }
            }
        }
        __syncthreads();                    // qualities are final (steps 4a and 4 wrote disjoint quads)

        // ---- 5. flat scatter over the staged quads
        {
            const uint32_t a_q = smem_u32(s_qual) + d_qual;               // d_qual is a multiple of 4
            const uint32_t a_s = smem_u32(s_seq) + d_seq, a_g = smem_u32(s_g2s);
            const uint32_t a_c = smem_u32(s_cnt);
            for (uint32_t g = tid; g < nq4; g += THREADS) {
                const uint32_t q = lds_u32(a_q + g * 4u);
                uint32_t x = lds_u8(a_s + g);
                const uint32_t j = lds_u8(a_g + g);                           // tile-relative quad
                x = (x * 4097u) & 0x000f000fu;                                // two 2-bit pairs per half word
                x = (x * 65u) & 0x03030303u;                                  // one base per byte lane
                const uint32_t v = (q & 0x7f7f7f7fu) + 0x73737373u;           // bit 7 of a lane: quality >= 13
                const uint32_t ok = ((v & ~q) >> 7) & 0x01010101u;            // ... and the base is A/C/G/T
                const uint32_t hi = x >> 1;
                const uint32_t a = a_c + j * 4u;
                red_shared_add<0>(a, ok & ~x & ~hi);
                red_shared_add<TILE>(a, ok & x & ~hi);
                red_shared_add<2 * TILE>(a, ok & ~x & hi);
                red_shared_add<3 * TILE>(a, ok & x & hi);
                const uint32_t nn = v & q & 0x80808080u;                      // rare: counted non-ACGT bases
                if (nn) red_shared_add<4 * TILE>(a, nn >> 7);
            }
        }
        __syncthreads();

        // ---- 6. fold this chunk's byte lanes (a position sees at most m <= 255 reads per chunk)
        #pragma unroll
        for (int k = 0; k < QUADS_PER_THREAD; ++k) {
            #pragma unroll
            for (int c = 0; c < 5; ++c) {
                const uint32_t slot = c * TILE_QUADS + k * THREADS + tid;
                const uint32_t w = s_cnt[slot];
                if (w) {
                    s_cnt[slot] = 0;
                    acc[k][c][0] += w & 0x00ff00ffu;
                    acc[k][c][1] += (w >> 8) & 0x00ff00ffu;
                }
            }
        }
        c0 += m;
        // no barrier here: the next chunk only touches the counters again after two more barriers
    }

    // ---- 7. flush: 8 B + 2 B per position; a thread owns four consecutive positions
    #pragma unroll
    for (int k = 0; k < QUADS_PER_THREAD; ++k) {
        const size_t o = (size_t)blockIdx.x * TILE + 4u * (k * THREADS + tid);
        uint32_t w[4][2], nw[2];
        #pragma unroll
        for (int i = 0; i < 4; ++i) {
            // position i of the quad sits in lane (i >> 1) of acc[..][i & 1]
            const uint32_t sh = 16u * (uint32_t)(i >> 1);
            const uint32_t A = (acc[k][0][i & 1] >> sh) & 0xffffu, C = (acc[k][1][i & 1] >> sh) & 0xffffu;
            const uint32_t G = (acc[k][2][i & 1] >> sh) & 0xffffu, T = (acc[k][3][i & 1] >> sh) & 0xffffu;
            w[i][0] = A | C << 16; w[i][1] = G | T << 16;
        }
        nw[0] = ((acc[k][4][0]) & 0xffffu) | (acc[k][4][1] & 0xffffu) << 16;
        nw[1] = (acc[k][4][0] >> 16) | (acc[k][4][1] >> 16) << 16;
        uint4* dst = reinterpret_cast<uint4*>(acgt + o);
        dst[0] = make_uint4(w[0][0], w[0][1], w[1][0], w[1][1]);
        dst[1] = make_uint4(w[2][0], w[2][1], w[3][0], w[3][1]);
        *reinterpret_cast<uint2*>(ncnt + o) = make_uint2(nw[0], nw[1]);
    }
}

// ------------------------------------------------------------------------------------------------
// call: one CTA per tile, one thread per position. Sums the per-sample counts of the tile's items
// and applies snpCall's tests (call_vC.cpp:545-601). Output: one flag byte per position
// (low nibble: population mask, high nibble: individual mask, allele order A,C,G,T).
// ------------------------------------------------------------------------------------------------
// reference character -> channel of the base mpileup renders as '.'/',' (0..3 = A,C,G,T, 4 = N-like,
// 5 = none: IUPAC ambiguity codes match no read base). Derived from SAMv1's nt16 coding, the table
// mpileup's pileup_seq compares with (SURVEY.md Annex A.4).
__device__ __forceinline__ uint32_t ref_channel(uint32_t c)
{
    switch (c) {
        case 'A': case 'a': case '0': return 0;
        case 'C': case 'c': case '1': return 1;
        case 'G': case 'g': case '2': return 2;
        case 'T': case 't': case '3': return 3;
        case 'M': case 'm': case 'R': case 'r': case 'S': case 's': case 'V': case 'v': case 'W': case 'w':
        case 'Y': case 'y': case 'H': case 'h': case 'K': case 'k': case 'D': case 'd': case 'B': case 'b':
        case '=': return 5;
        default: return 4;
    }
}

__global__ void __launch_bounds__(TILE, 2048 / TILE)
call_kernel(const uint64_t* __restrict__ acgt, const uint16_t* __restrict__ ncnt, const uint32_t* __restrict__ tile_begin,
            const uint8_t* __restrict__ ref, CallParamsDev prm, int text_mode, uint8_t* __restrict__ flags,
            uint32_t* __restrict__ tile_hits)
{
    __shared__ uint32_t s_warp[33];
    const uint32_t t = blockIdx.x, tid = threadIdx.x;
    const uint32_t i0 = tile_begin[t], i1 = tile_begin[t + 1];
    const size_t p = (size_t)t * TILE + tid;
    const int32_t thr = prm.thr;
    // any: bit a set when some sample has count[a] >= thr (samples without reads count 0, which only
    // matters for the degenerate thr <= 0)
    uint32_t sum[4] = {0, 0, 0, 0}, sum_n = 0, any = thr <= 0 ? 15u : 0u;
    uint32_t i = i0;
    for (; i + 4 <= i1; i += 4) {                              // 4 independent loads in flight per thread
        uint64_t w[4]; uint32_t nn[4];
        #pragma unroll
        for (int u = 0; u < 4; ++u) { w[u] = __ldg(acgt + (size_t)(i + u) * TILE + tid); nn[u] = __ldg(ncnt + (size_t)(i + u) * TILE + tid); }
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            #pragma unroll
            for (int a = 0; a < 4; ++a) {
                const uint32_t c = (uint32_t)(w[u] >> (16 * a)) & 0xffffu;
                sum[a] += c;
                any |= ((int32_t)c >= thr ? 1u : 0u) << a;
            }
            sum_n += nn[u];
        }
    }
    for (; i < i1; ++i) {
        const uint64_t w = __ldg(acgt + (size_t)i * TILE + tid);
        #pragma unroll
        for (int a = 0; a < 4; ++a) {
            const uint32_t c = (uint32_t)(w >> (16 * a)) & 0xffffu;
            sum[a] += c;
            any |= ((int32_t)c >= thr ? 1u : 0u) << a;
        }
        sum_n += __ldg(ncnt + (size_t)i * TILE + tid);
    }
    uint32_t flag = 0;
    const uint32_t rc = ref[p];
    if (rc != 0 && i1 > i0) {
        // text mode (counts parsed from mpileup text): letters are never the reference's own base and the
        // fifth plane holds the '.'/',' matches, so nothing is masked (call_vC.cpp:545,550,583-584)
        const uint32_t ch = text_mode ? 6u : ref_channel(rc);
        const int64_t cov = (int64_t)sum[0] + sum[1] + sum[2] + sum[3] + ((ch == 4 || text_mode) ? sum_n : 0);
        int64_t nonref = 0;
        #pragma unroll
        for (int a = 0; a < 4; ++a) if ((uint32_t)a != ch) nonref += sum[a];
        // call_vC.cpp:547-552. (int) casts mirror the reference's `int cov`.
        if (cov >= prm.min_cov && nonref >= thr) {
            const double lim = __dmul_rn((double)cov, prm.frac);       // cov*calling_min_fraction, no FMA contraction
            #pragma unroll
            for (int a = 0; a < 4; ++a) {
                // call_vC.cpp:580: the allele is skipped only when the reference character is its lower-case letter
                const uint32_t lower = a == 0 ? 'a' : a == 1 ? 'c' : a == 2 ? 'g' : 't';
                if (rc == lower) continue;
                const int64_t n = ((uint32_t)a == ch) ? 0 : (int64_t)sum[a];
                if (n >= thr && (double)n >= lim) flag |= 1u << a;                       // population variant
                else {
                    const bool indiv = ((uint32_t)a == ch) ? (0 >= thr) : ((any >> a) & 1u);
                    if (indiv) flag |= 16u << a;                                         // individual variant
                }
            }
        }
    }
    flags[p] = (uint8_t)flag;
    uint32_t total;
    block_rank(flag != 0, s_warp, total);
    if (tid == 0) tile_hits[t] = total;
}

// ordered compaction of the flagged positions: tile_hits holds exclusive offsets on entry
__global__ void __launch_bounds__(TILE)
compact_kernel(const uint8_t* __restrict__ flags, const uint32_t* __restrict__ tile_hits, uint32_t* __restrict__ hit_pos,
               uint8_t* __restrict__ hit_pop, uint8_t* __restrict__ hit_ind)
{
    __shared__ uint32_t s_warp[33];
    const uint32_t t = blockIdx.x, tid = threadIdx.x;
    const size_t p = (size_t)t * TILE + tid;
    const uint32_t f = flags[p];
    uint32_t total;
    const uint32_t r = block_rank(f != 0, s_warp, total);
    if (f) {
        const uint32_t slot = tile_hits[t] + r;
        hit_pos[slot] = (uint32_t)p;
        hit_pop[slot] = (uint8_t)(f & 15u);
        hit_ind[slot] = (uint8_t)(f >> 4);
    }
}

// per hit: per-sample coverage and allele counts (zero for samples without an item on the tile)
// plus population totals. One CTA of 128 threads per hit; outputs were zero-filled by the host side.
__global__ void __launch_bounds__(128)
gather_kernel(const uint64_t* __restrict__ acgt, const uint16_t* __restrict__ ncnt, const Item* __restrict__ items,
              const uint32_t* __restrict__ tile_begin, const uint8_t* __restrict__ ref, const uint32_t* __restrict__ hit_pos,
              uint32_t n_samples, int text_mode, uint16_t* __restrict__ cov, uint16_t* __restrict__ allele,
              uint32_t* __restrict__ total)
{
    __shared__ uint32_t s_tot[5];
    const uint32_t h = blockIdx.x, tid = threadIdx.x;
    const uint32_t p = hit_pos[h];
    const uint32_t t = p / TILE, off = p % TILE;
    const uint32_t ch = text_mode ? 6u : ref_channel(ref[p]);
    if (tid < 5) s_tot[tid] = 0;
    __syncthreads();
    uint32_t tc = 0, ta[4] = {0, 0, 0, 0};
    for (uint32_t i = tile_begin[t] + tid; i < tile_begin[t + 1]; i += blockDim.x) {
        const uint32_t s = items[i].sample;
        const uint64_t w = __ldg(acgt + (size_t)i * TILE + off);
        const uint32_t nn = __ldg(ncnt + (size_t)i * TILE + off);
        uint32_t c[4], cv = (ch == 4 || text_mode) ? nn : 0;
        #pragma unroll
        for (int a = 0; a < 4; ++a) { c[a] = (uint32_t)(w >> (16 * a)) & 0xffffu; cv += c[a]; if ((uint32_t)a == ch) c[a] = 0; }
        cov[(size_t)h * n_samples + s] = (uint16_t)cv;
        #pragma unroll
        for (int a = 0; a < 4; ++a) { allele[((size_t)h * 4 + a) * n_samples + s] = (uint16_t)c[a]; ta[a] += c[a]; }
        tc += cv;
    }
    atomicAdd(&s_tot[0], tc);
    #pragma unroll
    for (int a = 0; a < 4; ++a) atomicAdd(&s_tot[1 + a], ta[a]);
    __syncthreads();
    if (tid < 5) total[(size_t)h * 5 + tid] = s_tot[tid];
}

// inspection hook: expand one sample's counts for a position range into [n][5] u16
__global__ void counts_kernel(const uint64_t* __restrict__ acgt, const uint16_t* __restrict__ ncnt, const Item* __restrict__ items,
                              const uint32_t* __restrict__ tile_begin, uint32_t sample, uint32_t first, uint32_t n,
                              uint16_t* __restrict__ out)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t p = first + k, t = p / TILE, off = p % TILE;
    uint16_t v[5] = {0, 0, 0, 0, 0};
    for (uint32_t i = tile_begin[t]; i < tile_begin[t + 1]; ++i) {
        if (items[i].sample != sample) continue;
        const uint64_t w = acgt[(size_t)i * TILE + off];
        for (int a = 0; a < 4; ++a) v[a] = (uint16_t)(w >> (16 * a));
        v[4] = ncnt[(size_t)i * TILE + off];
        break;
    }
    for (int a = 0; a < 5; ++a) out[(size_t)k * 5 + a] = v[a];
}

// ------------------------------------------------------------------------------------------------
// coverage (qaCompute.cpp:530-552 scatter, :142-165 prefix sum + histogram).
// Contigs that have reads are laid out in a "coverage coordinate" space, each starting at a multiple
// of COV_CHUNK. A block [beg,end) contributes +1 at beg and -1 at end like the reference's
// difference array, but every chunk it enters gets its own +1 at the chunk's first index, so each
// chunk is prefix-summed independently by one CTA: no carries between CTAs, one pass over HBM.
// ------------------------------------------------------------------------------------------------
constexpr int COV_CHUNK = 4096;
constexpr int COV_THREADS = 1024;
constexpr int COV_MAX_BINS = 1024;

__global__ void __launch_bounds__(256)
cov_scatter_kernel(const uint32_t* __restrict__ beg, const uint32_t* __restrict__ end, const uint64_t* __restrict__ blk_off,
                   const uint32_t* __restrict__ contig_chunk0, uint32_t n_contigs, uint64_t n_blocks, int32_t* __restrict__ diff)
{
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_blocks) return;
    uint32_t lo = 0, hi = n_contigs;                       // contig of this block: blk_off[lo] <= g < blk_off[lo+1]
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (__ldg(blk_off + mid) <= g) lo = mid; else hi = mid; }
    const uint32_t b = __ldg(beg + g), e = __ldg(end + g);
    if (e <= b) return;
    int32_t* d = diff + (size_t)__ldg(contig_chunk0 + lo) * COV_CHUNK;
    for (uint32_t c = b / COV_CHUNK; c <= (e - 1) / COV_CHUNK; ++c) {
        const uint32_t c0 = c * COV_CHUNK;
        atomicAdd(d + (b > c0 ? b : c0), 1);
        if (e < c0 + COV_CHUNK) atomicAdd(d + e, -1);
    }
}

__global__ void __launch_bounds__(COV_THREADS)
cov_scan_kernel(const int32_t* __restrict__ diff, const uint32_t* __restrict__ chunk_contig, const uint32_t* __restrict__ contig_chunk0,
                const uint32_t* __restrict__ contig_len, uint32_t max_cov, unsigned long long* __restrict__ cov_sum,
                unsigned long long* __restrict__ hist /*[n_contigs][max_cov+1]*/)
{
    __shared__ int32_t s_warp[32];
    __shared__ uint32_t s_hist[COV_MAX_BINS];
    __shared__ unsigned long long s_sum;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t k = chunk_contig[blockIdx.x];
    const uint32_t first = (blockIdx.x - contig_chunk0[k]) * COV_CHUNK;        // contig coordinate of the chunk's first index
    const uint32_t len = contig_len[k];
    for (uint32_t i = tid; i <= max_cov; i += COV_THREADS) s_hist[i] = 0;
    if (tid == 0) s_sum = 0;
    const int4 v = reinterpret_cast<const int4*>(diff + (size_t)blockIdx.x * COV_CHUNK)[tid];
    int32_t c0 = v.x, c1 = c0 + v.y, c2 = c1 + v.z, c3 = c2 + v.w;
    int32_t incl = c3;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int32_t w = s_warp[lane], wi = w;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int32_t o = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += o; }
        s_warp[lane] = wi - w;
    }
    __syncthreads();
    const int32_t base = s_warp[warp] + incl - c3;
    int32_t cv[4] = {base + c0, base + c1, base + c2, base + c3};
    unsigned long long local = 0;
    #pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t p = first + tid * 4 + j;
        const bool valid = p < len;
        // negative coverage cannot occur for blocks that satisfy the ABI contract (beg < end)
        uint32_t bin = cv[j] < 0 ? 0u : ((uint32_t)cv[j] > max_cov ? max_cov : (uint32_t)cv[j]);
        if (!valid) bin = 0xffffffffu;
        else local += (unsigned long long)(cv[j] < 0 ? 0 : cv[j]);
        const uint32_t peers = __match_any_sync(0xffffffffu, bin);
        if (valid && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&s_hist[bin], (uint32_t)__popc(peers));
    }
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) local += __shfl_down_sync(0xffffffffu, local, d);
    if (lane == 0 && local) atomicAdd(&s_sum, local);
    __syncthreads();
    for (uint32_t i = tid; i <= max_cov; i += COV_THREADS)
        if (s_hist[i]) atomicAdd(hist + (size_t)k * (max_cov + 1) + i, (unsigned long long)s_hist[i]);
    if (tid == 0 && s_sum) atomicAdd(cov_sum + k, s_sum);
}

}  // namespace msnv_gpu
