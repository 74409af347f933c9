// kernels.cuh -- sm_100a kernels of the pileup + call path (device side of include/msnv.h).
//
// Data flow for one shard (all samples of one genome bin):
//   index_kernel     (tile, sample) pairs that have reads  -> ordered work items (ballot compaction)
//   pileup_kernel    per item: stage reads (TMA bulk copies) -> CIGAR walk + mate-overlap quality
//                    correction (SURVEY.md Annex A.2) in shared memory -> per-position gather
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

namespace msnv_gpu {

constexpr int TILE = MSNV_TILE;                 // positions per tile == threads per pileup CTA
constexpr int PILEUP_THREADS = 256;             // threads per pileup CTA (each folds TILE/256 positions)
// reads staged per chunk: <= 255 (8-bit per-chunk counters, one walk thread per read). Two
// instantiations of the pileup kernel: a small one for shallow data (8 CTAs per SM) and a large
// one for deep data (fewer, longer chunks per tile)
constexpr int CHUNK_READS_SMALL = 127, CHUNK_READS_LARGE = 255;
constexpr int PILEUP_CTAS_SMALL = TILE >= 1024 ? 6 : 8, PILEUP_CTAS_LARGE = TILE >= 1024 ? 4 : 5;   // launch-bound targets (register budget)
constexpr int CHUNK_Q4_MAX = 4096;              // upper bound of the 4-base groups staged per chunk (chosen per launch)
constexpr int CHUNK_Q4_MIN = MSNV_MAX_READ_BASES / 4;   // a single read always fits
constexpr int CHUNK_SEGS_SMALL = 256, CHUNK_SEGS_LARGE = 512;   // aligned segments per chunk

static_assert(TILE % PILEUP_THREADS == 0 && PILEUP_THREADS >= 256, "one walk thread per staged read");
static_assert(MSNV_MAX_READ_CIGAR * 2 <= CHUNK_SEGS_SMALL, "one read's segments must fit a chunk");

struct SampleDev {
    const int32_t*  pos;
    const uint32_t* cig_off;
    const uint32_t* seg_off;
    const uint32_t* q4_off;
    const int32_t*  mate;
    const uint32_t* cigar;
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
// 1 << s for 64-bit with PTX clamping semantics (s >= 64 gives 0)
__device__ __forceinline__ uint64_t shl1_clamped(uint32_t s)
{
    uint64_t r;
    asm("shl.b64 %0, 1, %1;" : "=l"(r) : "r"(s));
    return r;
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
                const int64_t t0 = (int64_t)t * TILE;
                r_lo = lower_bound_i32(pos, n, t0 - (int64_t)samples[s].max_span + 1);
                r_hi = lower_bound_i32(pos, n, t0 + TILE);
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
// 1 << s with PTX clamping semantics (s >= 32 gives 0)
__device__ __forceinline__ uint32_t shl1_clamped32(uint32_t s)
{
    uint32_t r;
    asm("shl.b32 %0, 1, %1;" : "=r"(r) : "r"(s));
    return r;
}

constexpr uint32_t CODE_N = 32, CODE_SKIP = 64;

// mpileup's mate-overlap rule (htslib tweak_overlap_quality, SURVEY.md Annex A.2) for one reference
// position that both mates align to. va/vb: staged quality bytes (bit 7: non-ACGT base) of the mate
// that comes first in the file (a) and of the later one (b); same: the two bases are equal.
// Qualities live in 7 bits, so htslib's cap of 200 becomes 127: only "q >= 13" is ever used.
__device__ __forceinline__ void overlap_rule(uint32_t va, uint32_t vb, bool same, uint32_t& na, uint32_t& nb)
{
    const uint32_t fa = va & 0x80u, fb = vb & 0x80u, qa = va & 0x7fu, qb = vb & 0x7fu;
    if (same) { uint32_t q = qa + qb; if (q > 127u) q = 127u; na = fa | q; nb = fb; }
    else if (qa >= qb) { na = fa | (uint32_t)(0.8 * (double)qa); nb = fb; }
    else { na = fa; nb = fb | (uint32_t)(0.8 * (double)qb); }
}

// ------------------------------------------------------------------------------------------------
// pileup: one CTA of PILEUP_THREADS threads per work item (sample, tile of TILE positions).
// Per chunk of reads (<= CHUNK_READS reads, chunk_q4*4 bases, CHUNK_SEGS aligned segments):
//   1. metadata of the chunk's reads -> shared memory; __syncthreads_count sizes the chunk
//   2. two TMA bulk copies (2-bit bases, qualities) are issued; while they fly,
//   3. one thread per read walks its CIGAR (global, L2) into aligned segments and a descriptor of
//      its first segment clipped to the tile, and tags its 4-base groups with the read's index
//   4. mate-overlap quality correction in shared memory, one warp per pair, lanes over positions
//      (mates staged in another chunk are read, pristine, from global memory)
//   5. flat scatter: thread g takes the g-th 4-base group of the staged bytes, turns (2-bit base,
//      quality) into four codes with SWAR arithmetic and adds each base that lies on the tile to
//      its position's shared-memory counter (one byte lane per base letter) with a predicated
//      red.shared.add. Every staged base costs the same handful of instructions at full lane
//      occupancy, whatever the depth; the counter words are XOR-swizzled so the stride-4 access of
//      a warp is conflict free (measured: >= 21 shared atomics per clock per SM,
//      profiles/r01_microbench_shared_atomics.txt)
//   6. each thread folds the byte lanes of its positions into 16-bit packed registers and clears them
// The reads in HBM are never modified. CTAs are small (8 warps) and the staging buffers are sized
// at launch from the mean work per item, so that 6-8 CTAs per SM overlap each other's barriers
// and copy latency.
//
// Shared memory (dynamic), regions 16-byte aligned; the counters are aligned to their own size:
//   s_meta   4 x META_STRIDE u32      pos | q4_off | seg_off | mate of the chunk's reads
//   s_rd     256 x 16 bytes           {first segment: p_rel - base index, lo, span; first seg | n segs << 16}
//   s_seg    CHUNK_SEGS x 16 bytes    {ref begin, length, byte address of first quality, read index}
//   s_cnt    2 x TILE u32             A|C|G|T byte lanes, non-ACGT count
//   s_seq    chunk_q4 + 32 bytes      2-bit bases           (TMA destination)
//   s_qual   4*chunk_q4 + 32 bytes    qualities             (TMA destination)
//   s_g2r    chunk_q4 bytes           read index of every 4-base group
// Codes are shift amounts: 0,8,16,24 = A,C,G,T with quality >= 13; 32 = non-ACGT base with
// quality >= 13; 64 = not counted.
// ------------------------------------------------------------------------------------------------
constexpr int PILEUP_POS_PER_THREAD = TILE / PILEUP_THREADS;
__host__ __device__ constexpr size_t pileup_smem_bytes(int chunk_reads, int chunk_segs, uint32_t chunk_q4)
{
    return (size_t)(4 * ((chunk_reads + 4) / 4 * 4) * 4 + ((chunk_reads + 4) / 4 * 4) * 16 + chunk_segs * 16 + ((chunk_reads + 4) / 4 * 4) * 2 + 64 +
                    (2 * TILE * 4) + TILE * 4 /*alignment slack*/) +
           (chunk_q4 + 32) + (4 * (size_t)chunk_q4 + 32) + chunk_q4 + 32;
}

// explicit shared-window accesses (32-bit addresses): the compiler otherwise rebuilds generic
// pointers from the CTA's shared base in every iteration of the scatter loop
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds_v4(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void red_shared_add(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }

// counter word of tile-relative position p: XOR swizzle so that positions 4 apart fall in different banks
__device__ __forceinline__ uint32_t cnt_slot(uint32_t p) { return p ^ ((p >> 5) & 3u); }

// if (a < b) shared[addr] += val, without a branch
__device__ __forceinline__ void red_shared_add_if_lt(uint32_t addr, uint32_t val, uint32_t a, uint32_t b)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %2, %3;\n\t@p red.shared.add.u32 [%0], %1;\n\t}"
                 :: "r"(addr), "r"(val), "r"(a), "r"(b) : "memory");
}

template <int CHUNK_READS, int CHUNK_SEGS, int MIN_CTAS>
__global__ void __launch_bounds__(PILEUP_THREADS, MIN_CTAS)
pileup_kernel(const SampleDev* __restrict__ samples, const Item* __restrict__ items, uint32_t n_items, uint32_t chunk_q4,
              uint64_t* __restrict__ acgt /*[n_items][TILE]*/, uint16_t* __restrict__ ncnt /*[n_items][TILE]*/,
              int* __restrict__ err_flag)
{
    constexpr int META_STRIDE = (CHUNK_READS + 4) / 4 * 4, RD_SLOTS = META_STRIDE;
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* s_pos = (uint32_t*)smem;
    uint32_t* s_q4  = s_pos + META_STRIDE;
    uint32_t* s_sgo = s_q4 + META_STRIDE;
    int32_t*  s_mate = (int32_t*)(s_sgo + META_STRIDE);
    uint4*    s_rd   = (uint4*)(s_mate + META_STRIDE);
    uint4*    s_seg  = s_rd + RD_SLOTS;
    uint16_t* s_pairs = (uint16_t*)(s_seg + CHUNK_SEGS);
    uint64_t* s_bar  = (uint64_t*)(s_pairs + RD_SLOTS);
    uint32_t* s_misc = (uint32_t*)(s_bar + 1);        // [0] number of overlap tasks of the chunk
    // counters: aligned to TILE*4 bytes in the shared window so that base | offset == base + offset
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t after_fixed = smem_u32(s_misc) + 56;
    const uint32_t cnt_base = (after_fixed + TILE * 4 - 1) & ~(uint32_t)(TILE * 4 - 1);
    uint32_t* s_cnt  = (uint32_t*)(smem + (cnt_base - smem_base));
    uint32_t* s_cntn = s_cnt + TILE;
    uint8_t*  s_seq  = (uint8_t*)(s_cntn + TILE);
    uint8_t*  s_qual = s_seq + chunk_q4 + 32;
    uint8_t*  s_g2r  = s_qual + 4 * chunk_q4 + 32;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Item it = items[blockIdx.x];
    const int32_t p0 = (int32_t)(it.tile * TILE);
    const SampleDev* __restrict__ sd = samples + it.sample;

    if (tid == 0) { mbar_init(s_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    #pragma unroll
    for (int k = 0; k < PILEUP_POS_PER_THREAD; ++k) { s_cnt[tid + k * PILEUP_THREADS] = 0; s_cntn[tid + k * PILEUP_THREADS] = 0; }

    uint64_t acc[PILEUP_POS_PER_THREAD];      // A | C<<16 | G<<32 | T<<48 of positions p0 + tid + k*PILEUP_THREADS
    uint32_t acc_n[PILEUP_POS_PER_THREAD];
    #pragma unroll
    for (int k = 0; k < PILEUP_POS_PER_THREAD; ++k) { acc[k] = 0; acc_n[k] = 0; }
    uint32_t parity = 0;

    for (uint32_t c0 = it.r_lo; c0 < it.r_hi;) {
        // ---- 1. metadata of up to CHUNK_READS reads (+1 for the end offsets); the chunk takes the
        // longest prefix within the byte and segment budgets (prefix sums: the predicate is monotone)
        uint32_t n = it.r_hi - c0; if (n > CHUNK_READS) n = CHUNK_READS;
        bool fits = false;
        if (tid <= n) {
            const uint32_t* q4p = sd->q4_off + c0; const uint32_t* sgp = sd->seg_off + c0;
            const uint32_t q = __ldg(q4p + tid), g = __ldg(sgp + tid);
            s_q4[tid] = q; s_sgo[tid] = g;
            fits = tid >= 1 && q - __ldg(q4p) <= chunk_q4 && g - __ldg(sgp) <= CHUNK_SEGS;
            if (tid < n) { s_pos[tid] = (uint32_t)__ldg(sd->pos + c0 + tid); s_mate[tid] = __ldg(sd->mate + c0 + tid); }
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

        // ---- 3. CIGAR walk while the copies are in flight: one thread per read
        if (tid < m) {
            const uint32_t cg0 = __ldg(sd->cig_off + c0 + tid), nops = __ldg(sd->cig_off + c0 + tid + 1) - cg0;
            const uint32_t k0 = s_sgo[tid] - sg_0;
            uint32_t k = k0;
            const uint32_t gq0 = s_q4[tid] - q4_0, gq1 = s_q4[tid + 1] - q4_0;
            int32_t x = (int32_t)s_pos[tid];
            uint32_t y = d_qual + gq0 * 4;                        // byte address of the read's first base in s_qual
            for (uint32_t o = 0; o < nops; ++o) {
                const uint32_t w = __ldg(sd->cigar + cg0 + o), op = w & 0xf, len = w >> 4;
                if (op == 0 || op == 7 || op == 8) {
                    s_seg[k++] = make_uint4((uint32_t)x, len, y, tid);
                    x += (int32_t)len; y += len;
                } else if (op == 2 || op == 3) x += (int32_t)len;
                else if (op == 1 || op == 4) y += len;
            }
            // first segment clipped to the tile, in the coordinates the scatter uses: a base with staged
            // index b (4*group + k) lies at tile-relative position rd.x + b if (rd.x + b - rd.y) < rd.z
            uint4 rd = make_uint4(0u, 0u, 0u, k0 | (k - k0) << 16);
            if (k > k0) {
                const uint4 f = s_seg[k0];
                const int32_t lo = max((int32_t)f.x - p0, 0), hi = min((int32_t)(f.x + f.y) - p0, (int32_t)TILE);
                rd.x = (uint32_t)((int32_t)f.x - p0 - (int32_t)(f.z - d_qual));
                rd.y = (uint32_t)lo;
                rd.z = hi > lo ? (uint32_t)(hi - lo) : 0u;
            }
            s_rd[tid] = rd;
            for (uint32_t g = gq0; g < gq1; ++g) s_g2r[g] = (uint8_t)tid;
            const int32_t mt = s_mate[tid];
            if (mt >= 0) {                                        // overlap task: once per pair when both mates are here
                const bool mate_here = (uint32_t)mt >= c0 && (uint32_t)mt < c0 + m;
                if (!mate_here || c0 + tid < (uint32_t)mt) s_pairs[atomicAdd(&s_misc[0], 1u)] = (uint16_t)tid;
            }
        }
        if (tid == 0) mbar_wait(s_bar, parity);
        parity ^= 1;
        __syncthreads();                    // segments written, copies landed (thread 0 observed the barrier)

        // ---- 4. mate-overlap quality correction, restricted to this tile's positions (other tiles
        // are counted by other CTAs). Pairs with both mates in the chunk: one warp rewrites both from
        // pristine values. Mates outside the chunk (only when a tile needs several chunks): this read
        // alone is rewritten, the mate's pristine data come from global memory.
        const uint32_t n_tasks = s_misc[0];
        if (n_tasks) {
            // eight lanes per task (four tasks per warp): the work per pair is a few dozen shared-memory
            // bytes -- a whole warp per pair wastes ~200 instructions on set-up, one thread per pair
            // serialises ~50 dependent read-modify-writes
            const uint32_t l8 = lane & 7u;
            for (uint32_t t = tid >> 3; t < n_tasks; t += PILEUP_THREADS / 8) {
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
                            for (int32_t p = lo + (int32_t)l8; p < hi; p += 8) {
                                const uint32_t za = A.z + (uint32_t)(p - (int32_t)A.x), zb = B.z + (uint32_t)(p - (int32_t)B.x);
                                const uint32_t va = s_qual[za], vb = s_qual[zb];
                                const uint32_t ia = za - d_qual, ib = zb - d_qual;     // base index inside the staged range
                                const uint32_t ba = (s_seq[d_seq + (ia >> 2)] >> ((ia & 3) * 2)) & 3u;
                                const uint32_t bb = (s_seq[d_seq + (ib >> 2)] >> ((ib & 3) * 2)) & 3u;
                                const bool same = ((va | vb) & 0x80u) ? ((va & vb & 0x80u) != 0) : (ba == bb);
                                uint32_t na, nb;
                                overlap_rule(va, vb, same, na, nb);
                                s_qual[za] = (uint8_t)na; s_qual[zb] = (uint8_t)nb;
                            }
                        }
                    }
                } else {
                    // the mate is staged in another chunk: walk its CIGAR and read its pristine bytes from global memory
                    const uint32_t mc0 = __ldg(sd->cig_off + mt), mn = __ldg(sd->cig_off + mt + 1) - mc0;
                    const uint32_t mq4 = __ldg(sd->q4_off + mt);
                    const uint8_t* mq = sd->qual + (size_t)mq4 * 4;
                    const uint8_t* ms = sd->seq2 + mq4;
                    int32_t bx = __ldg(sd->pos + mt); uint32_t by = 0;
                    for (uint32_t o = 0; o < mn; ++o) {
                        const uint32_t w = __ldg(sd->cigar + mc0 + o), op = w & 0xf, len = w >> 4;
                        if (op == 0 || op == 7 || op == 8) {
                            for (uint32_t ka = sa0; ka < sa1; ++ka) {
                                const uint4 A = s_seg[ka];
                                const int32_t lo = max(max((int32_t)A.x, bx), p0);
                                const int32_t hi = min(min((int32_t)(A.x + A.y), bx + (int32_t)len), p0 + TILE);
                                for (int32_t p = lo + (int32_t)l8; p < hi; p += 8) {
                                    const uint32_t zs = A.z + (uint32_t)(p - (int32_t)A.x), im = by + (uint32_t)(p - bx);
                                    const uint32_t vs = s_qual[zs], vm = mq[im];
                                    const uint32_t is = zs - d_qual;
                                    const uint32_t bs = (s_seq[d_seq + (is >> 2)] >> ((is & 3) * 2)) & 3u;
                                    const uint32_t bm = (ms[im >> 2] >> ((im & 3) * 2)) & 3u;
                                    const bool same = ((vs | vm) & 0x80u) ? ((vs & vm & 0x80u) != 0) : (bs == bm);
                                    uint32_t na, nb;
                                    if (self_is_a) overlap_rule(vs, vm, same, na, nb); else overlap_rule(vm, vs, same, nb, na);
                                    s_qual[zs] = (uint8_t)na;
                                }
                            }
                            bx += (int32_t)len; by += len;
                        } else if (op == 2 || op == 3) bx += (int32_t)len;
                        else if (op == 1 || op == 4) by += len;
                    }
                }
            }
            __syncthreads();
        }

        // ---- 5. flat scatter over the staged 4-base groups
        {
            const uint32_t a_q = smem_u32(s_qual) + d_qual;               // d_qual is a multiple of 4
            const uint32_t a_s = smem_u32(s_seq) + d_seq, a_g = smem_u32(s_g2r), a_rd = smem_u32(s_rd);
            const uint32_t a_n = smem_u32(s_cntn);
            for (uint32_t g = tid; g < nq4; g += PILEUP_THREADS) {
                const uint32_t q = lds_u32(a_q + g * 4u);
                uint32_t x = lds_u8(a_s + g);
                const uint4 rd = lds_v4(a_rd + lds_u8(a_g + g) * 16u);
                // four codes at once (SWAR): no byte lane can carry into its neighbour
                x = (x * 4097u) & 0x000f000fu;                    // two 2-bit pairs per half word
                x = (x * 520u) & 0x18181818u;                     // base*8 in every byte lane
                const uint32_t pass = (((q & 0x7f7f7f7fu) + 0x73737373u) >> 7) & 0x01010101u;   // 1 where (q & 127) >= 13
                const uint32_t nm = ((q >> 7) & 0x01010101u) * 255u, pm = pass * 255u;
                x = (x & ~nm) | (0x20202020u & nm);               // CODE_N for non-ACGT bases
                x = (x & pm) | (0x40404040u & ~pm);               // CODE_SKIP below the quality threshold
                const uint32_t pr0 = rd.x + g * 4u;               // tile-relative position of the group's first base
                const uint32_t t0 = pr0 - rd.y;
                const uint32_t o0 = pr0 * 4u;
                if (t0 < rd.z && t0 + 3u < rd.z) {              // (t0 may have wrapped below zero: test both ends)
                    // all four bases lie on the tile inside the read's first segment (the common case):
                    // no clamping, no per-base validity test
                    #pragma unroll
                    for (uint32_t k = 0; k < 4; ++k) {
                        const uint32_t ak = cnt_base + o0 + 4u * k;
                        red_shared_add(ak ^ ((ak >> 5) & 12u), shl1_clamped32((x >> (8 * k)) & 0xffu));
                    }
                } else {
                    // a base off the tile (or outside the first segment) adds 0 to a clamped address, which
                    // is cheaper than a divergent branch around every atomic
                    #pragma unroll
                    for (uint32_t k = 0; k < 4; ++k) {
                        const uint32_t ak = cnt_base | ((o0 + 4u * k) & (uint32_t)(TILE * 4 - 1));   // wraps inside the counter array
                        uint32_t inc = shl1_clamped32((x >> (8 * k)) & 0xffu);
                        if (t0 + k >= rd.z) inc = 0;
                        red_shared_add(ak ^ ((ak >> 5) & 12u), inc);
                    }
                }
                if ((x & 0x20202020u) | (rd.w & 0xfffe0000u)) {   // rare: non-ACGT bases, reads with indels
                    if (x & 0x20202020u) {                        // non-ACGT bases count on their own plane
                        #pragma unroll
                        for (uint32_t k = 0; k < 4; ++k)
                            if (t0 + k < rd.z && ((x >> (8 * k)) & 0xffu) == CODE_N) red_shared_add(a_n + (pr0 + k) * 4u, 1u);
                    }
                    if ((rd.w >> 16) > 1u) {                      // further segments of a read with indels
                        const uint32_t s1 = (rd.w & 0xffffu) + (rd.w >> 16);
                        for (uint32_t sgi = (rd.w & 0xffffu) + 1u; sgi < s1; ++sgi) {
                            const uint4 sg = s_seg[sgi];
                            #pragma unroll
                            for (uint32_t k = 0; k < 4; ++k) {
                                const uint32_t off = d_qual + g * 4u + k - sg.z;
                                const uint32_t pr = (uint32_t)((int32_t)sg.x - p0) + off;
                                if (off < sg.y && pr < (uint32_t)TILE) {
                                    const uint32_t code = (x >> (8 * k)) & 0xffu;
                                    atomicAdd(&s_cnt[cnt_slot(pr)], shl1_clamped32(code));
                                    if (code == CODE_N) atomicAdd(&s_cntn[pr], 1u);
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();

        // ---- 6. fold this chunk's byte lanes (a position sees at most m <= 255 reads per chunk)
        #pragma unroll
        for (int k = 0; k < PILEUP_POS_PER_THREAD; ++k) {
            const uint32_t p = tid + k * PILEUP_THREADS;
            const uint32_t sl = cnt_slot(p);
            const uint32_t a8 = s_cnt[sl];
            if (a8) {
                s_cnt[sl] = 0;
                acc[k] += (uint64_t)(a8 & 0xffu) | (uint64_t)((a8 >> 8) & 0xffu) << 16 | (uint64_t)((a8 >> 16) & 0xffu) << 32 |
                          (uint64_t)(a8 >> 24) << 48;
            }
            const uint32_t cn = s_cntn[p];
            if (cn) { s_cntn[p] = 0; acc_n[k] += cn; }
        }
        c0 += m;
        // no barrier here: the next chunk only touches the counters again after two more barriers
    }

    // ---- 7. flush: 8 B + 2 B per position, fully coalesced
    #pragma unroll
    for (int k = 0; k < PILEUP_POS_PER_THREAD; ++k) {
        const size_t o = (size_t)blockIdx.x * TILE + tid + k * PILEUP_THREADS;
        acgt[o] = acc[k];
        ncnt[o] = (uint16_t)acc_n[k];
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
