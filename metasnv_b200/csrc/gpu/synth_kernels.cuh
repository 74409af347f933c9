// synth_kernels.cuh -- device-side producer of synthetic shards (msnv_shard_synth). Evaluates the
// stateless model of host/synth_model.h per read, in the order the BAM writer (host/synth.cc) emits
// accepted records: by position, first mates before second mates on ties, then fragment index.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../host/synth_model.h"
#include "kernels.cuh"

namespace msnv_gpu {

using msnv::synth::Model;

struct SynthSampleCtg {          // one (sample, contig) block with reads
    uint32_t ctg;                // contig (tid) = id used by the model's hashes
    uint32_t len;
    uint32_t offset;             // shard coordinate of the contig
    uint32_t genome, n_sub;
    uint32_t n_frag;
    uint32_t read0;              // first read (rank) of the block within the sample
};

__device__ __forceinline__ uint32_t block_of(const uint32_t* __restrict__ starts, uint32_t n, uint32_t v)
{
    uint32_t lo = 0, hi = n;     // largest b with starts[b] <= v
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (starts[mid] <= v) lo = mid; else hi = mid; }
    return lo;
}

__global__ void synth_ref_kernel(Model m, const uint32_t* __restrict__ ctg_off, const uint32_t* __restrict__ ctg_len,
                                 uint32_t n_ctg, uint32_t n_pos, uint8_t* __restrict__ ref)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pos) return;
    const uint32_t c = block_of(ctg_off, n_ctg, p);
    const uint32_t q = p - ctg_off[c];
    ref[p] = q < ctg_len[c] ? (uint8_t)msnv::synth::ref_base(m, c, q) : 0;
}

// number of fragments g of the block with start(g) + add < key  (strict) or <= key
__device__ __forceinline__ uint32_t count_starts_below(const Model& m, int sample, const SynthSampleCtg& b, uint32_t span,
                                                       int64_t key, bool inclusive)
{
    uint32_t lo = 0, hi = b.n_frag;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const int64_t x = (int64_t)msnv::synth::frag_start(m, sample, b.ctg, b.len, span, b.n_frag, mid);
        const bool below = inclusive ? (x <= key) : (x < key);
        if (below) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// One thread per fragment: ranks of its read(s), positions, op / segment counts, mate links.
__global__ void synth_meta_kernel(Model m, int sample, bool paired, int32_t D, bool overlap, const SynthSampleCtg* __restrict__ blocks,
                                  const uint32_t* __restrict__ frag0 /*[n_blocks+1]*/, uint32_t n_blocks, uint32_t n_frag_total,
                                  int32_t* __restrict__ pos, uint32_t* __restrict__ n_ops, uint32_t* __restrict__ n_segs,
                                  uint32_t* __restrict__ q4_off, int32_t* __restrict__ mate, uint32_t* __restrict__ frag_of_rank)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_frag_total) return;
    const uint32_t bi = block_of(frag0, n_blocks, g);
    const SynthSampleCtg b = blocks[bi];
    const uint32_t f = g - frag0[bi];
    const uint32_t span = msnv::synth::frag_span(m, paired, D);
    const uint32_t x = msnv::synth::frag_start(m, sample, b.ctg, b.len, span, b.n_frag, f);
    const uint32_t q4 = (uint32_t)(m.read_len + 3) / 4;
    uint32_t r1 = b.read0 + f, r2 = 0;
    if (paired) {
        r1 = b.read0 + f + count_starts_below(m, sample, b, span, (int64_t)x - D, false);     // second mates strictly before
        r2 = b.read0 + f + count_starts_below(m, sample, b, span, (int64_t)x + D, true);      // first mates at or before
    }
    for (int k = 0; k < (paired ? 2 : 1); ++k) {
        const uint32_t r = k == 0 ? r1 : r2;
        const msnv::synth::ReadShape sh = msnv::synth::read_shape(m, sample, b.ctg, f, k, paired);
        pos[r] = (int32_t)(b.offset + x + (k ? (uint32_t)D : 0u));
        n_ops[r] = (uint32_t)sh.n_ops;
        uint32_t ns = 0;
        for (int o = 0; o < sh.n_ops; ++o) ns += (sh.ops[o] & 0xf) == 0;
        n_segs[r] = ns;
        q4_off[r] = r * q4;
        mate[r] = overlap ? (int32_t)(k == 1 ? r1 : r2) : -1;
        frag_of_rank[r] = (g << 1) | (uint32_t)k;
    }
}

__global__ void synth_tail_kernel(uint32_t n_reads, uint32_t q4, uint32_t* __restrict__ q4_off)
{
    q4_off[n_reads] = n_reads * q4;
}

// One thread per (read, 4-base group): bases, qualities and (group 0) the CIGAR words.
__global__ void synth_fill_kernel(Model m, int sample, bool paired, const SynthSampleCtg* __restrict__ blocks,
                                  const uint32_t* __restrict__ frag0, uint32_t n_blocks, uint32_t n_reads,
                                  const int32_t* __restrict__ pos, const uint32_t* __restrict__ cig_off,
                                  const uint32_t* __restrict__ frag_of_rank, uint32_t* __restrict__ cigar,
                                  uint8_t* __restrict__ seq2, uint8_t* __restrict__ qual)
{
    const uint32_t q4 = (uint32_t)(m.read_len + 3) / 4;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)n_reads * q4) return;
    const uint32_t r = (uint32_t)(t / q4), grp = (uint32_t)(t - (uint64_t)r * q4);
    const uint32_t fr = frag_of_rank[r];
    const uint32_t g = fr >> 1; const int k = (int)(fr & 1);
    const uint32_t bi = block_of(frag0, n_blocks, g);
    const SynthSampleCtg b = blocks[bi];
    const uint32_t f = g - frag0[bi];
    const msnv::synth::ReadShape sh = msnv::synth::read_shape(m, sample, b.ctg, f, k, paired);
    const uint64_t rid = (uint64_t)f * 2 + (uint64_t)k;
    const int64_t rpos = (int64_t)pos[r] - (int64_t)b.offset;
    if (grp == 0) for (int o = 0; o < sh.n_ops; ++o) cigar[cig_off[r] + o] = sh.ops[o];
    uint32_t sbyte = 0, qword = 0;
    for (int j4 = 0; j4 < 4; ++j4) {
        const int j = (int)grp * 4 + j4;
        if (j >= m.read_len) break;
        // reference position of query base j (or -1 inside an insertion / clip)
        int64_t refp = -1; int qy = 0; int64_t rx = rpos;
        for (int o = 0; o < sh.n_ops; ++o) {
            const int op = (int)(sh.ops[o] & 0xf), len = (int)(sh.ops[o] >> 4);
            if (op == 0) { if (j < qy + len) { refp = rx + (j - qy); break; } qy += len; rx += len; }
            else if (op == 1 || op == 4) { if (j < qy + len) { refp = -1; break; } qy += len; }
            else if (op == 2) rx += len;
        }
        const char c = msnv::synth::read_base(m, sample, (int)b.genome, (int)b.n_sub, b.ctg, rid, j, refp);
        uint32_t q = msnv::synth::read_qual(m, sample, b.ctg, rid, j);
        const int code = msnv::synth::base_code(c);
        if (code < 0) q |= 0x80u; else sbyte |= (uint32_t)code << (2 * j4);
        qword |= q << (8 * j4);
    }
    seq2[(size_t)r * q4 + grp] = (uint8_t)sbyte;
    reinterpret_cast<uint32_t*>(qual)[(size_t)r * q4 + grp] = qword;
}

// exclusive scans of two u32 arrays of one sample by one CTA (n up to 2^31); writes n+1 entries
__global__ void __launch_bounds__(1024) synth_scan2_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t n,
                                                           uint32_t* __restrict__ a_out, uint32_t* __restrict__ b_out)
{
    __shared__ uint32_t s_wa[32], s_wb[32];
    __shared__ uint32_t s_ca, s_cb;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_ca = 0; s_cb = 0; }
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t xa = i < n ? a[i] : 0, xb = i < n ? b[i] : 0;
        uint32_t ia = xa, ib = xb;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t oa = __shfl_up_sync(0xffffffffu, ia, d), ob = __shfl_up_sync(0xffffffffu, ib, d);
            if (lane >= d) { ia += oa; ib += ob; }
        }
        if (lane == 31) { s_wa[warp] = ia; s_wb[warp] = ib; }
        __syncthreads();
        if (warp == 0) {
            uint32_t wa = s_wa[lane], wb = s_wb[lane], ja = wa, jb = wb;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t oa = __shfl_up_sync(0xffffffffu, ja, d), ob = __shfl_up_sync(0xffffffffu, jb, d);
                if (lane >= d) { ja += oa; jb += ob; }
            }
            s_wa[lane] = ja - wa; s_wb[lane] = jb - wb;
        }
        __syncthreads();
        const uint32_t ca = s_ca, cb = s_cb;
        if (i < n) { a_out[i] = ca + s_wa[warp] + ia - xa; b_out[i] = cb + s_wb[warp] + ib - xb; }
        __syncthreads();
        if (threadIdx.x == 1023) { s_ca = ca + s_wa[warp] + ia; s_cb = cb + s_wb[warp] + ib; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { a_out[n] = s_ca; b_out[n] = s_cb; }
}

}  // namespace msnv_gpu
