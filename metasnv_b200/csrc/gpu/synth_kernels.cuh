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

// sum of a u16 array (aligned bases of a sample = sum of its segment lengths)
__global__ void sum_u16_kernel(const uint16_t* __restrict__ v, uint32_t n, unsigned long long* __restrict__ out)
{
    unsigned long long acc = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc += v[i];
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
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

// One thread per fragment: ranks of its read(s), positions, segment / quad counts, mate links.
__global__ void synth_meta_kernel(Model m, int sample, bool paired, int32_t D, bool overlap, const SynthSampleCtg* __restrict__ blocks,
                                  const uint32_t* __restrict__ frag0 /*[n_blocks+1]*/, uint32_t n_blocks, uint32_t n_frag_total,
                                  int32_t* __restrict__ pos, uint32_t* __restrict__ n_segs, uint32_t* __restrict__ n_quads,
                                  int32_t* __restrict__ mate, uint32_t* __restrict__ frag_of_rank)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_frag_total) return;
    const uint32_t bi = block_of(frag0, n_blocks, g);
    const SynthSampleCtg b = blocks[bi];
    const uint32_t f = g - frag0[bi];
    const uint32_t span = msnv::synth::frag_span(m, paired, D);
    const uint32_t x = msnv::synth::frag_start(m, sample, b.ctg, b.len, span, b.n_frag, f);
    uint32_t r1 = b.read0 + f, r2 = 0;
    if (paired) {
        r1 = b.read0 + f + count_starts_below(m, sample, b, span, (int64_t)x - D, false);     // second mates strictly before
        r2 = b.read0 + f + count_starts_below(m, sample, b, span, (int64_t)x + D, true);      // first mates at or before
    }
    for (int k = 0; k < (paired ? 2 : 1); ++k) {
        const uint32_t r = k == 0 ? r1 : r2;
        const msnv::synth::ReadShape sh = msnv::synth::read_shape(m, sample, b.ctg, f, k, paired);
        const uint32_t p = b.offset + x + (k ? (uint32_t)D : 0u);
        pos[r] = (int32_t)p;
        uint32_t ns = 0, nq = 0, rx = p;
        for (int o = 0; o < sh.n_ops; ++o) {
            const uint32_t op = sh.ops[o] & 0xf, len = sh.ops[o] >> 4;
            if (op == 0) { ++ns; nq += ((rx & 3u) + len + 3u) >> 2; rx += len; }
            else if (op == 2) rx += len;
        }
        n_segs[r] = ns;
        n_quads[r] = nq;
        mate[r] = overlap ? (int32_t)(k == 1 ? r1 : r2) : -1;
        frag_of_rank[r] = (g << 1) | (uint32_t)k;
    }
}

// One thread per (read, quad slot): the quad's bases and qualities in the position-aligned layout
// of include/msnv.h; slot 0 also writes the read's segment records.
__global__ void synth_fill_kernel(Model m, int sample, bool paired, const SynthSampleCtg* __restrict__ blocks,
                                  const uint32_t* __restrict__ frag0, uint32_t n_blocks, uint32_t n_reads, uint32_t q_slots,
                                  const int32_t* __restrict__ pos, const uint32_t* __restrict__ seg_off,
                                  const uint32_t* __restrict__ q4_off, const uint32_t* __restrict__ frag_of_rank,
                                  int32_t* __restrict__ seg_pos, uint16_t* __restrict__ seg_len,
                                  uint8_t* __restrict__ seq2, uint8_t* __restrict__ qual)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)n_reads * q_slots) return;
    const uint32_t r = (uint32_t)(t / q_slots), slot = (uint32_t)(t - (uint64_t)r * q_slots);
    const uint32_t q0 = q4_off[r], nq_read = q4_off[r + 1] - q0;
    if (slot >= nq_read && slot != 0) return;
    const uint32_t fr = frag_of_rank[r];
    const uint32_t g = fr >> 1; const int k = (int)(fr & 1);
    const uint32_t bi = block_of(frag0, n_blocks, g);
    const SynthSampleCtg b = blocks[bi];
    const uint32_t f = g - frag0[bi];
    const msnv::synth::ReadShape sh = msnv::synth::read_shape(m, sample, b.ctg, f, k, paired);
    const uint64_t rid = (uint64_t)f * 2 + (uint64_t)k;
    // walk the aligned segments: rx = shard coordinate, qy = query index of the segment's first base
    uint32_t rx = (uint32_t)pos[r], qy = 0, quads_before = 0, seg = 0;
    uint32_t sbyte = 0, qword = 0;
    for (int o = 0; o < sh.n_ops; ++o) {
        const uint32_t op = sh.ops[o] & 0xf, len = sh.ops[o] >> 4;
        if (op == 0) {
            const uint32_t a = rx & 3u, nq = (a + len + 3u) >> 2;
            if (slot == 0) { seg_pos[seg_off[r] + seg] = (int32_t)rx; seg_len[seg_off[r] + seg] = (uint16_t)len; }
            if (slot >= quads_before && slot < quads_before + nq) {
                const uint32_t first = (rx - a) + 4u * (slot - quads_before);       // shard coordinate of the quad's first position
                for (uint32_t j4 = 0; j4 < 4; ++j4) {
                    const uint32_t p = first + j4;
                    if (p < rx || p >= rx + len) continue;                        // padding: quality 0, base bits 0
                    const int j = (int)(qy + (p - rx));
                    const char c = msnv::synth::read_base(m, sample, (int)b.genome, (int)b.n_sub, b.ctg, rid, j, (int64_t)p - (int64_t)b.offset);
                    uint32_t q = msnv::synth::read_qual(m, sample, b.ctg, rid, j);
                    const int code = msnv::synth::base_code(c);
                    if (code < 0) q |= 0x80u; else sbyte |= (uint32_t)code << (2 * j4);
                    qword |= q << (8 * j4);
                }
            }
            quads_before += nq; ++seg; qy += len; rx += len;
        } else if (op == 1 || op == 4) qy += len;
        else if (op == 2) rx += len;
    }
    if (slot < nq_read) {
        seq2[(size_t)q0 + slot] = (uint8_t)sbyte;
        reinterpret_cast<uint32_t*>(qual)[(size_t)q0 + slot] = qword;
    }
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
