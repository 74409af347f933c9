// msnv_gpu.cu -- implementation of the C ABI in include/msnv.h (libmsnv_gpu.so).
// Host-side orchestration only: device memory, the stream, kernel launches, result copies.
// Kernels are in kernels.cuh. There is no CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "synth_kernels.cuh"

using namespace msnv_gpu;

struct msnv_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;

    // ---- shard state
    bool open = false, has_run = false;
    uint32_t S = 0, P = 0, n_tiles = 0;
    std::vector<SampleDev> h_samples;
    struct Block { void* p; size_t bytes; bool in_slab; };
    struct Slab { uint8_t* base; size_t size, used; };
    std::vector<Block> sample_allocs;   // device blocks of the open shard
    std::vector<Block> pool;            // blocks of the previous shard, reused by the next one
    std::vector<Slab> slabs;            // small blocks are carved out of large allocations
    size_t slab_live = 0;               // carved blocks in sample_allocs + pool
    std::vector<msnv_sample_sizes> sizes;      // [S]
    uint64_t n_reads = 0, n_bases = 0, n_segs = 0;
    SampleDev* d_samples = nullptr;
    uint8_t* d_ref = nullptr;

    // ---- work buffers (grown on demand, kept across shards)
    Item* d_items = nullptr;        uint64_t cap_items = 0;
    uint8_t* d_tiles = nullptr;     // count planes, SLOT_BYTES per item
    uint8_t* d_expect = nullptr;    // expected letter per position (derived from d_ref)
    uint64_t* d_text_acgt = nullptr; uint16_t* d_text_match = nullptr; uint64_t cap_text = 0;   // classic text mode staging
    int sm_count = 0;
    uint32_t* d_tile_begin = nullptr; uint32_t* d_tile_hits = nullptr; uint8_t* d_flags = nullptr; uint64_t cap_tiles = 0;
    uint32_t* d_block_sums = nullptr; uint64_t cap_blocks = 0;
    uint2* d_range_cache = nullptr;   uint64_t cap_range = 0;
    uint32_t* d_bitmap = nullptr;     uint64_t cap_bitmap = 0;
    uint32_t* d_scalar = nullptr;   int* d_err = nullptr;
    uint32_t n_items = 0;

    // ---- hits (device + pinned host mirrors)
    uint64_t cap_hits = 0, cap_hits_S = 0;
    uint32_t *d_hit_pos = nullptr, *d_hit_total = nullptr; uint8_t *d_hit_pop = nullptr, *d_hit_ind = nullptr;
    uint16_t *d_hit_cov = nullptr, *d_hit_allele = nullptr;
    uint32_t *h_hit_pos = nullptr, *h_hit_total = nullptr; uint8_t *h_hit_pop = nullptr, *h_hit_ind = nullptr;
    uint16_t *h_hit_cov = nullptr, *h_hit_allele = nullptr;
    uint32_t* h_scalar = nullptr;   // pinned, 4 words

    cudaEvent_t ev[8] = {};
    msnv_timings tm = {};
    PileupShape tm_shape = {}; int tm_ctas = 0;   // what the last pileup launch used (MSNV_VERBOSE)
};

namespace {

int fail(msnv_ctx* c, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(ctx, MSNV_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <class T>
int grow(msnv_ctx* ctx, T*& p, uint64_t n)
{
    if (p) { CU(cudaFree(p)); p = nullptr; }
    CU(cudaMalloc((void**)&p, (size_t)(n ? n : 1) * sizeof(T)));
    return 0;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Blocks of a finished shard go to a pool: the next shard (e.g. the next genome bin of the same
// sample set) reuses them instead of paying cudaFree (a device-wide sync) and cudaMalloc per sample.
// Blocks below SLAB_BYTES / 4 are carved out of SLAB_BYTES allocations: a job with hundreds of small
// samples would otherwise make hundreds of cudaMalloc calls while the decoder threads fault pages in,
// and the two contend for the process's address-space lock (measured: 0.7 s vs 3.4 s for 400 samples).
constexpr size_t SLAB_BYTES = 256u << 20;

void release_block(msnv_ctx* ctx, const msnv_ctx::Block& b)
{
    if (!b.in_slab) { cudaFree(b.p); return; }
    if (--ctx->slab_live == 0) for (auto& sl : ctx->slabs) sl.used = 0;      // nothing carved is alive: start over
}

void free_samples(msnv_ctx* ctx)
{
    for (auto& b : ctx->pool) release_block(ctx, b);
    ctx->pool.swap(ctx->sample_allocs);
    ctx->sample_allocs.clear();
}

void* take_block(msnv_ctx* ctx, size_t bytes)
{
    size_t best = (size_t)-1, bi = 0;
    for (size_t i = 0; i < ctx->pool.size(); ++i)
        if (ctx->pool[i].bytes >= bytes && ctx->pool[i].bytes < best) { best = ctx->pool[i].bytes; bi = i; }
    if (best != (size_t)-1 && best <= bytes + bytes / 4 + (1u << 20)) {
        void* p = ctx->pool[bi].p;
        ctx->sample_allocs.push_back(ctx->pool[bi]);
        ctx->pool[bi] = ctx->pool.back(); ctx->pool.pop_back();
        return p;
    }
    if (bytes <= SLAB_BYTES / 4) {
        const size_t need = align_up(bytes, 256);
        for (int attempt = 0; attempt < 2; ++attempt) {
            for (auto& sl : ctx->slabs)
                if (sl.size - sl.used >= need) {
                    void* p = sl.base + sl.used;
                    sl.used += need;
                    ++ctx->slab_live;
                    ctx->sample_allocs.push_back({p, bytes, true});
                    return p;
                }
            void* base = nullptr;
            if (attempt || cudaMalloc(&base, SLAB_BYTES) != cudaSuccess) { cudaGetLastError(); break; }   // fall through to a plain block
            ctx->slabs.push_back({(uint8_t*)base, SLAB_BYTES, 0});
        }
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        for (auto& b : ctx->pool) release_block(ctx, b);     // give the pool back and retry once
        ctx->pool.clear();
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    ctx->sample_allocs.push_back({p, bytes, false});
    return p;
}

int ensure_hits(msnv_ctx* ctx, uint64_t n_hits)
{
    const uint64_t S = ctx->S;
    if (n_hits <= ctx->cap_hits && n_hits * S <= ctx->cap_hits_S) return 0;
    uint64_t cap = n_hits + n_hits / 4 + 1024;
    cudaFree(ctx->d_hit_pos); cudaFree(ctx->d_hit_total); cudaFree(ctx->d_hit_pop); cudaFree(ctx->d_hit_ind);
    cudaFree(ctx->d_hit_cov); cudaFree(ctx->d_hit_allele);
    cudaFreeHost(ctx->h_hit_pos); cudaFreeHost(ctx->h_hit_total); cudaFreeHost(ctx->h_hit_pop); cudaFreeHost(ctx->h_hit_ind);
    cudaFreeHost(ctx->h_hit_cov); cudaFreeHost(ctx->h_hit_allele);
    ctx->d_hit_pos = ctx->d_hit_total = nullptr; ctx->d_hit_pop = ctx->d_hit_ind = nullptr; ctx->d_hit_cov = ctx->d_hit_allele = nullptr;
    ctx->h_hit_pos = ctx->h_hit_total = nullptr; ctx->h_hit_pop = ctx->h_hit_ind = nullptr; ctx->h_hit_cov = ctx->h_hit_allele = nullptr;
    ctx->cap_hits = ctx->cap_hits_S = 0;
    CU(cudaMalloc((void**)&ctx->d_hit_pos, cap * 4));       CU(cudaMallocHost((void**)&ctx->h_hit_pos, cap * 4));
    CU(cudaMalloc((void**)&ctx->d_hit_total, cap * 20));    CU(cudaMallocHost((void**)&ctx->h_hit_total, cap * 20));
    CU(cudaMalloc((void**)&ctx->d_hit_pop, cap));           CU(cudaMallocHost((void**)&ctx->h_hit_pop, cap));
    CU(cudaMalloc((void**)&ctx->d_hit_ind, cap));           CU(cudaMallocHost((void**)&ctx->h_hit_ind, cap));
    CU(cudaMalloc((void**)&ctx->d_hit_cov, cap * S * 2));   CU(cudaMallocHost((void**)&ctx->h_hit_cov, cap * S * 2));
    CU(cudaMalloc((void**)&ctx->d_hit_allele, cap * S * 8)); CU(cudaMallocHost((void**)&ctx->h_hit_allele, cap * S * 8));
    ctx->cap_hits = cap; ctx->cap_hits_S = cap * S;
    return 0;
}

}  // namespace


// call -> ordered compaction -> per-hit gather -> copy back; shared by the BAM path and the text path.
// Records events 4..6 (the caller records event 3 once the count tiles are final).
static int run_call_phase(msnv_ctx* ctx, const msnv_call_params* prm, int text_mode, msnv_hits* hits, uint32_t& launches)
{
    cudaStream_t st = ctx->stream;
    const uint32_t S = ctx->S, n_tiles = ctx->n_tiles;
    CallParamsDev cp{prm->min_coverage, prm->calling_threshold, prm->min_fraction};
    if (cp.thr > 128)
        call_kernel<true><<<n_tiles, CALL_THREADS, 0, st>>>(ctx->d_tiles, ctx->d_items, ctx->d_tile_begin, ctx->d_ref, ctx->d_expect, cp, text_mode,
                                                            ctx->d_flags, ctx->d_tile_hits);
    else
        call_kernel<false><<<n_tiles, CALL_THREADS, 0, st>>>(ctx->d_tiles, ctx->d_items, ctx->d_tile_begin, ctx->d_ref, ctx->d_expect, cp, text_mode,
                                                             ctx->d_flags, ctx->d_tile_hits);
    ++launches;
    CU(cudaEventRecord(ctx->ev[4], st));

    // ---- ordered compaction of the called positions
    scan_kernel<<<1, 1024, 0, st>>>(ctx->d_tile_hits, n_tiles, ctx->d_scalar + 1);
    ++launches;
    CU(cudaMemcpyAsync(ctx->h_scalar + 1, ctx->d_scalar + 1, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ctx->h_scalar + 2, ctx->d_err, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (ctx->h_scalar[2]) return fail(ctx, MSNV_E_LIMIT, "a read exceeds MSNV_MAX_READ_BASES / MSNV_MAX_READ_SEGMENTS, or its offsets and segments disagree");
    const uint32_t n_hits = ctx->h_scalar[1];
    if (ensure_hits(ctx, n_hits)) return MSNV_E_CUDA;
    if (n_hits) {
        compact_kernel<<<n_tiles, TILE, 0, st>>>(ctx->d_flags, ctx->d_tile_hits, ctx->d_hit_pos, ctx->d_hit_pop, ctx->d_hit_ind);
        ++launches;
    }
    CU(cudaEventRecord(ctx->ev[5], st));

    // ---- per-hit, per-sample numbers for the formatter
    if (n_hits) {
        CU(cudaMemsetAsync(ctx->d_hit_cov, 0, (size_t)n_hits * S * 2, st));
        CU(cudaMemsetAsync(ctx->d_hit_allele, 0, (size_t)n_hits * S * 8, st));
        gather_kernel<<<n_hits, 128, 0, st>>>(ctx->d_tiles, ctx->d_items, ctx->d_tile_begin, ctx->d_ref, ctx->d_expect, ctx->d_hit_pos, S,
                                               text_mode, ctx->d_hit_cov, ctx->d_hit_allele, ctx->d_hit_total);
        ++launches;
    }
    CU(cudaEventRecord(ctx->ev[6], st));
    if (n_hits) {
        CU(cudaMemcpyAsync(ctx->h_hit_pos, ctx->d_hit_pos, (size_t)n_hits * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ctx->h_hit_total, ctx->d_hit_total, (size_t)n_hits * 20, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ctx->h_hit_pop, ctx->d_hit_pop, n_hits, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ctx->h_hit_ind, ctx->d_hit_ind, n_hits, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ctx->h_hit_cov, ctx->d_hit_cov, (size_t)n_hits * S * 2, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ctx->h_hit_allele, ctx->d_hit_allele, (size_t)n_hits * S * 8, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    hits->n_hits = n_hits;
    hits->n_samples = S;
    hits->pos = ctx->h_hit_pos; hits->pop_mask = ctx->h_hit_pop; hits->ind_mask = ctx->h_hit_ind;
    hits->cov = ctx->h_hit_cov; hits->allele = ctx->h_hit_allele; hits->total = ctx->h_hit_total;
    return MSNV_OK;
}

static int ensure_items(msnv_ctx* ctx, uint64_t n_items)
{
    if (n_items <= ctx->cap_items) return 0;
    const uint64_t cap = n_items + n_items / 8 + 64;
    if (grow(ctx, ctx->d_items, cap)) return MSNV_E_CUDA;
    cudaFree(ctx->d_tiles); ctx->d_tiles = nullptr; ctx->cap_items = 0;
    if (cudaMalloc((void**)&ctx->d_tiles, cap * SLOT_BYTES) != cudaSuccess)
        return fail(ctx, MSNV_E_NOMEM, "count tiles for %llu (sample,tile) items do not fit device memory; run smaller shards (metaSNV.py --n_splits bins the genomes)", (unsigned long long)n_items);
    ctx->cap_items = cap;
    return 0;
}

extern "C" {

int msnv_abi_version(void) { return MSNV_ABI_VERSION; }
int msnv_tile(void) { return MSNV_TILE; }

int msnv_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

// Staging limits of the pileup kernel for a shard (PileupShape) and the CTAs per SM they allow. A stage should hold
// a whole item in the common case: the limits follow the mean number of reads per item (measured by the index pass)
// with head-room for its spread; what is left of the SM's shared memory after fitting `ctas` CTAs goes into larger
// quad buffers. MSNV_MAX_READS / MSNV_CHUNK_Q4 / MSNV_PILEUP_CTAS override the choice (tuning hooks).
constexpr size_t SMEM_PER_SM = 233472, SMEM_PER_CTA_MAX = 232448, SMEM_CTA_RESERVED = 1024;
constexpr uint32_t CHUNK_Q4_CAP = 16384;

static PileupShape choose_pileup_shape(const msnv_ctx* ctx, uint64_t n_items, uint64_t item_reads, int& ctas)
{
    const double reads_per_item = (double)item_reads / (double)(n_items ? n_items : 1);
    const double q4_per_read = (double)ctx->n_bases / 4.0 / (double)(ctx->n_reads ? ctx->n_reads : 1);
    const double segs_per_read = (double)ctx->n_segs / (double)(ctx->n_reads ? ctx->n_reads : 1);
    PileupShape sh;
    uint32_t mr = (uint32_t)(reads_per_item * 1.3 + 24.0);
    if (const char* e = getenv("MSNV_MAX_READS")) mr = (uint32_t)atoi(e);
    if (mr < 16) mr = 16;
    if (mr > NARROW_MAX_READS) mr = NARROW_MAX_READS;
    sh.max_reads = mr;
    uint32_t ms = (uint32_t)(mr * segs_per_read * 1.1 + 16.0);
    if (ms < CHUNK_SEGS_MIN) ms = CHUNK_SEGS_MIN;
    sh.max_segs = up_to(ms, 8);
    const double want_reads = reads_per_item < (double)mr ? reads_per_item * 1.25 + 4.0 : (double)mr;
    uint32_t cq = (uint32_t)(want_reads * q4_per_read) + 64u;
    if (const char* e = getenv("MSNV_CHUNK_Q4")) cq = (uint32_t)atoi(e);
    if (cq < CHUNK_Q4_MIN) cq = CHUNK_Q4_MIN;
    if (cq > CHUNK_Q4_CAP) cq = CHUNK_Q4_CAP;
    sh.chunk_q4 = up_to(cq, 16);
    while (pileup_smem_layout(sh).total > SMEM_PER_CTA_MAX && sh.chunk_q4 > CHUNK_Q4_MIN) sh.chunk_q4 -= 16;
    ctas = (int)(SMEM_PER_SM / (pileup_smem_layout(sh).total + SMEM_CTA_RESERVED));
    if (ctas < 1) ctas = 1;
    if (ctas > 6) ctas = 6;
    if (const char* e = getenv("MSNV_PILEUP_CTAS")) { const int v = atoi(e); if (v >= 1 && v < ctas) ctas = v; }
    if (!getenv("MSNV_CHUNK_Q4")) {
        // spend the rest of the SM's shared memory on the quad buffers (fewer items need a second chunk)
        const size_t budget = SMEM_PER_SM / ctas - SMEM_CTA_RESERVED;
        const size_t per_q4 = 5 * PL_STAGES + 1;             // bytes per staged quad: bases + qualities per stage, one tag
        const size_t have = pileup_smem_layout(sh).total;
        if (budget > have + 256) {
            uint32_t extra = (uint32_t)((budget - have - 256) / per_q4) / 16 * 16;
            if (sh.chunk_q4 + extra > CHUNK_Q4_CAP) extra = CHUNK_Q4_CAP - sh.chunk_q4;
            sh.chunk_q4 += extra;
        }
    }
    return sh;
}

int msnv_create(int device, msnv_ctx** out)
{
    if (!out) return MSNV_E_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return MSNV_E_CUDA;
    msnv_ctx* ctx = new msnv_ctx();
    ctx->device = device;
    *out = ctx;                                   // returned even on failure so the caller can read the error
    CU(cudaSetDevice(device));
    CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    for (auto& e : ctx->ev) CU(cudaEventCreate(&e));
    CU(cudaMalloc((void**)&ctx->d_scalar, 32));
    CU(cudaMalloc((void**)&ctx->d_err, 4));
    CU(cudaMallocHost((void**)&ctx->h_scalar, 32));
    CU(cudaFuncSetAttribute(pileup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_PER_CTA_MAX));
    CU(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
    return MSNV_OK;
}

void msnv_destroy(msnv_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    free_samples(ctx);
    free_samples(ctx);                              // second call empties the pool as well
    for (auto& sl : ctx->slabs) cudaFree(sl.base);
    cudaFree(ctx->d_samples); cudaFree(ctx->d_ref);
    cudaFree(ctx->d_items); cudaFree(ctx->d_tiles); cudaFree(ctx->d_expect); cudaFree(ctx->d_text_acgt); cudaFree(ctx->d_text_match);
    cudaFree(ctx->d_tile_begin); cudaFree(ctx->d_tile_hits); cudaFree(ctx->d_flags); cudaFree(ctx->d_block_sums); cudaFree(ctx->d_range_cache); cudaFree(ctx->d_bitmap);
    cudaFree(ctx->d_scalar); cudaFree(ctx->d_err);
    cudaFree(ctx->d_hit_pos); cudaFree(ctx->d_hit_total); cudaFree(ctx->d_hit_pop); cudaFree(ctx->d_hit_ind);
    cudaFree(ctx->d_hit_cov); cudaFree(ctx->d_hit_allele);
    cudaFreeHost(ctx->h_hit_pos); cudaFreeHost(ctx->h_hit_total); cudaFreeHost(ctx->h_hit_pop); cudaFreeHost(ctx->h_hit_ind);
    cudaFreeHost(ctx->h_hit_cov); cudaFreeHost(ctx->h_hit_allele); cudaFreeHost(ctx->h_scalar);
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

void* msnv_pinned_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void msnv_pinned_free(void* p) { if (p) cudaFreeHost(p); }

const char* msnv_last_error(const msnv_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context (no CUDA device?)"; }

int msnv_shard_begin(msnv_ctx* ctx, uint32_t n_samples, uint32_t n_positions, const uint8_t* ref)
{
    if (!ctx) return MSNV_E_ARG;
    if (!ref || n_samples == 0 || n_positions == 0 || n_positions % TILE != 0)
        return fail(ctx, MSNV_E_ARG, "msnv_shard_begin: n_samples and n_positions must be positive, n_positions a multiple of %d", TILE);
    if (n_samples > 65535) return fail(ctx, MSNV_E_LIMIT, "msnv_shard_begin: at most 65535 samples (32-bit population sums)");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    free_samples(ctx);
    ctx->S = n_samples; ctx->P = n_positions; ctx->n_tiles = n_positions / TILE;
    ctx->h_samples.assign(n_samples, SampleDev{});
    ctx->sizes.assign(n_samples, msnv_sample_sizes{});
    ctx->n_reads = ctx->n_bases = ctx->n_segs = 0;
    ctx->has_run = false;
    cudaFree(ctx->d_samples); cudaFree(ctx->d_ref); cudaFree(ctx->d_expect);
    ctx->d_samples = nullptr; ctx->d_ref = nullptr; ctx->d_expect = nullptr;
    CU(cudaMalloc((void**)&ctx->d_samples, sizeof(SampleDev) * n_samples));
    CU(cudaMalloc((void**)&ctx->d_ref, n_positions));
    CU(cudaMalloc((void**)&ctx->d_expect, n_positions));
    CU(cudaMemcpyAsync(ctx->d_ref, ref, n_positions, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->n_tiles + 1 > ctx->cap_tiles) {
        if (grow(ctx, ctx->d_tile_begin, (uint64_t)ctx->n_tiles + 1)) return MSNV_E_CUDA;
        if (grow(ctx, ctx->d_tile_hits, (uint64_t)ctx->n_tiles + 1)) return MSNV_E_CUDA;
        if (grow(ctx, ctx->d_flags, (uint64_t)n_positions)) return MSNV_E_CUDA;
        ctx->cap_tiles = (uint64_t)ctx->n_tiles + 1;
    }
    ctx->open = true;
    return MSNV_OK;
}

int msnv_shard_add_sample(msnv_ctx* ctx, uint32_t sample, const msnv_sample_reads* r)
{
    if (!ctx || !r) return MSNV_E_ARG;
    if (!ctx->open) return fail(ctx, MSNV_E_STATE, "msnv_shard_add_sample: no open shard");
    if (sample >= ctx->S) return fail(ctx, MSNV_E_ARG, "msnv_shard_add_sample: sample %u out of range", sample);
    if (ctx->h_samples[sample].n_reads) return fail(ctx, MSNV_E_STATE, "msnv_shard_add_sample: sample %u added twice", sample);
    if (r->n_reads == 0) return MSNV_OK;
    if (r->max_span > 8u * MSNV_MAX_READ_BASES) return fail(ctx, MSNV_E_LIMIT, "sample %u: reference span %u exceeds the limit", sample, r->max_span);
    CU(cudaSetDevice(ctx->device));
    const size_t n = r->n_reads, n1 = n + 1;
    const size_t n_seg = r->seg_off[n], n_q4 = r->q4_off[n];
    if (n_seg < n || n_q4 < n_seg) return fail(ctx, MSNV_E_ARG, "sample %u: inconsistent offsets (%zu reads, %zu segments, %zu quads)", sample, n, n_seg, n_q4);
    // one allocation per sample, sub-arrays 256-byte aligned, 32 spare bytes behind every array
    // because the pileup kernel's bulk copies read whole 16-byte units
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes + 32, 256); return o; };
    const size_t o_pos = take(n * 4), o_sgo = take(n1 * 4), o_q4 = take(n1 * 4), o_mate = take(n * 4),
                 o_sp = take(n_seg * 4), o_sl = take(n_seg * 2), o_seq = take(n_q4), o_qual = take(n_q4 * 4);
    uint8_t* base = (uint8_t*)take_block(ctx, off);
    if (!base) return fail(ctx, MSNV_E_NOMEM, "sample %u: cannot allocate %zu bytes of device memory; run smaller shards (metaSNV.py --n_splits bins the genomes)", sample, off);
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(base + o_pos, r->pos, n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(base + o_sgo, r->seg_off, n1 * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(base + o_q4, r->q4_off, n1 * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(base + o_mate, r->mate, n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(base + o_sp, r->seg_pos, n_seg * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(base + o_sl, r->seg_len, n_seg * 2, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(base + o_seq, r->seq2, n_q4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(base + o_qual, r->qual, n_q4 * 4, cudaMemcpyHostToDevice, st));
    SampleDev& d = ctx->h_samples[sample];
    d.pos = (const int32_t*)(base + o_pos);
    d.seg_off = (const uint32_t*)(base + o_sgo); d.q4_off = (const uint32_t*)(base + o_q4);
    d.mate = (const int32_t*)(base + o_mate);
    d.seg_pos = (const int32_t*)(base + o_sp);   d.seg_len = (const uint16_t*)(base + o_sl);
    d.seq2 = base + o_seq;                      d.qual = base + o_qual;
    d.n_reads = r->n_reads; d.max_span = r->max_span ? r->max_span : 1;
    ctx->n_reads += n; ctx->n_bases += 4ull * n_q4; ctx->n_segs += n_seg;
    ctx->sizes[sample] = msnv_sample_sizes{r->n_reads, 0, d.max_span, 0, (uint64_t)n_seg, (uint64_t)n_q4};
    return MSNV_OK;
}

int msnv_shard_mask_position(msnv_ctx* ctx, uint32_t pos)
{
    if (!ctx) return MSNV_E_ARG;
    if (!ctx->open || pos >= ctx->P) return fail(ctx, MSNV_E_ARG, "msnv_shard_mask_position: position out of range");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemsetAsync(ctx->d_ref + pos, 0, 1, ctx->stream));
    return MSNV_OK;
}

int msnv_shard_sync(msnv_ctx* ctx)
{
    if (!ctx) return MSNV_E_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return MSNV_OK;
}

int msnv_shard_run(msnv_ctx* ctx, const msnv_call_params* prm, msnv_hits* hits)
{
    if (!ctx || !prm || !hits) return MSNV_E_ARG;
    if (!ctx->open) return fail(ctx, MSNV_E_STATE, "msnv_shard_run: no open shard");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t S = ctx->S, n_tiles = ctx->n_tiles;
    uint32_t launches = 0;
    memset(hits, 0, sizeof *hits);
    hits->n_samples = S;

    CU(cudaMemcpyAsync(ctx->d_samples, ctx->h_samples.data(), sizeof(SampleDev) * S, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(ctx->d_err, 0, 4, st));
    CU(cudaMemsetAsync(ctx->d_scalar, 0, 32, st));
    // expected letter per position (msnv_shard_mask_position may have changed the reference since the last run)
    expect_kernel<<<(ctx->P + 255) / 256, 256, 0, st>>>(ctx->d_ref, ctx->P, ctx->d_expect);
    ++launches;

    CU(cudaEventRecord(ctx->ev[0], st));
    // ---- index
    const uint64_t n_pairs_idx = (uint64_t)n_tiles * S;
    const uint64_t n_blocks = (n_pairs_idx + 255) / 256;
    if (n_blocks > 0x7fffffffull) return fail(ctx, MSNV_E_LIMIT, "shard too large: %u tiles x %u samples", n_tiles, S);
    if (n_blocks > ctx->cap_blocks) { if (grow(ctx, ctx->d_block_sums, n_blocks)) return MSNV_E_CUDA; ctx->cap_blocks = n_blocks; }
    // the (r_lo, r_hi) of every pair found by the counting pass is kept for the emitting pass when it fits 1 GiB
    uint2* cache = nullptr;
    if (n_pairs_idx * 8 <= (1ull << 30)) {
        if (n_pairs_idx > ctx->cap_range) { if (grow(ctx, ctx->d_range_cache, n_pairs_idx)) return MSNV_E_CUDA; ctx->cap_range = n_pairs_idx; }
        cache = ctx->d_range_cache;
    }
    {   // the index and the pileup rely on coordinate order within a sample; checked on the device (0.3 ms for 5e8 reads)
        uint64_t max_reads = 0;
        for (uint32_t s = 0; s < S; ++s) if (ctx->h_samples[s].n_reads > max_reads) max_reads = ctx->h_samples[s].n_reads;
        unsigned gx = (unsigned)((max_reads + 256 * 8 - 1) / (256 * 8)); if (gx < 1) gx = 1; if (gx > 1024) gx = 1024;
        order_check_kernel<<<dim3(gx, S), 256, 0, st>>>(ctx->d_samples, ctx->d_err);
        ++launches;
    }
    // sparse shards (fewer than ~1/3 of the pairs can be active): occupancy bitmap first
    uint32_t* bitmap = nullptr;
    const uint32_t words_per_sample = (n_tiles + 31) / 32;
    const bool sparse = getenv("MSNV_INDEX_BITMAP") ? atoi(getenv("MSNV_INDEX_BITMAP")) != 0
                                                     : ctx->n_reads / 20 < n_pairs_idx;     // < 20 reads per (tile, sample) pair on average
    if (sparse) {
        const uint64_t words = (uint64_t)words_per_sample * S;
        if (words > ctx->cap_bitmap) { if (grow(ctx, ctx->d_bitmap, words)) return MSNV_E_CUDA; ctx->cap_bitmap = words; }
        CU(cudaMemsetAsync(ctx->d_bitmap, 0, words * 4, st));
        uint64_t max_reads = 0;
        for (uint32_t s = 0; s < S; ++s) if (ctx->h_samples[s].n_reads > max_reads) max_reads = ctx->h_samples[s].n_reads;
        unsigned gx = (unsigned)((max_reads + 256 * 8 - 1) / (256 * 8)); if (gx < 1) gx = 1; if (gx > 1024) gx = 1024;
        mark_kernel<<<dim3(gx, S), 256, 0, st>>>(ctx->d_samples, words_per_sample, ctx->d_bitmap);
        ++launches;
        bitmap = ctx->d_bitmap;
    }
    unsigned long long* d_item_reads = reinterpret_cast<unsigned long long*>(ctx->d_scalar + 2);
    index_kernel<false><<<(unsigned)n_blocks, 256, 0, st>>>(ctx->d_samples, S, n_tiles, ctx->d_block_sums, nullptr, nullptr, cache, bitmap, words_per_sample,
                                                            d_item_reads);
    scan_kernel<<<1, 1024, 0, st>>>(ctx->d_block_sums, (uint32_t)n_blocks, ctx->d_scalar);
    launches += 2;
    CU(cudaMemcpyAsync(ctx->h_scalar, ctx->d_scalar, 16, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ctx->h_scalar + 4, ctx->d_err, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (ctx->h_scalar[4] == 3) return fail(ctx, MSNV_E_ARG, "the reads of a sample are not in coordinate order (pos must be ascending)");
    const uint32_t n_items = ctx->h_scalar[0];
    const uint64_t item_reads = (uint64_t)ctx->h_scalar[2] | (uint64_t)ctx->h_scalar[3] << 32;
    ctx->n_items = n_items;
    if (int rc = ensure_items(ctx, n_items)) return rc;
    index_kernel<true><<<(unsigned)n_blocks, 256, 0, st>>>(ctx->d_samples, S, n_tiles, ctx->d_block_sums, ctx->d_items, ctx->d_tile_begin, cache, bitmap,
                                                           words_per_sample, nullptr);
    ++launches;
    CU(cudaMemcpyAsync(ctx->d_tile_begin + n_tiles, ctx->d_scalar, 4, cudaMemcpyDeviceToDevice, st));
    CU(cudaEventRecord(ctx->ev[1], st));

    CU(cudaEventRecord(ctx->ev[2], st));

    // ---- pileup: persistent CTAs, as many per SM as the staging buffers allow
    if (n_items) {
        int ctas = 1;
        const PileupShape sh = choose_pileup_shape(ctx, n_items, item_reads, ctas);
        const size_t smem = pileup_smem_layout(sh).total;
        int fit = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, pileup_kernel, PL_THREADS, smem));
        if (fit < 1) return fail(ctx, MSNV_E_CUDA, "pileup kernel does not fit an SM (%zu bytes of shared memory)", smem);
        if (fit < ctas) ctas = fit;
        uint64_t grid = (uint64_t)ctas * (uint64_t)ctx->sm_count;
        if (grid > n_items) grid = n_items;
        pileup_kernel<<<(unsigned)grid, PL_THREADS, smem, st>>>(ctx->d_samples, ctx->d_items, n_items, sh, ctx->d_expect, ctx->d_tiles, ctx->d_err);
        ++launches;
        ctx->tm_shape = sh; ctx->tm_ctas = ctas;
        if (getenv("MSNV_VERBOSE"))
            fprintf(stderr, "msnv: pileup %u items (%.1f reads each), %u CTAs (%d per SM) x %d threads, stage limits %u reads / %u segments / %u quads, %zu B shared memory\n",
                    n_items, (double)item_reads / n_items, (unsigned)grid, ctas, PL_THREADS, sh.max_reads, sh.max_segs, sh.chunk_q4, smem);
    }
    CU(cudaEventRecord(ctx->ev[3], st));

    // ---- call, compaction, gather, copy back
    if (int rc = run_call_phase(ctx, prm, 0, hits, launches)) return rc;

    msnv_timings& tm = ctx->tm;
    cudaEventElapsedTime(&tm.ms_index, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&tm.ms_reserved, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&tm.ms_pileup, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&tm.ms_call, ctx->ev[3], ctx->ev[4]);
    cudaEventElapsedTime(&tm.ms_compact, ctx->ev[4], ctx->ev[5]);
    cudaEventElapsedTime(&tm.ms_gather, ctx->ev[5], ctx->ev[6]);
    cudaEventElapsedTime(&tm.ms_total, ctx->ev[0], ctx->ev[6]);
    tm.n_items = n_items; tm.n_reads = ctx->n_reads; tm.n_bases = ctx->n_bases; tm.n_tiles = n_tiles;
    tm.kernel_launches = launches;
    ctx->has_run = true;
    return MSNV_OK;
}

int msnv_call_counts(msnv_ctx* ctx, uint32_t n_samples, uint32_t n_positions, const uint8_t* ref, const uint64_t* acgt,
                     const uint16_t* matches, const msnv_call_params* prm, msnv_hits* hits)
{
    if (!ctx || !ref || !acgt || !matches || !prm || !hits) return MSNV_E_ARG;
    if (int rc = msnv_shard_begin(ctx, n_samples, n_positions, ref)) return rc;
    cudaStream_t st = ctx->stream;
    const uint64_t n_items = (uint64_t)ctx->n_tiles * n_samples;
    if (n_items > 0xffffffffull) return fail(ctx, MSNV_E_LIMIT, "msnv_call_counts: batch too large");
    if (int rc = ensure_items(ctx, n_items)) return rc;
    memset(hits, 0, sizeof *hits);
    uint32_t launches = 0;
    CU(cudaMemsetAsync(ctx->d_err, 0, 4, st));
    CU(cudaEventRecord(ctx->ev[0], st));
    if (n_items > ctx->cap_text) {
        if (grow(ctx, ctx->d_text_acgt, n_items * TILE)) return MSNV_E_CUDA;
        if (grow(ctx, ctx->d_text_match, n_items * TILE)) return MSNV_E_CUDA;
        ctx->cap_text = n_items;
    }
    CU(cudaMemcpyAsync(ctx->d_text_acgt, acgt, n_items * TILE * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_text_match, matches, n_items * TILE * 2, cudaMemcpyHostToDevice, st));
    expect_kernel<<<(ctx->P + 255) / 256, 256, 0, st>>>(ctx->d_ref, ctx->P, ctx->d_expect);
    text_tiles_kernel<<<(unsigned)((n_items * TILE + 255) / 256), 256, 0, st>>>(ctx->d_text_acgt, ctx->d_text_match, n_items * TILE, ctx->d_tiles);
    launches += 2;
    dense_items_kernel<<<(unsigned)((n_items + 1 + 255) / 256), 256, 0, st>>>(n_samples, ctx->n_tiles, ctx->d_items, ctx->d_tile_begin);
    ++launches;
    ctx->n_items = (uint32_t)n_items;
    for (int k = 1; k <= 3; ++k) CU(cudaEventRecord(ctx->ev[k], st));
    if (int rc = run_call_phase(ctx, prm, 1, hits, launches)) return rc;
    msnv_timings& tm = ctx->tm;
    tm = msnv_timings{};
    cudaEventElapsedTime(&tm.ms_call, ctx->ev[3], ctx->ev[4]);
    cudaEventElapsedTime(&tm.ms_compact, ctx->ev[4], ctx->ev[5]);
    cudaEventElapsedTime(&tm.ms_gather, ctx->ev[5], ctx->ev[6]);
    cudaEventElapsedTime(&tm.ms_total, ctx->ev[0], ctx->ev[6]);
    tm.n_items = n_items; tm.n_tiles = ctx->n_tiles; tm.kernel_launches = launches;
    ctx->has_run = true;
    return MSNV_OK;
}

int msnv_shard_synth(msnv_ctx* ctx, const msnv_synth_desc* d, int64_t* first_column)
{
    if (!ctx || !d || !d->contig_len || !d->contig_genome || !d->genome_n_sub) return MSNV_E_ARG;
    if (d->n_samples == 0 || d->n_contigs == 0 || d->read_len < 20 || d->read_len > 1000)
        return fail(ctx, MSNV_E_ARG, "msnv_shard_synth: bad description");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    msnv::synth::Model m;
    m.seed = d->seed; m.n_samples = (int32_t)d->n_samples; m.read_len = (int32_t)d->read_len; m.depth_x100 = d->depth_x100;
    m.presence_ppm = d->presence_ppm; m.paired_pct = d->paired_pct; m.site_ppm = d->site_ppm; m.err_ppm = d->err_ppm;
    m.nbase_ppm = d->nbase_ppm; m.refn_ppm = d->refn_ppm; m.indel_pct_x10 = d->indel_pct_x10; m.clip_pct_x10 = d->clip_pct_x10;
    m.mapq0_pct_x10 = d->mapq0_pct_x10;
    const uint32_t K = d->n_contigs, S = d->n_samples, L = d->read_len, q4 = (L + 3) / 4;
    std::vector<uint32_t> off(K + 1);
    uint64_t P = 0;
    for (uint32_t k = 0; k < K; ++k) {
        off[k] = (uint32_t)P;
        P += ((uint64_t)d->contig_len[k] + TILE - 1) / TILE * TILE + (d->contig_len[k] == 0 ? TILE : 0);
        if (P > 0x7ff00000ull) return fail(ctx, MSNV_E_LIMIT, "msnv_shard_synth: shard larger than 2^31 positions");
    }
    off[K] = (uint32_t)P;
    // reference: generate on the device, begin the shard around it
    std::vector<uint8_t> dummy(1, 'N');
    {
        // msnv_shard_begin wants a host reference; give it a zero page and overwrite on the device
        std::vector<uint8_t> zero((size_t)P, 0);
        if (int rc = msnv_shard_begin(ctx, S, (uint32_t)P, zero.data())) return rc;
        CU(cudaStreamSynchronize(st));
    }
    uint32_t *d_off = nullptr, *d_len = nullptr;
    CU(cudaMalloc((void**)&d_off, ((size_t)K + 1) * 4));
    CU(cudaMalloc((void**)&d_len, (size_t)K * 4));
    CU(cudaMemcpyAsync(d_off, off.data(), ((size_t)K + 1) * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_len, d->contig_len, (size_t)K * 4, cudaMemcpyHostToDevice, st));
    synth_ref_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(m, d_off, d_len, K, (uint32_t)P, ctx->d_ref);
    CU(cudaStreamSynchronize(st));
    cudaFree(d_off); cudaFree(d_len);

    int64_t first_col = -1;
    std::vector<SynthSampleCtg> blocks;
    std::vector<uint32_t> frag0;
    for (uint32_t s = 0; s < S; ++s) {
        const bool paired = msnv::synth::sample_paired(m, (int)s);
        const int32_t D = paired ? msnv::synth::sample_mate_offset(m, (int)s) : 0;
        const bool overlap = paired && D < (int32_t)L;
        const uint32_t span = msnv::synth::frag_span(m, paired, D);
        blocks.clear(); frag0.assign(1, 0);
        uint64_t n_reads = 0, n_mated = 0;
        for (uint32_t k = 0; k < K; ++k) {
            const uint32_t g = d->contig_genome[k];
            if (g >= d->n_genomes) return fail(ctx, MSNV_E_ARG, "msnv_shard_synth: contig_genome out of range");
            if (!msnv::synth::sample_has_genome(m, (int)s, (int)g) || d->contig_len[k] <= span) continue;
            const uint32_t nf = msnv::synth::n_fragments(m, d->contig_len[k], paired);
            if (!nf) continue;
            SynthSampleCtg b{k, d->contig_len[k], off[k], g, d->genome_n_sub[g], nf, (uint32_t)n_reads};
            blocks.push_back(b);
            frag0.push_back(frag0.back() + nf);
            n_reads += (uint64_t)nf * (paired ? 2 : 1);
            if (overlap) n_mated += 2ull * nf;
            if (n_reads > 0x7fffffffull) return fail(ctx, MSNV_E_LIMIT, "msnv_shard_synth: more than 2^31 reads in one sample");
        }
        if (blocks.empty()) continue;
        {   // first pileup column of this sample: first fragment of its first block
            const SynthSampleCtg& b = blocks[0];
            const int64_t c = (int64_t)b.offset + msnv::synth::frag_start(m, (int)s, b.ctg, b.len, span, b.n_frag, 0);
            if (first_col < 0 || c < first_col) first_col = c;
        }
        const size_t n = (size_t)n_reads, n1 = n + 1, nb = blocks.size(), nft = frag0.back();
        // ---- phase 1: per-read metadata (segments and quads per read, then their prefix sums)
        size_t o = 0;
        auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes + 32, 256); return r; };
        const size_t o_pos = take(n * 4), o_sgo = take(n1 * 4), o_q4 = take(n1 * 4), o_mate = take(n * 4);
        uint8_t* meta = (uint8_t*)take_block(ctx, o);
        if (!meta) return fail(ctx, MSNV_E_NOMEM, "msnv_shard_synth: out of device memory (sample %u)", s);
        SynthSampleCtg* d_blocks = nullptr; uint32_t *d_frag0 = nullptr, *d_nq = nullptr, *d_nsegs = nullptr, *d_for = nullptr;
        CU(cudaMalloc((void**)&d_blocks, nb * sizeof(SynthSampleCtg)));
        CU(cudaMalloc((void**)&d_frag0, (nb + 1) * 4));
        CU(cudaMalloc((void**)&d_nq, n * 4));
        CU(cudaMalloc((void**)&d_nsegs, n * 4));
        CU(cudaMalloc((void**)&d_for, n * 4));
        CU(cudaMemcpyAsync(d_blocks, blocks.data(), nb * sizeof(SynthSampleCtg), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_frag0, frag0.data(), (nb + 1) * 4, cudaMemcpyHostToDevice, st));
        synth_meta_kernel<<<(unsigned)((nft + 127) / 128), 128, 0, st>>>(m, (int)s, paired, D, overlap, d_blocks, d_frag0, (uint32_t)nb, (uint32_t)nft,
            (int32_t*)(meta + o_pos), d_nsegs, d_nq, (int32_t*)(meta + o_mate), d_for);
        synth_scan2_kernel<<<1, 1024, 0, st>>>(d_nsegs, d_nq, (uint32_t)n, (uint32_t*)(meta + o_sgo), (uint32_t*)(meta + o_q4));
        uint32_t n_seg = 0, n_q4_u = 0;
        CU(cudaMemcpyAsync(&n_seg, meta + o_sgo + n * 4, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(&n_q4_u, meta + o_q4 + n * 4, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        // ---- phase 2: segment records, bases, qualities
        const size_t n_q4 = n_q4_u;
        const uint32_t q_slots = q4 + 4;          // upper bound of the quads of one read (two segments)
        o = 0;
        const size_t o_sp = take((size_t)n_seg * 4), o_sl = take((size_t)n_seg * 2), o_seq = take(n_q4), o_qual = take(n_q4 * 4);
        uint8_t* data = (uint8_t*)take_block(ctx, o);
        if (!data) return fail(ctx, MSNV_E_NOMEM, "msnv_shard_synth: out of device memory (sample %u)", s);
        synth_fill_kernel<<<(unsigned)((n * q_slots + 255) / 256), 256, 0, st>>>(m, (int)s, paired, d_blocks, d_frag0, (uint32_t)nb, (uint32_t)n, q_slots,
            (const int32_t*)(meta + o_pos), (const uint32_t*)(meta + o_sgo), (const uint32_t*)(meta + o_q4), d_for,
            (int32_t*)(data + o_sp), (uint16_t*)(data + o_sl), data + o_seq, data + o_qual);
        CU(cudaStreamSynchronize(st));
        CU(cudaGetLastError());
        cudaFree(d_blocks); cudaFree(d_frag0); cudaFree(d_nq); cudaFree(d_nsegs); cudaFree(d_for);
        SampleDev& sd = ctx->h_samples[s];
        sd.pos = (const int32_t*)(meta + o_pos);
        sd.seg_off = (const uint32_t*)(meta + o_sgo); sd.q4_off = (const uint32_t*)(meta + o_q4);
        sd.mate = (const int32_t*)(meta + o_mate);
        sd.seg_pos = (const int32_t*)(data + o_sp);  sd.seg_len = (const uint16_t*)(data + o_sl);
        sd.seq2 = data + o_seq;                      sd.qual = data + o_qual;
        sd.n_reads = (uint32_t)n; sd.max_span = L + 3;
        ctx->n_reads += n; ctx->n_bases += 4ull * n_q4; ctx->n_segs += n_seg;
        ctx->sizes[s] = msnv_sample_sizes{(uint32_t)n, (uint32_t)n_mated, L + 3, 0, (uint64_t)n_seg, (uint64_t)n_q4};
    }
    if (first_column) *first_column = first_col;
    return MSNV_OK;
}

int msnv_shard_sample_sizes(msnv_ctx* ctx, uint32_t sample, msnv_sample_sizes* sizes)
{
    if (!ctx || !sizes) return MSNV_E_ARG;
    if (!ctx->open || sample >= ctx->S) return fail(ctx, MSNV_E_ARG, "msnv_shard_sample_sizes: no such sample");
    *sizes = ctx->sizes[sample];
    return MSNV_OK;
}

int msnv_shard_export_sample(msnv_ctx* ctx, uint32_t sample, int32_t* pos, uint32_t* seg_off, uint32_t* q4_off, int32_t* mate,
                             int32_t* seg_pos, uint16_t* seg_len, uint8_t* seq2, uint8_t* qual)
{
    if (!ctx) return MSNV_E_ARG;
    if (!ctx->open || sample >= ctx->S) return fail(ctx, MSNV_E_ARG, "msnv_shard_export_sample: no such sample");
    const msnv_sample_sizes z = ctx->sizes[sample];
    if (z.n_reads == 0) return MSNV_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const SampleDev& d = ctx->h_samples[sample];
    const size_t n = z.n_reads, n1 = n + 1;
    CU(cudaMemcpyAsync(pos, d.pos, n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(seg_off, d.seg_off, n1 * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(q4_off, d.q4_off, n1 * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(mate, d.mate, n * 4, cudaMemcpyDeviceToHost, st));
    if (z.n_segs) {
        CU(cudaMemcpyAsync(seg_pos, d.seg_pos, (size_t)z.n_segs * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(seg_len, d.seg_len, (size_t)z.n_segs * 2, cudaMemcpyDeviceToHost, st));
    }
    if (z.n_q4) {
        CU(cudaMemcpyAsync(seq2, d.seq2, (size_t)z.n_q4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(qual, d.qual, (size_t)z.n_q4 * 4, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    return MSNV_OK;
}

int msnv_shard_export_ref(msnv_ctx* ctx, uint8_t* ref)
{
    if (!ctx || !ref) return MSNV_E_ARG;
    if (!ctx->open) return fail(ctx, MSNV_E_STATE, "msnv_shard_export_ref: no open shard");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(ref, ctx->d_ref, ctx->P, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MSNV_OK;
}

int msnv_shard_counts(msnv_ctx* ctx, uint32_t sample, uint32_t first, uint32_t n, uint16_t* out)
{
    if (!ctx || !out) return MSNV_E_ARG;
    if (!ctx->has_run) return fail(ctx, MSNV_E_STATE, "msnv_shard_counts: run the shard first");
    if (sample >= ctx->S || (uint64_t)first + n > ctx->P) return fail(ctx, MSNV_E_ARG, "msnv_shard_counts: range out of bounds");
    if (n == 0) return MSNV_OK;
    CU(cudaSetDevice(ctx->device));
    uint16_t* d = nullptr;
    CU(cudaMalloc((void**)&d, (size_t)n * 10));
    counts_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_tiles, ctx->d_items, ctx->d_tile_begin, ctx->d_expect, sample, first, n, d);
    cudaError_t e = cudaMemcpyAsync(out, d, (size_t)n * 10, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, MSNV_E_CUDA, "msnv_shard_counts: %s", cudaGetErrorString(e));
    return MSNV_OK;
}

int msnv_get_timings(const msnv_ctx* ctx, msnv_timings* out)
{
    if (!ctx || !out) return MSNV_E_ARG;
    *out = ctx->tm;
    return MSNV_OK;
}

int msnv_cov_run(msnv_ctx* ctx, const msnv_cov_blocks* b, uint32_t max_cov, uint64_t* cov_sum, uint64_t* hist)
{
    if (!ctx || !b || !cov_sum || !hist) return MSNV_E_ARG;
    if (max_cov + 1 > (uint32_t)COV_MAX_BINS) return fail(ctx, MSNV_E_LIMIT, "msnv_cov_run: max_cov must be below %d", COV_MAX_BINS);
    const uint32_t K = b->n_contigs;
    if (K == 0) return MSNV_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t n_blocks = b->blk_off[K];
    std::vector<uint32_t> chunk0(K + 1), chunk_contig;
    uint64_t n_chunks = 0;
    for (uint32_t k = 0; k < K; ++k) {
        chunk0[k] = (uint32_t)n_chunks;
        const uint64_t c = ((uint64_t)b->contig_len[k] + COV_CHUNK - 1) / COV_CHUNK;
        n_chunks += c ? c : 1;
        if (n_chunks > 0x7fffffffull) return fail(ctx, MSNV_E_LIMIT, "msnv_cov_run: too many positions in one call");
    }
    chunk0[K] = (uint32_t)n_chunks;
    chunk_contig.resize(n_chunks);
    for (uint32_t k = 0; k < K; ++k) for (uint32_t c = chunk0[k]; c < chunk0[k + 1]; ++c) chunk_contig[c] = k;

    int32_t* d_diff = nullptr; uint32_t *d_beg = nullptr, *d_end = nullptr, *d_chunk0 = nullptr, *d_cc = nullptr, *d_len = nullptr;
    uint64_t* d_off = nullptr; unsigned long long *d_sum = nullptr, *d_hist = nullptr;
    int rc = MSNV_OK;
    auto cleanup = [&]() {
        cudaFree(d_diff); cudaFree(d_beg); cudaFree(d_end); cudaFree(d_chunk0); cudaFree(d_cc); cudaFree(d_len);
        cudaFree(d_off); cudaFree(d_sum); cudaFree(d_hist);
    };
#define CUC(call)                                                                                                   \
    do {                                                                                                            \
        cudaError_t e_ = (call);                                                                                    \
        if (e_ != cudaSuccess) { rc = fail(ctx, MSNV_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); cleanup(); return rc; } \
    } while (0)
    const size_t hist_n = (size_t)K * (max_cov + 1);
    CUC(cudaMalloc((void**)&d_diff, n_chunks * COV_CHUNK * 4));
    CUC(cudaMalloc((void**)&d_beg, (n_blocks ? n_blocks : 1) * 4));
    CUC(cudaMalloc((void**)&d_end, (n_blocks ? n_blocks : 1) * 4));
    CUC(cudaMalloc((void**)&d_chunk0, ((size_t)K + 1) * 4));
    CUC(cudaMalloc((void**)&d_cc, n_chunks * 4));
    CUC(cudaMalloc((void**)&d_len, (size_t)K * 4));
    CUC(cudaMalloc((void**)&d_off, ((size_t)K + 1) * 8));
    CUC(cudaMalloc((void**)&d_sum, (size_t)K * 8));
    CUC(cudaMalloc((void**)&d_hist, hist_n * 8));
    CUC(cudaMemsetAsync(d_diff, 0, n_chunks * COV_CHUNK * 4, st));
    CUC(cudaMemsetAsync(d_sum, 0, (size_t)K * 8, st));
    CUC(cudaMemsetAsync(d_hist, 0, hist_n * 8, st));
    if (n_blocks) {
        CUC(cudaMemcpyAsync(d_beg, b->beg, n_blocks * 4, cudaMemcpyHostToDevice, st));
        CUC(cudaMemcpyAsync(d_end, b->end, n_blocks * 4, cudaMemcpyHostToDevice, st));
    }
    CUC(cudaMemcpyAsync(d_chunk0, chunk0.data(), ((size_t)K + 1) * 4, cudaMemcpyHostToDevice, st));
    CUC(cudaMemcpyAsync(d_cc, chunk_contig.data(), n_chunks * 4, cudaMemcpyHostToDevice, st));
    CUC(cudaMemcpyAsync(d_len, b->contig_len, (size_t)K * 4, cudaMemcpyHostToDevice, st));
    CUC(cudaMemcpyAsync(d_off, b->blk_off, ((size_t)K + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_blocks)
        cov_scatter_kernel<<<(unsigned)((n_blocks + 255) / 256), 256, 0, st>>>(d_beg, d_end, d_off, d_chunk0, K, n_blocks, d_diff);
    cov_scan_kernel<<<(unsigned)n_chunks, COV_THREADS, 0, st>>>(d_diff, d_cc, d_chunk0, d_len, max_cov, d_sum, d_hist);
    CUC(cudaMemcpyAsync(cov_sum, d_sum, (size_t)K * 8, cudaMemcpyDeviceToHost, st));
    CUC(cudaMemcpyAsync(hist, d_hist, hist_n * 8, cudaMemcpyDeviceToHost, st));
    CUC(cudaStreamSynchronize(st));
    CUC(cudaGetLastError());
#undef CUC
    cleanup();
    return rc;
}

}  // extern "C"
