// Microbenchmark: shared-memory atomic throughput on B200 (decides scatter vs gather for the pileup kernel).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 atoms_bench.cu -o atoms_bench && ./atoms_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(unsigned* out, int iters)
{
    __shared__ unsigned s[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned a = (warp * 97 + lane) & 4095, acc = 0;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) atomicAdd(&s[a], 1u << ((i & 3) * 8));              // conflict-free: consecutive words per warp
        else if (MODE == 1) atomicAdd(&s[(a & ~31u) | (lane >> 1)], 1u);     // 2 lanes per address
        else if (MODE == 2) atomicAdd(&s[(a * 33) & 4095], 1u);              // scattered, bank conflicts
        else if (MODE == 3) acc += s[a];                                     // LDS reference
        else if (MODE == 4) s[a] = i;                                        // STS reference
        else if (MODE == 5) atomicAdd(&s[a & ~31u], 1u);                     // whole warp one address
        a = (a + 32 * 7 + 1) & 4095;
    }
    __syncthreads();
    if (MODE == 3) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    else if (threadIdx.x == 0) out[blockIdx.x] = s[5];
}

template <int MODE>
void run(const char* name)
{
    unsigned* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
    const int iters = 4096, blocks = 148 * 2, threads = 1024;
    k<MODE><<<blocks, threads>>>(d, 16);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<blocks, threads>>>(d, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)blocks * threads * iters;
    printf("%-34s %8.3f ms  %8.2f Gops/s (lane ops)  %6.2f lane-ops/clk/SM @1.9GHz\n", name, ms, ops / ms / 1e6, ops / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(d);
}

int main()
{
    run<0>("atomicAdd smem conflict-free");
    run<1>("atomicAdd smem 2 lanes/address");
    run<2>("atomicAdd smem scattered");
    run<5>("atomicAdd smem warp-uniform address");
    run<3>("LDS reference");
    run<4>("STS reference");
    return 0;
}
