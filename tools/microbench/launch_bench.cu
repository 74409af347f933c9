// launch_bench.cu -- how long does the GPU need just to run N tiny CTAs (the "skeleton" of one-CTA-per-item kernels)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 launch_bench.cu -o launch_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
extern __shared__ uint32_t smem[];
template <int MODE>
__global__ void k(const uint4* items, uint32_t* out, int nzero)
{
    if (MODE >= 1) for (int i = threadIdx.x; i < nzero; i += blockDim.x) smem[i] = 0;
    if (MODE >= 2) { uint4 it = items[blockIdx.x]; if (it.x == 0xffffffffu) out[0] = 1; }
    if (MODE >= 3) { __syncthreads(); if (smem[threadIdx.x] == 77u) out[1] = 1; }
    if (MODE >= 4) { __syncthreads(); reinterpret_cast<uint4*>(out + 16)[(size_t)blockIdx.x * 640 + threadIdx.x] = make_uint4(smem[threadIdx.x], 0, 0, 0); }
}
template <int MODE> float run(int n, int threads, int smem_bytes, const uint4* items, uint32_t* out)
{
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<n, threads, smem_bytes>>>(items, out, 1280);
    cudaEventRecord(a);
    for (int i = 0; i < 3; ++i) k<MODE><<<n, threads, smem_bytes>>>(items, out, 1280);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 3;
}
int main()
{
    const int n = 4883000;
    uint4* items; uint32_t* out;
    cudaMalloc(&items, (size_t)n * 16); cudaMemset(items, 0, (size_t)n * 16);
    cudaMalloc(&out, (size_t)n * 640 * 16 + 4096);
    for (int threads : {128, 256})
        for (int smem_kb : {8, 30, 46}) {
            printf("threads %d smem %2d KB: empty %.2f ms | +zero 5KB %.2f | +item load %.2f | +2 barriers %.2f | +10KB store %.2f\n", threads, smem_kb,
                   run<0>(n, threads, smem_kb * 1024, items, out), run<1>(n, threads, smem_kb * 1024, items, out),
                   run<2>(n, threads, smem_kb * 1024, items, out), run<3>(n, threads, smem_kb * 1024, items, out),
                   threads == 128 ? run<4>(n, threads, smem_kb * 1024, items, out) : 0.f);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
