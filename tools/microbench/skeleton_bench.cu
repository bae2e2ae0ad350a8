// skeleton_bench.cu — calibrates fixed costs of a persistent-CTA kernel: launch with 55 KB dynamic smem,
// table staging, barriers per tile.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o skeleton_bench skeleton_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(128, 4) skel(const int* __restrict__ table, int n_table, int tiles, int* sink) {
    extern __shared__ int smem[];
    const int tid = threadIdx.x;
    if (MODE >= 1) for (int a = tid; a < n_table; a += 128) smem[a] = table[a];
    __syncthreads();
    int acc = 0;
    for (int t = 0; t < tiles; ++t) {
        if (MODE >= 2) {
            // three barriers and a few dependent shared-memory round trips per tile
            smem[4096 + tid] = acc + t;
            __syncthreads();
            acc += smem[4096 + ((tid + 17) & 127)];
            smem[4300 + tid] = acc;
            __syncthreads();
            acc += smem[4300 + ((tid + 5) & 127)];
            __syncthreads();
        }
        if (MODE >= 3) {
            // 21 stores + 22 loads like the pool
#pragma unroll
            for (int k = 0; k < 21; ++k) smem[5000 + k * 144 + tid] = acc + k;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 11; ++k) acc += smem[5000 + k * 287 + (tid >> 1)];
        }
    }
    if (acc == 0x7fffffff) sink[0] = acc;
}

int main() {
    int* table; int* sink; CK(cudaMalloc(&table, 4 * 2660)); CK(cudaMalloc(&sink, 4)); CK(cudaMemset(table, 0, 4 * 2660));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int smem = 55 * 1024;
    CK(cudaFuncSetAttribute(skel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(skel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(skel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(skel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int mode = 0; mode < 4; ++mode) {
        for (int grid : {148, 566, 592}) {
            auto launch = [&]() {
                if (mode == 0) skel<0><<<grid, 128, smem>>>(table, 2660, 15, sink);
                if (mode == 1) skel<1><<<grid, 128, smem>>>(table, 2660, 15, sink);
                if (mode == 2) skel<2><<<grid, 128, smem>>>(table, 2660, 15, sink);
                if (mode == 3) skel<3><<<grid, 128, smem>>>(table, 2660, 15, sink);
            };
            for (int i = 0; i < 3; ++i) launch();
            CK(cudaEventRecord(e0));
            for (int i = 0; i < 20; ++i) launch();
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("mode %d grid %3d: %.2f us per launch\n", mode, grid, ms * 1000 / 20);
        }
    }
    return 0;
}
