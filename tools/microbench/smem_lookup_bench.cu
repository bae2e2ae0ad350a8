// smem_lookup_bench.cu — clean measurement of random 4-byte shared-memory look-ups per clock per SM
// (ids generated in registers, nothing else in flight), conflict-free vs random vs staged-id patterns.
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int A = 2660, ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(128, 4) k(int* out, long long* cycles) {
    __shared__ int sTab[A + 4];
    for (int a = threadIdx.x; a < A; a += 128) sTab[a] = a;
    __syncthreads();
    unsigned s = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    int run = 0;
    const long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < ITERS; ++i) {
        s = s * 1664525u + 1013904223u;
        int id;
        if (MODE == 0) id = (int)(((s >> 9) % 83u) * 32u + (threadIdx.x & 31));   // conflict-free: lane l -> bank l
        if (MODE == 1) id = (int)((s >> 9) % (unsigned)A);                           // uniform random
        if (MODE == 2) id = (int)((s >> 9) % 83u) * 32;                              // all lanes bank 0, different rows: 32-way
        if (MODE == 3) id = (int)(((s >> 9) % 83u) * 32u + ((threadIdx.x & 15) * 2)); // 2-way conflicts exactly
        run += sTab[id];
    }
    const long long t1 = clock64();
    out[blockIdx.x * 128 + threadIdx.x] = run;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = sms * 4;
    int* out; long long* cyc; CK(cudaMalloc(&out, grid * 128 * 4)); CK(cudaMalloc(&cyc, grid * 8));
    long long* h = new long long[grid];
    const char* names[] = {"conflict-free", "uniform random", "32-way conflict", "2-way conflict"};
    for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            if (mode == 0) k<0><<<grid, 128>>>(out, cyc);
            if (mode == 1) k<1><<<grid, 128>>>(out, cyc);
            if (mode == 2) k<2><<<grid, 128>>>(out, cyc);
            if (mode == 3) k<3><<<grid, 128>>>(out, cyc);
            CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
        // per SM: 4 CTAs x 128 threads x ITERS lookups in ~avg cycles
        printf("%-16s %9.0f cycles/CTA -> %6.2f lookups/clk/SM = %5.2f cycles per warp-instruction\n", names[mode], avg,
               4.0 * 128 * ITERS / avg, avg / (4.0 * 4 * ITERS));
    }
    return 0;
}
