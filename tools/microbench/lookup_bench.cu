// lookup_bench.cu — throughput of random 4-byte table look-ups (2,660-entry table) per SM:
// shared memory vs texture fetch vs cached global load vs mixes.  One CTA of 256 threads x 4 per SM.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int A = 2660;
constexpr int PER_THREAD = 52 * 16;

template <int MODE>
__global__ void __launch_bounds__(256) bench(const int* __restrict__ ids, const int* __restrict__ table, cudaTextureObject_t tex, int* out) {
    __shared__ int sTab[A];
    for (int a = threadIdx.x; a < A; a += 256) sTab[a] = table[a];
    __syncthreads();
    const int4* v = reinterpret_cast<const int4*>(ids) + (size_t)(blockIdx.x * 256 + threadIdx.x) * (PER_THREAD / 4);
    int run = 0;
#pragma unroll 4
    for (int i = 0; i < PER_THREAD / 4; ++i) {
        const int4 id = v[i];
        if (MODE == 0) { run += sTab[id.x]; run += sTab[id.y]; run += sTab[id.z]; run += sTab[id.w]; }
        if (MODE == 1) { run += tex1Dfetch<int>(tex, id.x); run += tex1Dfetch<int>(tex, id.y); run += tex1Dfetch<int>(tex, id.z); run += tex1Dfetch<int>(tex, id.w); }
        if (MODE == 2) { run += __ldg(table + id.x); run += __ldg(table + id.y); run += __ldg(table + id.z); run += __ldg(table + id.w); }
        if (MODE == 3) { run += sTab[id.x]; run += tex1Dfetch<int>(tex, id.y); run += sTab[id.z]; run += tex1Dfetch<int>(tex, id.w); }
        if (MODE == 4) { run += sTab[id.x]; run += __ldg(table + id.y); run += sTab[id.z]; run += __ldg(table + id.w); }
        if (MODE == 5) { run += sTab[id.x]; run += sTab[id.y]; run += sTab[id.z]; run += __ldg(table + id.w); }
        if (MODE == 6) { run += id.x + id.y + id.z + id.w; }
    }
    out[blockIdx.x * 256 + threadIdx.x] = run;
}

int main() {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = sms * 4;
    const size_t n = (size_t)grid * 256 * PER_THREAD;
    std::vector<int> h(n);
    unsigned s = 12345;
    for (size_t i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) % A; }
    std::vector<int> ht(A); for (int a = 0; a < A; ++a) ht[a] = a * 7 + 1;
    int *ids, *table, *out; CK(cudaMalloc(&ids, n * 4)); CK(cudaMalloc(&table, A * 4)); CK(cudaMalloc(&out, grid * 256 * 4));
    CK(cudaMemcpy(ids, h.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(table, ht.data(), A * 4, cudaMemcpyHostToDevice));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = table; rd.res.linear.desc = cudaCreateChannelDesc<int>(); rd.res.linear.sizeInBytes = A * 4;
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const char* names[] = {"smem", "tex", "ldg", "smem+tex 1:1", "smem+ldg 1:1", "smem+ldg 3:1", "no lookup (ids only)"};
    for (int mode = 0; mode < 7; ++mode) {
        auto launch = [&]() {
            switch (mode) {
                case 0: bench<0><<<grid, 256>>>(ids, table, tex, out); break;
                case 1: bench<1><<<grid, 256>>>(ids, table, tex, out); break;
                case 2: bench<2><<<grid, 256>>>(ids, table, tex, out); break;
                case 3: bench<3><<<grid, 256>>>(ids, table, tex, out); break;
                case 4: bench<4><<<grid, 256>>>(ids, table, tex, out); break;
                case 5: bench<5><<<grid, 256>>>(ids, table, tex, out); break;
                case 6: bench<6><<<grid, 256>>>(ids, table, tex, out); break;
            }
        };
        launch(); launch();
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 5; ++i) launch();
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
        printf("%-22s %8.1f us  %6.2f lookups/clk/SM (at 1.95 GHz)  %.0f GB/s of ids\n", names[mode], ms * 1e3,
               (double)n / (ms * 1e-3) / sms / 1.95e9, n * 4.0 / (ms * 1e-3) / 1e9);
    }
    return 0;
}
