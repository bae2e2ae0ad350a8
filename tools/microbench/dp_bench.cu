// dp_bench.cu — stand-alone micro-benchmark of the windowed forward-backward stage (W = 20).
// Compares a scalar one-window-per-thread recursion with packed f32x2 two-windows-per-thread
// variants.  All windows are treated as valid (one endless contig, indices clamped at the ends),
// so only the arithmetic / shared-memory structure is measured.  Build + run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o dp_bench dp_bench.cu && ./dp_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int W = 20;
__constant__ float c_m01, c_m10, c_m11;

__device__ __forceinline__ float rcpf(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// ---------------------------------------------------------------- V0: scalar, 1 window / thread
template <int NT>
__global__ void __launch_bounds__(NT) dp_v0(const float* __restrict__ u, float* __restrict__ out, int G, int tiles_per_cta, int num_tiles) {
    constexpr int TOUT = NT - W;           // outputs per tile
    constexpr int NG = NT + W - 1;         // genes staged
    __shared__ float sU[NG + 1];
    __shared__ float pool[W * NT];
    const int tid = threadIdx.x;
    const float m01 = c_m01, m10 = c_m10, m11 = c_m11;
    for (int tile = blockIdx.x * tiles_per_cta; tile < min(num_tiles, (blockIdx.x + 1) * tiles_per_cta); ++tile) {
        const int T0 = tile * TOUT;
        const int Gs = T0 - (W - 1);
        for (int j = tid; j < NG; j += NT) sU[j] = u[min(max(Gs + j, 0), G - 1)];
        __syncthreads();
        float ra[W];
        float r = sU[tid];
        ra[0] = r;
#pragma unroll
        for (int k = 1; k < W; ++k) {
            const float uu = sU[tid + k];
            const float num = fmaf(r, m11, m01), den = fmaf(r, m10, 1.0f);
            r = num * uu * rcpf(den);
            ra[k] = r;
        }
        float s = 1.0f;
        pool[(W - 1) * NT + tid] = ra[W - 1];
#pragma unroll
        for (int k = W - 2; k >= 0; --k) {
            const float w = sU[tid + k + 1] * s;
            s = fmaf(w, m11, m10) * rcpf(fmaf(w, m01, 1.0f));
            pool[k * NT + tid] = ra[k] * s;
        }
        __syncthreads();
        if (tid >= W - 1 && tid < W - 1 + TOUT && T0 + tid - (W - 1) < G) {
            float q = 0.f;
#pragma unroll
            for (int k = 0; k < W; ++k) q = fmaxf(q, pool[k * NT + tid - k]);
            out[T0 + tid - (W - 1)] = q * rcpf(1.0f + q);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- V1: packed f32x2, windows (2t, 2t+1) per thread
// KEEP = keep the 20 unary pairs in registers for the backward pass instead of re-reading shared memory.
template <int NTH, bool KEEP, int MINB>
__global__ void __launch_bounds__(NTH, MINB) dp_v1(const float* __restrict__ u, float* __restrict__ out, int G, int tiles_per_cta, int num_tiles) {
    constexpr int NW = 2 * NTH;            // windows per tile
    constexpr int TOUT = NW - W;           // outputs per tile (even)
    constexpr int NG = NW + W;             // genes staged (even, >= NW + W - 1)
    constexpr int P = NTH + 16;            // pool pitch: odd/even genes land 16 banks apart
    __shared__ __align__(16) float sU0[NG + 2];   // sU0[j] = u[Gs + j]
    __shared__ __align__(16) float sU1[NG + 2];   // sU1[j] = u[Gs + j + 1]
    __shared__ float pool[(W + 1) * P];
    const int tid = threadIdx.x;
    const float2 M01 = make_float2(c_m01, c_m01), M10 = make_float2(c_m10, c_m10), M11 = make_float2(c_m11, c_m11);
    const float2 ONE = make_float2(1.f, 1.f);
    for (int tile = blockIdx.x * tiles_per_cta; tile < min(num_tiles, (blockIdx.x + 1) * tiles_per_cta); ++tile) {
        const int T0 = tile * TOUT;
        const int Gs = T0 - W;             // local gene j <-> global Gs + j ; outputs are local [W, W + TOUT)
        for (int j = tid; j < NG + 1; j += NTH) {
            const float v = u[min(max(Gs + j, 0), G - 1)];
            if (j < NG) sU0[j] = v;
            if (j >= 1) sU1[j - 1] = v;
        }
        __syncthreads();
        const int b0 = 2 * tid;            // first window of this thread starts at local gene b0
        auto upair = [&](int k) -> float2 {  // (u[b0+k], u[b0+k+1]) with an aligned 8-byte load
            return (k & 1) ? *reinterpret_cast<const float2*>(&sU1[b0 + k - 1]) : *reinterpret_cast<const float2*>(&sU0[b0 + k]);
        };
        float2 ra[W];
        float2 uk[KEEP ? W : 1];
        float2 R = upair(0);
        if (KEEP) uk[0] = R;
        ra[0] = R;
#pragma unroll
        for (int k = 1; k < W; ++k) {
            const float2 U = upair(k);
            if (KEEP) uk[k] = U;
            const float2 num = __ffma2_rn(R, M11, M01);
            const float2 den = __ffma2_rn(R, M10, ONE);
            const float2 inv = make_float2(rcpf(den.x), rcpf(den.y));
            R = __fmul2_rn(__fmul2_rn(num, U), inv);
            ra[k] = R;
        }
        // backward + odds; m[j] = best odds for local gene b0 + j over this thread's two windows
        float2 S = ONE;
        float2 Qprev = ra[W - 1];          // Q[19]
        pool[W * P + tid] = Qprev.y;       // m[20] = q_b[19]
#pragma unroll
        for (int k = W - 2; k >= 0; --k) {
            const float2 U = KEEP ? uk[k + 1] : upair(k + 1);
            const float2 Wv = __fmul2_rn(U, S);
            const float2 num = __ffma2_rn(Wv, M11, M10);
            const float2 den = __ffma2_rn(Wv, M01, ONE);
            const float2 inv = make_float2(rcpf(den.x), rcpf(den.y));
            S = __fmul2_rn(num, inv);
            const float2 Q = __fmul2_rn(ra[k], S);
            pool[(k + 1) * P + tid] = fmaxf(Qprev.x, Q.y);   // m[k+1] = max(q_a[k+1], q_b[k])
            Qprev = Q;
        }
        pool[tid] = Qprev.x;               // m[0] = q_a[0]
        __syncthreads();
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
            const int g = W + tid + rep * NTH;         // local gene
            if (g < W + TOUT && Gs + g < G) {
                const int par = g & 1;
                float q = 0.f;
#pragma unroll
                for (int i = 0; i <= W / 2; ++i) {
                    const int j = par + 2 * i;
                    if (j <= W) q = fmaxf(q, pool[j * P + ((g - j) >> 1)]);
                }
                out[Gs + g] = q * rcpf(1.0f + q);
            }
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------- V3: pair form (no division in the recursions), lanes of a
// float2 = the two labels of ONE window.  alpha_k = (alpha_k[other], alpha_k[pos]) un-normalised, E_k = (e0, e1) with max 1.
//   fwd  A' = (A o (1, m11) + swap(A) o (m10, m01)) o E_k
//   bwd  X = B o E_{k+1};  B' = X o (1, m11) + swap(X) o (m01, m10)
//   p_k  = a1 b1 / (a0 b0 + a1 b1)                                 (one MUFU per position instead of two)
// Power-of-two renormalisation every RN steps keeps the pairs inside the FP32 range.
__device__ __forceinline__ float2 renorm(float2 A) {
    const float mx = fmaxf(A.x, A.y);
    const float sc = __int_as_float(0x7f000000 - (__float_as_int(mx) & 0x7f800000));
    return __fmul2_rn(A, make_float2(sc, sc));
}

template <int NT, int MINB, int RN>
__global__ void __launch_bounds__(NT, MINB) dp_v3(const float* __restrict__ u, float* __restrict__ out, int G, int tiles_per_cta, int num_tiles) {
    constexpr int TOUT = NT - W;
    constexpr int NG = NT + W - 1;
    __shared__ __align__(16) float2 sE[NG + 1];
    __shared__ float pool[W * NT];
    const int tid = threadIdx.x;
    const float2 Md = make_float2(1.0f, c_m11), Mxf = make_float2(c_m10, c_m01), Mxb = make_float2(c_m01, c_m10);
    for (int tile = blockIdx.x * tiles_per_cta; tile < min(num_tiles, (blockIdx.x + 1) * tiles_per_cta); ++tile) {
        const int T0 = tile * TOUT;
        const int Gs = T0 - (W - 1);
        for (int j = tid; j < NG; j += NT) {
            const float uu = u[min(max(Gs + j, 0), G - 1)];
            sE[j] = uu > 1.0f ? make_float2(rcpf(uu), 1.0f) : make_float2(1.0f, uu);
        }
        __syncthreads();
        float2 al[W];
        float2 A = sE[tid];
        al[0] = A;
#pragma unroll
        for (int k = 1; k < W; ++k) {
            const float2 E = sE[tid + k];
            const float2 T = __ffma2_rn(make_float2(A.y, A.x), Mxf, __fmul2_rn(A, Md));
            A = __fmul2_rn(T, E);
            if (k % RN == 0) A = renorm(A);
            al[k] = A;
        }
        float2 B = make_float2(1.0f, 1.0f);
#pragma unroll
        for (int k = W - 1; k >= 0; --k) {
            const float2 N = __fmul2_rn(al[k], B);
            pool[k * NT + tid] = N.y * rcpf(N.x + N.y);
            if (k > 0) {
                const float2 X = __fmul2_rn(B, sE[tid + k]);
                B = __ffma2_rn(make_float2(X.y, X.x), Mxb, __fmul2_rn(X, Md));
                if (k % RN == 0) B = renorm(B);
            }
        }
        __syncthreads();
        if (tid >= W - 1 && tid < W - 1 + TOUT && T0 + tid - (W - 1) < G) {
            float p = 0.f;
#pragma unroll
            for (int k = 0; k < W; ++k) p = fmaxf(p, pool[k * NT + tid - k]);
            out[T0 + tid - (W - 1)] = p;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- host
template <typename K>
float time_kernel(K launch, int iters) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGetLastError());
    return ms / iters;
}

int main() {
    const int G = 2000810;
    std::vector<float> hu(G);
    unsigned long long s = 88172645463325252ULL;
    for (int i = 0; i < G; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const double x = ((s >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;  // uniform(-1,1)
        hu[i] = (float)exp(8.0 * x * x * x);                                    // heavy-ish tails
    }
    float *du, *dout; CK(cudaMalloc(&du, G * 4)); CK(cudaMalloc(&dout, G * 4));
    CK(cudaMemcpy(du, hu.data(), G * 4, cudaMemcpyHostToDevice));
    const float m01 = expf(-2.599571900486168f - 2.669891070463728f), m10 = expf(-2.6019205422130995f - 2.669891070463728f),
                m11 = expf(2.5683226020688488f - 2.669891070463728f);
    CK(cudaMemcpyToSymbol(c_m01, &m01, 4)); CK(cudaMemcpyToSymbol(c_m10, &m10, 4)); CK(cudaMemcpyToSymbol(c_m11, &m11, 4));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    std::vector<float> ref(G), got(G);

    auto report = [&](const char* name, float ms, bool is_ref) {
        CK(cudaMemcpy(got.data(), dout, G * 4, cudaMemcpyDeviceToHost));
        double maxd = 0, sum = 0;
        for (int i = 64; i < G - 64; ++i) { sum += got[i]; if (!is_ref) maxd = fmax(maxd, fabs((double)got[i] - ref[i])); }
        if (is_ref) ref = got;
        printf("%-28s %8.2f us  %7.2f Gwindows/s  checksum %.6f  max|d| vs v0 %.2e\n", name, ms * 1e3, G / ms / 1e6, sum, maxd);
        CK(cudaMemset(dout, 0, G * 4));
    };

#define RUN(NAME, KERNEL, NTHREADS, TOUT_, IS_REF) do { \
        int per_sm = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, KERNEL, NTHREADS, 0)); \
        const int num_tiles = (G + (TOUT_) - 1) / (TOUT_); \
        int grid = sms * per_sm; if (grid > num_tiles) grid = num_tiles; \
        const int tpc = (num_tiles + grid - 1) / grid; \
        float ms = time_kernel([&] { KERNEL<<<grid, NTHREADS>>>(du, dout, G, tpc, num_tiles); }, 20); \
        char nm[96]; snprintf(nm, sizeof nm, "%s (occ %d/SM)", NAME, per_sm); report(nm, ms, IS_REF); } while (0)

    RUN("v0 scalar NT=256", (dp_v0<256>), 256, 256 - W, true);
    RUN("v0 scalar NT=128", (dp_v0<128>), 128, 128 - W, false);
    RUN("v0 scalar NT=512", (dp_v0<512>), 512, 512 - W, false);
    RUN("v1 packed NTH=128 reload mb4", (dp_v1<128, false, 4>), 128, 256 - W, false);
    RUN("v1 packed NTH=128 reload mb6", (dp_v1<128, false, 6>), 128, 256 - W, false);
    RUN("v1 packed NTH=128 reload mb8", (dp_v1<128, false, 8>), 128, 256 - W, false);
    RUN("v1 packed NTH=128 keep mb4", (dp_v1<128, true, 4>), 128, 256 - W, false);
    RUN("v1 packed NTH=128 keep mb5", (dp_v1<128, true, 5>), 128, 256 - W, false);
    RUN("v1 packed NTH=256 reload mb2", (dp_v1<256, false, 2>), 256, 512 - W, false);
    RUN("v1 packed NTH=256 reload mb3", (dp_v1<256, false, 3>), 256, 512 - W, false);
    RUN("v1 packed NTH=256 reload mb4", (dp_v1<256, false, 4>), 256, 512 - W, false);
    RUN("v1 packed NTH=256 keep mb2", (dp_v1<256, true, 2>), 256, 512 - W, false);
    RUN("v3 pair NT=256 mb4 rn5", (dp_v3<256, 4, 5>), 256, 256 - W, false);
    RUN("v3 pair NT=256 mb3 rn5", (dp_v3<256, 3, 5>), 256, 256 - W, false);
    RUN("v3 pair NT=256 mb2 rn5", (dp_v3<256, 2, 5>), 256, 256 - W, false);
    RUN("v3 pair NT=256 mb4 rn4", (dp_v3<256, 4, 4>), 256, 256 - W, false);
    RUN("v3 pair NT=256 mb4 rn99", (dp_v3<256, 4, 99>), 256, 256 - W, false);
    RUN("v3 pair NT=128 mb8 rn5", (dp_v3<128, 8, 5>), 128, 128 - W, false);
    RUN("v3 pair NT=512 mb2 rn5", (dp_v3<512, 2, 5>), 512, 512 - W, false);
    RUN("v1 packed NTH=64 reload mb12", (dp_v1<64, false, 12>), 64, 128 - W, false);
    return 0;
}
