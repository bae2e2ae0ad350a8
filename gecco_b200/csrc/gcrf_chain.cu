// gcrf_chain.cu — whole-contig (un-windowed) marginals: the primitive equal to
// sklearn_crfsuite.CRF.predict_marginals_single on a full row (call site gecco/crf/__init__.py:253),
// and the deep-chain case of BASELINE config 5.
//
// A chain of n items is not a serial problem: with odds-ratio messages both recursions are Moebius
// maps, i.e. 2x2 matrices acting on (x; 1),
//     forward   r_k     = u_k (m01 + m11 r_{k-1}) / (1 + m10 r_{k-1})   F_k = [[u_k m11, u_k m01], [m10, 1]]
//     backward  s_{k-1} = (m10 + m11 u_k s_k) / (1 + m01 u_k s_k)       B_k = [[m11 u_k, m10], [m01 u_k, 1]]
// so prefix/suffix products give every message.  One warp owns a contig: each lane composes the
// matrices of its contiguous segment, a shuffle scan combines the 32 segment products (each product
// renormalised by its largest entry), and the lane then replays its segment from the exact incoming
// message.  f64 arithmetic: this path is the numerically delicate one (thousands of steps) and is
// not the HBM-bound headline path.
#include "gcrf_kernels.cuh"

#include <cstdlib>

namespace gcrf {

namespace {

constexpr int kThreads = 256;
constexpr double kClamp64 = 300.0;

__device__ __forceinline__ int64_t load_gene_ptr(const CsrDev &csr, int64_t g) {
    return csr.gene_ptr64 ? __ldg(csr.gene_ptr64 + g) : (int64_t)__ldg(csr.gene_ptr32 + g);
}

// u_g = exp(clamp(sum_a delta_a)) in f64, one thread per gene.
__global__ void __launch_bounds__(kThreads) unary64_kernel(const CsrDev csr, const double *__restrict__ table64,
                                                          int32_t A, double *__restrict__ u) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < csr.G; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = load_gene_ptr(csr, g), e = load_gene_ptr(csr, g + 1);
        double acc = 0.0;
        for (int64_t p = s; p < e; ++p) {
            const uint32_t id = (uint32_t)__ldg(csr.attr_idx + p);
            acc += __ldg(table64 + (id < (uint32_t)A ? id : (uint32_t)A));
        }
        u[g] = exp(fmin(fmax(acc, -kClamp64), kClamp64));
    }
}

struct M2 {
    double a, b, c, d;  // [[a, b], [c, d]] acting on (x; 1): x -> (a x + b) / (c x + d)
};

__device__ __forceinline__ M2 identity() { return M2{1.0, 0.0, 0.0, 1.0}; }

// X after Y: (X o Y)(x) = X(Y(x)); entries are non-negative, rescale by the largest one.
__device__ __forceinline__ M2 compose(const M2 &X, const M2 &Y) {
    M2 r;
    r.a = X.a * Y.a + X.b * Y.c;
    r.b = X.a * Y.b + X.b * Y.d;
    r.c = X.c * Y.a + X.d * Y.c;
    r.d = X.c * Y.b + X.d * Y.d;
    const double mx = fmax(fmax(r.a, r.b), fmax(r.c, r.d));
    const double inv = mx > 0.0 ? 1.0 / mx : 1.0;
    r.a *= inv; r.b *= inv; r.c *= inv; r.d *= inv;
    return r;
}

__device__ __forceinline__ M2 shfl_up(const M2 &m, int d) {
    return M2{__shfl_up_sync(0xffffffffu, m.a, d), __shfl_up_sync(0xffffffffu, m.b, d),
              __shfl_up_sync(0xffffffffu, m.c, d), __shfl_up_sync(0xffffffffu, m.d, d)};
}

__device__ __forceinline__ M2 shfl_down(const M2 &m, int d) {
    return M2{__shfl_down_sync(0xffffffffu, m.a, d), __shfl_down_sync(0xffffffffu, m.b, d),
              __shfl_down_sync(0xffffffffu, m.c, d), __shfl_down_sync(0xffffffffu, m.d, d)};
}

// One warp per contig (grid-stride).  r_buf[G] holds the forward odds between the two passes.
__global__ void __launch_bounds__(kThreads)
chain_kernel(const CsrDev csr, const double *__restrict__ u, double *__restrict__ r_buf, void *__restrict__ out,
             int out_f32, double m01, double m10, double m11) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp0; c < csr.C; c += nwarps) {
        const int64_t g0 = __ldg(csr.contig_ptr + c), g1 = __ldg(csr.contig_ptr + c + 1);
        const int64_t n = g1 - g0;
        if (n <= 0) continue;
        const int64_t seg = (n + 31) / 32;
        int64_t b = g0 + (int64_t)lane * seg, e = b + seg;
        if (b > g1) b = g1;
        if (e > g1) e = g1;

        // ---- forward: segment products, inclusive scan, replay
        M2 S = identity();
        for (int64_t g = b; g < e; ++g) {
            const double ug = u[g];
            const M2 F = (g == g0) ? M2{0.0, ug, 0.0, 1.0} : M2{ug * m11, ug * m01, m10, 1.0};
            S = compose(F, S);
        }
        M2 P = S;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const M2 Y = shfl_up(P, d);
            if (lane >= d) P = compose(P, Y);
        }
        M2 X = shfl_up(P, 1);  // product of all earlier segments
        double r = 0.0;        // lane 0 starts at g0, whose map ignores its input
        if (lane > 0) r = X.b / X.d;  // first column of X is zero: every chain starts with F_0
        for (int64_t g = b; g < e; ++g) {
            const double ug = u[g];
            r = (g == g0) ? ug : ((m01 + m11 * r) / (1.0 + m10 * r)) * ug;
            r_buf[g] = r;
        }

        // ---- backward: segment products over B_g (g = b+1 .. e maps s_e... see header), suffix scan
        // R_lane maps s at the last item of the lane's segment back to s at the last item of the
        // previous lane's segment: R = B_b o B_{b+1} o ... o B_{e-1}, B_g: s_g -> s_{g-1}.
        M2 R = identity();
        for (int64_t g = b; g < e; ++g) {
            if (g == g0) continue;  // B_{g0} would map to the item before the chain
            const double ug = u[g];
            const M2 B = M2{m11 * ug, m10, m01 * ug, 1.0};
            R = compose(R, B);
        }
        M2 Q = R;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const M2 Y = shfl_down(Q, d);
            if (lane + d < 32) Q = compose(Q, Y);
        }
        M2 E = shfl_down(Q, 1);  // product of all later segments
        if (lane == 31) E = identity();
        double s = (E.a + E.b) / (E.c + E.d);  // applied to s_{n-1} = 1
        for (int64_t g = e - 1; g >= b; --g) {
            const double q = r_buf[g] * s;
            const double p = q / (1.0 + q);
            if (out_f32) static_cast<float *>(out)[g] = (float)p;
            else static_cast<double *>(out)[g] = p;
            const double w = u[g] * s;
            s = (m10 + m11 * w) / (1.0 + m01 * w);
        }
    }
}

// One CTA per contig (grid-stride), for batches of long contigs: 256 lanes share a chain instead of 32, so a
// 5,000-gene contig is ~20 sequential steps per lane, and 100 such contigs fill the machine (one warp per contig
// leaves all but 100 warps idle).  Same mathematics as chain_kernel: per-lane segment products, an inclusive scan
// (shuffles inside a warp, the eight warp totals through shared memory), replay from the exact incoming message.
__global__ void __launch_bounds__(kThreads)
chain_block_kernel(const CsrDev csr, const double *__restrict__ u, double *__restrict__ r_buf, void *__restrict__ out,
                   int out_f32, double m01, double m10, double m11) {
    constexpr int kWarps = kThreads / 32;
    __shared__ M2 sTot[kWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int64_t c = blockIdx.x; c < csr.C; c += gridDim.x) {
        const int64_t g0 = __ldg(csr.contig_ptr + c), g1 = __ldg(csr.contig_ptr + c + 1);
        const int64_t n = g1 - g0;
        if (n <= 0) continue;
        const int64_t seg = (n + kThreads - 1) / kThreads;
        int64_t b = g0 + (int64_t)tid * seg, e = b + seg;
        if (b > g1) b = g1;
        if (e > g1) e = g1;

        // ---- forward
        M2 S = identity();
        for (int64_t g = b; g < e; ++g) {
            const double ug = u[g];
            const M2 F = (g == g0) ? M2{0.0, ug, 0.0, 1.0} : M2{ug * m11, ug * m01, m10, 1.0};
            S = compose(F, S);
        }
        M2 P = S;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const M2 Y = shfl_up(P, d);
            if (lane >= d) P = compose(P, Y);
        }
        if (lane == 31) sTot[warp] = P;
        __syncthreads();
        M2 X = shfl_up(P, 1);  // earlier lanes of this warp
        if (lane == 0) X = identity();
        for (int w = warp - 1; w >= 0; --w) X = compose(X, sTot[w]);  // then the earlier warps, nearest first
        double r = 0.0;
        if (tid > 0) r = X.b / X.d;  // first column of X is zero: every chain starts with F_0
        for (int64_t g = b; g < e; ++g) {
            const double ug = u[g];
            r = (g == g0) ? ug : ((m01 + m11 * r) / (1.0 + m10 * r)) * ug;
            r_buf[g] = r;
        }
        __syncthreads();  // sTot is reused below

        // ---- backward
        M2 R = identity();
        for (int64_t g = b; g < e; ++g) {
            if (g == g0) continue;
            const double ug = u[g];
            const M2 B = M2{m11 * ug, m10, m01 * ug, 1.0};
            R = compose(R, B);
        }
        M2 Q = R;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const M2 Y = shfl_down(Q, d);
            if (lane + d < 32) Q = compose(Q, Y);
        }
        if (lane == 0) sTot[warp] = Q;
        __syncthreads();
        M2 E = shfl_down(Q, 1);  // later lanes of this warp
        if (lane == 31) E = identity();
        for (int w = warp + 1; w < kWarps; ++w) E = compose(E, sTot[w]);  // then the later warps, nearest first
        double s = (E.a + E.b) / (E.c + E.d);  // applied to s_{n-1} = 1
        for (int64_t g = e - 1; g >= b; --g) {
            const double q = r_buf[g] * s;
            const double p = q / (1.0 + q);
            if (out_f32) static_cast<float *>(out)[g] = (float)p;
            else static_cast<double *>(out)[g] = p;
            const double w = u[g] * s;
            s = (m10 + m11 * w) / (1.0 + m01 * w);
        }
        __syncthreads();  // before the next contig overwrites sTot
    }
}

}  // namespace

cudaError_t launch_chain(const ChainArgs &args, int num_sms, cudaStream_t stream, int64_t *launches) {
    if (args.csr.G <= 0) return cudaSuccess;
    double *u = args.scratch;
    double *r = args.scratch + args.csr.G;
    int64_t blocks_u = (args.csr.G + kThreads - 1) / kThreads;
    if (blocks_u > (int64_t)num_sms * 16) blocks_u = (int64_t)num_sms * 16;
    unary64_kernel<<<(int)blocks_u, kThreads, 0, stream>>>(args.csr, args.table64, args.model.A, u);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    // a warp per contig, or a CTA per contig when the contigs are long on average (GCRF_CHAIN_TEAM=warp|block overrides)
    const char *team = getenv("GCRF_CHAIN_TEAM");
    const bool block_team = team ? team[0] == 'b' : args.csr.G / args.csr.C >= 256;
    if (block_team) {
        int64_t blocks_c = args.csr.C;
        if (blocks_c > (int64_t)num_sms * 8) blocks_c = (int64_t)num_sms * 8;
        chain_block_kernel<<<(int)blocks_c, kThreads, 0, stream>>>(args.csr, u, r, args.out, args.out_f32, args.m01, args.m10,
                                                                   args.m11);
    } else {
        int64_t blocks_c = (args.csr.C * 32 + kThreads - 1) / kThreads;
        if (blocks_c > (int64_t)num_sms * 8) blocks_c = (int64_t)num_sms * 8;
        chain_kernel<<<(int)blocks_c, kThreads, 0, stream>>>(args.csr, u, r, args.out, args.out_f32, args.m01, args.m10,
                                                             args.m11);
    }
    if (launches) *launches += 2;
    return cudaGetLastError();
}

}  // namespace gcrf
