// gcrf_exact.cu — GCRF_FLAG_F64: the windowed marginals in the reference's own arithmetic.
//
// The default device path folds the two-label model into odds ratios and computes in FP32 (gcrf_stream.cu); its
// results stay within 1e-5 of the reference's.  This path instead evaluates, in f64 and OPERATION BY OPERATION,
// what the reference's tagger evaluates for every window (gecco/crf/__init__.py:253 -> sklearn-crfsuite ->
// python-crfsuite -> CRFsuite 0.12 crf1d_tag.c / crf1d_context.c; restated in SURVEY.md Appendix B):
//
//     state score   s_t[l] = sum over the item's attributes, in row order, of W[a][l]      (0.0 for an empty item)
//     E_t[l] = exp(s_t[l]),  M[i][j] = exp(T[i][j])
//     forward       alpha_0 = E_0;  alpha_t[j] = (alpha_{t-1}[0] M[0][j] + alpha_{t-1}[1] M[1][j]) E_t[j]
//                   c_t = 1 / (alpha_t[0] + alpha_t[1])  (1 when the sum is 0);  alpha_t *= c_t
//     backward      beta_{W-1} = c_{W-1};  beta_t[i] = (M[i][0] E_{t+1}[0] beta_{t+1}[0] + M[i][1] E_{t+1}[1] beta_{t+1}[1]) c_t
//     marginal      alpha_t[l] beta_t[l] / c_t
//
// with the same association of every sum and product, IEEE round-to-nearest and no fused multiply-adds (the file is
// compiled with -fmad=false and spells the arithmetic with __dmul_rn / __dadd_rn / __ddiv_rn), and exp() rounded
// correctly (gcrf_exp.cuh).  On the reference's golden fixture the results are bit-identical to python-crfsuite's
// (tests/golden/bgc0001866.json, 17 digits); everywhere else they agree with the f64 oracle to the last few ulps,
// the residue being the host libm's own rounding of exp().
//
// Any window size is accepted (the reference takes any window_size >= 1, gecco/crf/__init__.py:134-137): W = 20 keeps
// the stored half of the recursion in registers, other sizes park it in a global work area.
//
// Three launches: (1) per gene: both state scores, their exponentials, and the initial value of the max-pool (0.0, or
// NaN for the genes of a skipped short contig, :228-234); (2) per window: the recursion above, max-pooled into the
// output with an integer atomic max (for doubles >= 0 the bit patterns order like the values; a NaN marginal has the
// largest pattern and therefore wins, like numpy.maximum at :254); (3) only for float output: the narrowing pass.
#include "gcrf_exp.cuh"
#include <type_traits>
#include "gcrf_kernels.cuh"

namespace gcrf {

namespace {

constexpr int kExactThreads = 128;

// Largest c in [0, C) with contig_ptr[c] <= g
__device__ __forceinline__ int64_t find_contig(const int32_t *__restrict__ contig_ptr, int64_t C, int64_t g) {
    int64_t lo = 0, hi = C;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(contig_ptr + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ int64_t row_ptr(const CsrDev &csr, int64_t g) {
    return csr.gene_ptr64 ? __ldg(csr.gene_ptr64 + g) : (int64_t)__ldg(csr.gene_ptr32 + g);
}

__device__ __forceinline__ double quiet_nan() { return __longlong_as_double(0x7ff8000000000000ll); }

__global__ void __launch_bounds__(kExactThreads)
exact_unary_kernel(const ExactArgs args, double *__restrict__ pool) {
    const CsrDev &csr = args.csr;
    const uint32_t A = (uint32_t)args.A;
    for (int64_t g = (int64_t)blockIdx.x * kExactThreads + threadIdx.x; g < csr.G; g += (int64_t)gridDim.x * kExactThreads) {
        const int64_t rb = row_ptr(csr, g), re = row_ptr(csr, g + 1);
        double s0 = 0.0, s1 = 0.0;
        for (int64_t p = rb; p < re; ++p) {
            const uint32_t a = (uint32_t)__ldg(csr.attr_idx + p);
            if (a < A) {  // ids outside [0, A): attributes the model does not know are dropped
                const double2 w = __ldg(reinterpret_cast<const double2 *>(args.state_w) + a);
                s0 = __dadd_rn(s0, w.x);
                s1 = __dadd_rn(s1, w.y);
            }
        }
        reinterpret_cast<double2 *>(args.unary)[g] = make_double2(expdd::exp_cr(s0), expdd::exp_cr(s1));
        double init = 0.0;
        if (!args.pad) {  // genes of a contig shorter than the window keep "no probability" (:228-234)
            const int64_t c = find_contig(csr.contig_ptr, csr.C, g + csr.gene_base);
            const int n = __ldg(csr.contig_ptr + c + 1) - __ldg(csr.contig_ptr + c);
            if (n < args.window) init = quiet_nan();
        }
        pool[g] = init;
    }
}

// One forward step / the scaling, in the oracle's association.
struct Fwd {
    double a0, a1, c;
};
__device__ __forceinline__ Fwd fwd_first(double e0, double e1) {
    const double sum = __dadd_rn(e0, e1);
    const double c = sum != 0.0 ? __drcp_rn(sum) : 1.0;  // the correctly rounded 1 / sum, i.e. what 1.0 / sum is
    return {__dmul_rn(e0, c), __dmul_rn(e1, c), c};
}
__device__ __forceinline__ Fwd fwd_step(const Fwd &p, const double *M, double e0, double e1) {
    double n0 = __dadd_rn(__dmul_rn(p.a0, M[0]), __dmul_rn(p.a1, M[2]));
    double n1 = __dadd_rn(__dmul_rn(p.a0, M[1]), __dmul_rn(p.a1, M[3]));
    n0 = __dmul_rn(n0, e0);
    n1 = __dmul_rn(n1, e1);
    const double sum = __dadd_rn(n0, n1);
    const double c = sum != 0.0 ? __drcp_rn(sum) : 1.0;  // the correctly rounded 1 / sum, i.e. what 1.0 / sum is
    return {__dmul_rn(n0, c), __dmul_rn(n1, c), c};
}

// Largest c in [0, C) with contig_ptr[c] <= g, found by the whole warp: 32 probes per round trip (three dependent
// loads for 10,000 contigs instead of fourteen).  g is the same in all lanes; so is the result.
__device__ __forceinline__ int64_t warp_find_contig(const int32_t *__restrict__ contig_ptr, int64_t C, int64_t g, int lane) {
    int64_t lo = 0, hi = C;  // contig_ptr[lo] <= g < contig_ptr[hi]
    while (hi - lo > 1) {
        const int64_t st = (hi - lo + 31) / 32;
        const int64_t probe = lo + (int64_t)(lane + 1) * st;
        const bool le = probe < hi && (int64_t)__ldg(contig_ptr + probe) <= g;  // monotone in the lane
        const int cnt = __popc(__ballot_sync(0xffffffffu, le));
        const int64_t nlo = lo + (int64_t)cnt * st, nhi = nlo + st;
        lo = nlo;
        hi = nhi < hi ? nhi : hi;
    }
    return lo;
}

// WT > 0: compile-time window, alpha[pos] and the scales stay in registers.  WT == 0: runtime window, both live in
// `work` ([2*W][threads of the grid], coalesced).
// A warp takes 32 consecutive window starts: the contig of the first is found by the warp together, the others walk on
// from it.  Windows inside a contig of at least W genes — all of them, short contigs apart — read their items and
// write their marginals without a range test per position (CHECKED = false).
#ifndef GCRF_EXACT_MINB
#define GCRF_EXACT_MINB 4  // 128 registers (a few doubles of the stored half spill): 0.41 ms on config 2 against 0.47 at 3 x 152
#endif
template <int WT>
__global__ void __launch_bounds__(kExactThreads, GCRF_EXACT_MINB)
exact_window_kernel(const ExactArgs args, double *__restrict__ pool, double *__restrict__ work) {
    const CsrDev &csr = args.csr;
    const int W = WT > 0 ? WT : args.window;
    const int step = args.step, pos = args.pos_label;
    const double M[4] = {args.M[0], args.M[1], args.M[2], args.M[3]};
    const int64_t nthreads = (int64_t)gridDim.x * kExactThreads;
    const int64_t me = (int64_t)blockIdx.x * kExactThreads + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int G = (int)csr.G;  // the ABI keeps G below 2^31
    const int32_t *__restrict__ cp = csr.contig_ptr;
    const int gene_base = (int)csr.gene_base;
    const double2 *__restrict__ unary = reinterpret_cast<const double2 *>(args.unary);
    unsigned long long *__restrict__ pool_bits = reinterpret_cast<unsigned long long *>(pool);

    for (int64_t base = me - lane; base < G; base += nthreads) {  // the same trip count in all lanes of a warp
        const int i = (int)base + lane;
        int64_t c = warp_find_contig(cp, csr.C, base + gene_base, lane);
        if (i >= G) continue;
        while ((int)__ldg(cp + c + 1) - gene_base <= i) ++c;  // usually no step at all
        const int c0 = (int)__ldg(cp + c) - gene_base, c1 = (int)__ldg(cp + c + 1) - gene_base;
        const int n = c1 - c0;
        int first;  // gene of window position 0 (may lie in front of the contig for a padded window)
        if (n >= W) {
            // gecco/_meta.py:124-132: windows start at c0, c0 + step, ... while they fit
            if (i > c1 - W || (i - c0) % step != 0) continue;
            first = i;
        } else {
            // :216-227: one window over delta/2 empty items, the contig's genes, and the rest of the padding
            if (!args.pad || i != c0) continue;
            first = c0 - ((W - n) >> 1);
        }
        auto window = [&](auto checked) {
            constexpr bool CHECKED = decltype(checked)::value;
            auto item = [&](int k) -> double2 {  // exp of the state scores of window position k; an empty item scores 0
                const int g = first + k;
                if (CHECKED && !(g >= c0 && g < c1)) return make_double2(1.0, 1.0);
                return __ldg(unary + g);
            };
            auto emit = [&](int k, double alpha_p, double beta_p, double ck) {
                const int g = first + k;
                if (!CHECKED || (g >= c0 && g < c1)) {
                    const double m = __ddiv_rn(__dmul_rn(alpha_p, beta_p), ck);
                    atomicMax(pool_bits + g, (unsigned long long)__double_as_longlong(m));
                }
            };
            if constexpr (WT > 0) {
                double ap[WT], sc[WT];
                double2 e = item(0);
                Fwd f = fwd_first(e.x, e.y);
                ap[0] = pos ? f.a1 : f.a0;
                sc[0] = f.c;
#pragma unroll
                for (int k = 1; k < WT; ++k) {
                    e = item(k);
                    f = fwd_step(f, M, e.x, e.y);
                    ap[k] = pos ? f.a1 : f.a0;
                    sc[k] = f.c;
                }
                double b0 = sc[WT - 1], b1 = b0;
#pragma unroll
                for (int k = WT - 1; k >= 0; --k) {
                    emit(k, ap[k], pos ? b1 : b0, sc[k]);
                    if (k > 0) {
                        e = item(k);
                        const double r0 = __dmul_rn(b0, e.x), r1 = __dmul_rn(b1, e.y);
                        const double d0 = __dadd_rn(__dmul_rn(M[0], r0), __dmul_rn(M[1], r1));
                        const double d1 = __dadd_rn(__dmul_rn(M[2], r0), __dmul_rn(M[3], r1));
                        b0 = __dmul_rn(d0, sc[k - 1]);
                        b1 = __dmul_rn(d1, sc[k - 1]);
                    }
                }
            } else {
                double *ap = work + me, *sc = work + (int64_t)W * nthreads + me;  // element k at [k * nthreads]
                double2 e = item(0);
                Fwd f = fwd_first(e.x, e.y);
                ap[0] = pos ? f.a1 : f.a0;
                sc[0] = f.c;
                for (int k = 1; k < W; ++k) {
                    e = item(k);
                    f = fwd_step(f, M, e.x, e.y);
                    ap[(int64_t)k * nthreads] = pos ? f.a1 : f.a0;
                    sc[(int64_t)k * nthreads] = f.c;
                }
                double ck = f.c;
                double b0 = ck, b1 = ck;
                for (int k = W - 1; k >= 0; --k) {
                    emit(k, ap[(int64_t)k * nthreads], pos ? b1 : b0, ck);
                    if (k > 0) {
                        e = item(k);
                        ck = sc[(int64_t)(k - 1) * nthreads];
                        const double r0 = __dmul_rn(b0, e.x), r1 = __dmul_rn(b1, e.y);
                        const double d0 = __dadd_rn(__dmul_rn(M[0], r0), __dmul_rn(M[1], r1));
                        const double d1 = __dadd_rn(__dmul_rn(M[2], r0), __dmul_rn(M[3], r1));
                        b0 = __dmul_rn(d0, ck);
                        b1 = __dmul_rn(d1, ck);
                    }
                }
            }
        };
        if (n >= W) window(std::false_type{});
        else window(std::true_type{});
    }
}

__global__ void __launch_bounds__(256)
exact_narrow_kernel(const double *__restrict__ pool, float *__restrict__ out, int64_t G) {
    for (int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x; g < G; g += (int64_t)gridDim.x * 256) out[g] = (float)pool[g];
}

}  // namespace

// Threads of the window kernel for runtime window sizes: bounded so that the work area (2 * W doubles per thread)
// stays under 256 MB whatever the window.
int64_t exact_window_threads(int64_t G, int32_t window, int num_sms) {
    int64_t threads = (int64_t)num_sms * 8 * kExactThreads;
    const int64_t need = (G + kExactThreads - 1) / kExactThreads * kExactThreads;
    if (threads > need) threads = need;
    const int64_t cap = ((int64_t)256 << 20) / (16 * (int64_t)window) / kExactThreads * kExactThreads;
    if (threads > cap) threads = cap;
    return threads < kExactThreads ? kExactThreads : threads;
}

size_t exact_work_bytes(int64_t G, int32_t window, int num_sms) {
    if (window == 20) return 0;
    return (size_t)exact_window_threads(G, window, num_sms) * 2 * (size_t)window * sizeof(double);
}

cudaError_t launch_exact(const ExactArgs &args, double *work, int num_sms, cudaStream_t stream, int64_t *launches) {
    const int64_t G = args.csr.G;
    if (G <= 0) return cudaSuccess;
    double *pool = args.pool ? args.pool : static_cast<double *>(args.out);
    int64_t blocks = (G + kExactThreads - 1) / kExactThreads;
    const int64_t cap = (int64_t)num_sms * 16;
    exact_unary_kernel<<<(unsigned)(blocks < cap ? blocks : cap), kExactThreads, 0, stream>>>(args, pool);
    if (args.window == 20) {
        exact_window_kernel<20><<<(unsigned)(blocks < cap ? blocks : cap), kExactThreads, 0, stream>>>(args, pool, nullptr);
    } else {
        const int64_t threads = exact_window_threads(G, args.window, num_sms);
        exact_window_kernel<0><<<(unsigned)(threads / kExactThreads), kExactThreads, 0, stream>>>(args, pool, work);
    }
    int64_t n = 2;
    if (args.out_f32) {
        exact_narrow_kernel<<<(unsigned)(num_sms * 4), 256, 0, stream>>>(pool, static_cast<float *>(args.out), G);
        ++n;
    }
    if (launches) *launches += n;
    return cudaGetLastError();
}

}  // namespace gcrf
