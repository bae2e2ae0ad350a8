// gcrf_device.cuh — device helpers shared by the windowed kernels (gcrf_stream.cu):
// fast reciprocal / exponential, the fixed-point table look-up, mbarrier + bulk-async copy (TMA) wrappers,
// contig searches and the rare slow paths (direct row sums, padded short contigs).  sm_100a only.
#pragma once

#include "gcrf_kernels.cuh"

namespace gcrf {
namespace {

#ifndef GCRF_KWALK
#define GCRF_KWALK 52
#endif
constexpr int kWalk = GCRF_KWALK;  // ids walked per thread and staging round (4 * odd: conflict-free LDS.128); -DGCRF_KWALK: A/B builds

__host__ __device__ constexpr int round_up4s(int x) { return (x + 3) & ~3; }

__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// exp(x) for |x| <= ~40 with ~2e-7 relative error: exponent split in two floats, one MUFU.EX2
__device__ __forceinline__ float exp_fast(float x) {
    const float l2e_hi = 1.44269502162933349609375f, l2e_lo = 1.925963033500011e-8f;
    const float t = x * l2e_hi;
    const float lo = fmaf(x, l2e_hi, -t) + x * l2e_lo;
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return fmaf(r, lo * 0.693147180559945f, r);
}

// x / kWalk for 0 <= x < 13376 (kWalk = 52): one multiply and one shift
__device__ __forceinline__ int walk_thread(int x) {
    if constexpr (kWalk == 52) return (int)(((unsigned)x * 10083u) >> 19);
    else return x / kWalk;  // A/B builds
}

__device__ __forceinline__ int lookup(const int *sTab, int32_t id, uint32_t A) {
    return sTab[min((uint32_t)id, A)];  // ids outside [0, A) (e.g. -1) hit the zero slot A
}

// ---- mbarrier / bulk-async copy (TMA) wrappers ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    // order earlier generic-proxy accesses to the buffer (the -1 masks, the walk's stores) before the async write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Stage ids [pa, pb) (from their 16-byte aligned start) if they fit one round; otherwise only arrive, and the
// tile takes the direct path.
template <int kCap>
__device__ __forceinline__ void stage_ids(int32_t *sIdx, const int32_t *attr_idx, int64_t pa, int64_t pb, uint64_t *bar) {
    const int64_t a0 = pa & ~(int64_t)3;
    const int64_t total = pb - a0;
    if (total > 0 && total <= kCap) {
        const uint32_t bytes = (uint32_t)(4 * ((total + 3) & ~(int64_t)3));
        mbar_expect_tx(bar, bytes);
        tma_load_1d(sIdx, attr_idx + a0, bytes, bar);
    } else {
        mbar_arrive(bar);
    }
}

// One result of the streaming kernel -> the local output array and, for a contig-sharded batch, straight into the
// output arrays of the peer GPUs (plain stores through the NVLink peer mapping, or one multimem.st to an NVLS multicast
// address that the switch replicates): the "gather" of the marginals happens while the kernel computes.
__device__ __forceinline__ void store_result(const WindowedArgs &args, int gg, float p) {
    if (args.out_f32) {
        if (args.out) static_cast<float *>(args.out)[gg] = p;
        if (args.n_peer_out) {
            if (args.peer_multicast) {
                asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(static_cast<float *>(args.peer_out[0]) + gg), "f"(p) : "memory");
            } else {
                for (int k = 0; k < args.n_peer_out; ++k) static_cast<float *>(args.peer_out[k])[gg] = p;
            }
        }
    } else {
        const double pd = (double)p;
        if (args.out) static_cast<double *>(args.out)[gg] = pd;
        if (args.n_peer_out) {
            if (args.peer_multicast) {
                asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(static_cast<double *>(args.peer_out[0]) + gg), "d"(pd) : "memory");
            } else {
                for (int k = 0; k < args.n_peer_out; ++k) static_cast<double *>(args.peer_out[k])[gg] = pd;
            }
        }
    }
}

// Largest c in [0, C) with contig_ptr[c] <= g (g >= 0), one warp, 32 probes per round.
__device__ int64_t warp_find_contig(const int32_t *contig_ptr, int64_t C, int64_t g, int lane) {
    int64_t lo = 0, hi = C;
    while (hi - lo > 1) {
        const int64_t span = hi - lo;
        const int64_t st = (span + 32) / 33;
        const int64_t probe = lo + (int64_t)(lane + 1) * st;
        const bool ok = probe < hi && (int64_t)__ldg(contig_ptr + probe) <= g;
        const int cnt = __popc(__ballot_sync(0xffffffffu, ok));
        const int64_t nhi = lo + (int64_t)(cnt + 1) * st;
        lo += (int64_t)cnt * st;
        hi = nhi < hi ? nhi : hi;
    }
    return lo;
}

// Largest k in [0, kmax] with sCp[k] <= j.  Requires sCp[0] <= j and sCp[kmax + 1] > j.
__device__ __forceinline__ int find_slice_contig(const int *sCp, int j, int kmax) {
    int k = 0, hi = kmax + 1;
    while (hi - k > 1) {
        const int mid = (k + hi) >> 1;
        if (sCp[mid] <= j) k = mid; else hi = mid;
    }
    return k;
}

// Unary odds of one gene straight from global memory (prologue halo, oversized tiles, very long rows).
template <typename PtrT>
__device__ __noinline__ float direct_unary(const PtrT *gene_ptr, const int32_t *attr_idx, const float *table, uint32_t A,
                                           int g, float clampv) {
    const int64_t rb = (int64_t)__ldg(gene_ptr + g), re = (int64_t)__ldg(gene_ptr + g + 1);
    float delta = 0.f;
    for (int64_t p = rb; p < re; ++p) delta += __ldg(table + min((uint32_t)__ldg(attr_idx + p), A));
    return exp_fast(fminf(fmaxf(delta, -clampv), clampv));
}

// Window of a padded short contig (gecco/crf/__init__.py:216-227): n < W genes starting at local gene j,
// (W-n)/2 empty items in front and the rest behind; writes the odds of its n genes to sQ.
template <int W>
__device__ __noinline__ void padded_window(const float *sU0, float *sQ, int j, int n, float m01, float m10, float m11) {
    const int front = (W - n) >> 1;
    float ra[W];
    float rr = 0.f;
#pragma unroll
    for (int q = 0; q < W; ++q) {
        const int p = q - front;
        const float u = (p >= 0 && p < n) ? sU0[j + p] : 1.0f;
        rr = q == 0 ? u : (fmaf(rr, m11, m01) * u) * rcp_fast(fmaf(rr, m10, 1.0f));
        ra[q] = rr;
    }
    float ss = 1.0f;
#pragma unroll
    for (int q = W - 1; q >= 0; --q) {
        const int p = q - front;
        const bool real = p >= 0 && p < n;
        if (real) sQ[j + p] = ra[q] * ss;
        if (q > 0) {
            const float w = (real ? sU0[j + p] : 1.0f) * ss;
            ss = fmaf(w, m11, m10) * rcp_fast(fmaf(w, m01, 1.0f));
        }
    }
}

}  // namespace
}  // namespace gcrf
