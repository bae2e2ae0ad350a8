// gcrf_windowed.cu — fused sm_100a kernel for ClusterCRF.predict_probabilities' hot loop
// (reference: gecco/crf/__init__.py:209-258; arithmetic of the third-party tagger it calls at :253,
// restated in SURVEY.md Appendix B).
//
// One persistent CTA walks a contiguous range of gene tiles.  Per tile of `tile_out` genes (plus a
// (W-1)-gene halo on both sides, because every gene is covered by the W windows that contain it):
//
//   A. gene_ptr slice -> shared memory (rebased to the tile's first attribute);
//   B. contig_ptr slice -> shared memory; each thread finds the contig of "its" gene;
//   C. gather: the tile's attr_idx range is streamed with coalesced 16-byte loads, every id is
//      resolved through the delta table staged in shared memory (delta_a = W[a][pos]-W[a][other],
//      slot A = 0 for unknown ids) and the per-gene sums are formed            [HBM-bound part]
//   D. u_g = exp(clamp(delta_g)) — the odds of the gene's unary potential;
//   E. one thread per window start: W-step forward and backward recursions on ODDS RATIOS
//        fwd  r_k = u_k (m01 + r_{k-1} m11) / (1 + r_{k-1} m10)
//        bwd  s_k = (m10 + m11 u_{k+1} s_{k+1}) / (1 + m01 u_{k+1} s_{k+1}),  s_{W-1} = 1
//      and odds of the marginal q_k = r_k s_k, written to a [W][threads] matrix in shared memory;
//      short contigs use the reference's padding (delta/2 empty items in front, :226-227);
//   F. one thread per gene: max over the windows covering it (max of odds == max of
//      probabilities), p = q/(1+q), store.
//
// FP32 throughout; odds stay finite because |delta_g| is clamped (ModelDev::clamp) — a clamp that
// only bites where the f64 marginal is already within 1e-20 of 0 or 1.
#include "gcrf_kernels.cuh"

#include <cfloat>
#include <climits>

namespace gcrf {

namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 8192;  // attribute ids staged per gather round
constexpr int kMaxWindow = 128;

__host__ __device__ constexpr int round_up4(int x) { return (x + 3) & ~3; }

struct Tiling {
    int W, tile_out, ng_max;
    int off_ptr, off_cp, off_u, off_big;  // float/int offsets into dynamic smem
    int big_elems;
    __host__ __device__ Tiling(int A, int window) {
        W = window;
        tile_out = (kThreads - window) & ~3;
        ng_max = tile_out + 2 * (window - 1);
        off_ptr = round_up4(A + 1);
        off_cp = off_ptr + round_up4(ng_max + 1);
        off_u = off_cp + round_up4(ng_max + 2);
        off_big = off_u + round_up4(ng_max);
        const int pool = window * kThreads;
        big_elems = pool > kChunk ? pool : kChunk;
    }
    __host__ __device__ size_t bytes() const { return sizeof(float) * (size_t)(off_big + big_elems); }
};

__device__ __forceinline__ int64_t load_gene_ptr(const CsrDev &csr, int64_t g) {
    return csr.gene_ptr64 ? __ldg(csr.gene_ptr64 + g) : (int64_t)__ldg(csr.gene_ptr32 + g);
}

// streaming 16-byte load of attribute ids: read once, keep out of L1
__device__ __forceinline__ int4 ld_stream_v4(const int32_t *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// a * (1/b) with one MUFU.RCP; operands here are positive and far from the FP32 range limits
__device__ __forceinline__ float fast_div(float a, float b) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return a * r;
}

__device__ __forceinline__ float lookup(const float *sTab, int32_t id, uint32_t A) {
    return sTab[min((uint32_t)id, A)];  // ids outside [0, A) (e.g. -1) hit the zero slot A
}

// Largest c in [0, C) with contig_ptr[c] <= g, found by one warp with 32 probes per round.
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

// Forward/backward on odds ratios for one window.  Position k of the window is gene `base + k`
// when klo <= k < khi and an empty (padding) item otherwise.  WT > 0: window size known at compile
// time, forward ratios live in registers.  WT == 0: runtime W, forward ratios are parked in the pool.
template <int WT, bool MASKED>
__device__ __forceinline__ void dp_window(const float *__restrict__ sU, float *__restrict__ pool,
                                          int W, int tid, int base, int klo, int khi, float m01,
                                          float m10, float m11) {
    auto unary = [&](int k) -> float {
        if (MASKED) {
            float u = 1.0f;
            if (k >= klo && k < khi) u = sU[base + k];
            return u;
        }
        return sU[base + k];
    };
    if constexpr (WT > 0) {
        float ra[WT];
        float r = unary(0);
        ra[0] = r;
#pragma unroll
        for (int k = 1; k < WT; ++k) {
            const float u = unary(k);
            const float num = fmaf(r, m11, m01), den = fmaf(r, m10, 1.0f);
            r = fast_div(num * u, den);
            ra[k] = r;
        }
        float s = 1.0f;
        pool[(WT - 1) * kThreads + tid] = ra[WT - 1];
#pragma unroll
        for (int k = WT - 2; k >= 0; --k) {
            const float w = unary(k + 1) * s;
            s = fast_div(fmaf(w, m11, m10), fmaf(w, m01, 1.0f));
            pool[k * kThreads + tid] = ra[k] * s;
        }
    } else {
        float r = unary(0);
        pool[tid] = r;
        for (int k = 1; k < W; ++k) {
            const float u = unary(k);
            const float num = fmaf(r, m11, m01), den = fmaf(r, m10, 1.0f);
            r = fast_div(num * u, den);
            pool[k * kThreads + tid] = r;
        }
        float s = 1.0f;
        for (int k = W - 2; k >= 0; --k) {
            const float w = unary(k + 1) * s;
            s = fast_div(fmaf(w, m11, m10), fmaf(w, m01, 1.0f));
            pool[k * kThreads + tid] *= s;
        }
    }
}

template <int WT>
__global__ void __launch_bounds__(kThreads, 4)
windowed_kernel(const WindowedArgs args, const int64_t num_tiles, const int tiles_per_cta) {
    const CsrDev &csr = args.csr;
    const int W = WT > 0 ? WT : args.window;
    const Tiling tl(args.model.A, W);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t A = (uint32_t)args.model.A;
    const float m01 = args.model.m01, m10 = args.model.m10, m11 = args.model.m11;
    const float clampv = args.model.clamp;
    const int step = args.step;

    extern __shared__ __align__(16) float smem[];
    float *sTab = smem;
    int *sPtr = reinterpret_cast<int *>(smem + tl.off_ptr);
    int *sCp = reinterpret_cast<int *>(smem + tl.off_cp);
    float *sU = smem + tl.off_u;
    float *sBig = smem + tl.off_big;
    __shared__ int64_t sCursor;

    // delta table: staged once per CTA, reused by every tile it walks
    for (int a = tid; a <= (int)A; a += kThreads) sTab[a] = __ldg(args.model.table + a);

    // consecutive tiles per CTA, spread evenly: the first num_tiles % grid CTAs take one more than the others
    (void)tiles_per_cta;
    const int64_t tiles_base = num_tiles / gridDim.x, tiles_rem = num_tiles % gridDim.x;
    const int64_t tile_begin = (int64_t)blockIdx.x * tiles_base + ((int64_t)blockIdx.x < tiles_rem ? (int64_t)blockIdx.x : tiles_rem);
    const int64_t tile_end = tile_begin + tiles_base + ((int64_t)blockIdx.x < tiles_rem ? 1 : 0);

    const int G = (int)csr.G;  // G < 2^31 (checked on the host): tile arithmetic stays 32-bit
    int64_t c_first = 0;
    for (int64_t tile = tile_begin; tile < tile_end; ++tile) {
        const int T0 = (int)tile * tl.tile_out;
        const int lo = T0 < W - 1 ? T0 : W - 1;  // halo genes in front of the tile
        const int Gs = T0 - lo;                  // first gene staged
        const int nout = min(G - T0, tl.tile_out);
        const int ngt = min(lo + nout + W - 1, G - Gs);  // genes staged: halo + tile + halo

        // ---- A. gene_ptr slice, rebased to the first attribute of the tile
        const int64_t p0 = load_gene_ptr(csr, Gs);
        for (int j = tid; j <= ngt; j += kThreads) sPtr[j] = (int)(load_gene_ptr(csr, Gs + j) - p0);

        // ---- B. contig slice: sCp[k] = contig_ptr[c_first + k] - Gs, k = 0 .. ngt+1
        if (tile == tile_begin) {
            if (warp == 0) {
                const int64_t c = warp_find_contig(csr.contig_ptr, csr.C, Gs + csr.gene_base, lane);
                if (lane == 0) sCursor = c;
            }
            __syncthreads();
        }
        c_first = sCursor;
        for (int k = tid; k <= ngt + 1; k += kThreads) {
            const int64_t c = c_first + k;
            sCp[k] = c <= csr.C ? __ldg(csr.contig_ptr + c) - (Gs + (int)csr.gene_base) : INT_MAX;
        }
        __syncthreads();  // sTab, sPtr, sCp visible

        // contig [c0, c1) of gene `tid`, in tile-local gene coordinates (c0 may be negative,
        // c1 may exceed ngt: the contig continues outside the staged range)
        int c0 = 0, c1 = 0;
        if (tid < ngt) {
            // sCp[ngt+1] > tid for a valid batch (contig_ptr strictly increasing); the bound also ends the search on
            // a malformed device-pointer batch instead of spinning
            int hi = 1;
            while (hi < ngt + 1 && sCp[hi] <= tid) hi <<= 1, hi = hi > ngt + 1 ? ngt + 1 : hi;
            int lo_k = hi >> 1;
            if (sCp[lo_k] > tid) lo_k = 0;  // only when hi was clamped to a non power of two
            while (hi - lo_k > 1) {
                const int mid = (lo_k + hi) >> 1;
                if (sCp[mid] <= tid) lo_k = mid; else hi = mid;
            }
            c0 = sCp[lo_k];
            c1 = sCp[lo_k + 1];
        }

        // ---- C. gather: per-gene sums of delta over the tile's attribute ids
        const int tile_nnz = sPtr[ngt];
        const int j1 = tid + kThreads;  // second gene owned by this thread (ng_max < 2*threads)
        float acc0 = 0.0f, acc1 = 0.0f;
        if (tile_nnz > 0) {
            const int mis = (int)(p0 & 3);  // attr_idx is 16-byte aligned (checked on host): start on a vector
            const int32_t *src = csr.attr_idx + (p0 - mis);
            const int64_t avail = csr.nnz - (p0 - mis);  // elements readable from src
            const int total = tile_nnz + mis;            // elements to stage, counted from src
            for (int cb = 0; cb < total; cb += kChunk) {
                const int span = min(total - cb, kChunk);
                const int nvec = (span + 3) >> 2;
                const int64_t safe64 = (avail - cb) >> 2;  // full vectors inside attr_idx
                const int safe = safe64 < nvec ? (int)safe64 : nvec;
                constexpr int kBatch = 4;
                for (int v0 = tid; v0 < nvec; v0 += kBatch * kThreads) {
                    int4 id[kBatch];
#pragma unroll
                    for (int b = 0; b < kBatch; ++b) {
                        const int v = v0 + b * kThreads;
                        id[b] = make_int4(-1, -1, -1, -1);
                        if (v < safe) {
                            id[b] = ld_stream_v4(src + cb + 4 * v);
                        } else if (v < nvec) {
                            const int64_t e = (int64_t)cb + 4 * v;
                            if (e + 0 < avail) id[b].x = __ldg(src + e + 0);
                            if (e + 1 < avail) id[b].y = __ldg(src + e + 1);
                            if (e + 2 < avail) id[b].z = __ldg(src + e + 2);
                        }
                    }
#pragma unroll
                    for (int b = 0; b < kBatch; ++b) {
                        const int v = v0 + b * kThreads;
                        if (v < nvec) {
                            float4 d;
                            d.x = lookup(sTab, id[b].x, A);
                            d.y = lookup(sTab, id[b].y, A);
                            d.z = lookup(sTab, id[b].z, A);
                            d.w = lookup(sTab, id[b].w, A);
                            reinterpret_cast<float4 *>(sBig)[v] = d;
                        }
                    }
                }
                __syncthreads();
                const int cbase = cb - mis;  // chunk start relative to the tile's first attribute
                if (tid < ngt) {
                    const int s = max(sPtr[tid] - cbase, 0), e = min(sPtr[tid + 1] - cbase, kChunk);
                    for (int p = s; p < e; ++p) acc0 += sBig[p];
                }
                if (j1 < ngt) {
                    const int s = max(sPtr[j1] - cbase, 0), e = min(sPtr[j1 + 1] - cbase, kChunk);
                    for (int p = s; p < e; ++p) acc1 += sBig[p];
                }
                __syncthreads();
            }
        }

        // ---- D. unary odds
        if (tid < ngt) sU[tid] = expf(fminf(fmaxf(acc0, -clampv), clampv));
        if (j1 < ngt) sU[j1] = expf(fminf(fmaxf(acc1, -clampv), clampv));
        __syncthreads();

        // ---- E. one thread per window start
        {
            const int n = c1 - c0;
            bool valid = false;
            int base = tid, klo = 0, khi = W;
            if (tid < lo + nout && tid < ngt) {
                if (n >= W) {
                    valid = tid <= c1 - W && (step == 1 || (tid - c0) % step == 0);
                } else if (args.pad && tid == c0) {
                    // gecco/crf/__init__.py:226-227: delta//2 empty items in front, the rest behind
                    valid = true;
                    klo = (W - n) >> 1;
                    khi = klo + n;
                    base = tid - klo;
                }
            }
            if (valid) {
                if (klo == 0 && khi == W) dp_window<WT, false>(sU, sBig, W, tid, base, klo, khi, m01, m10, m11);
                else dp_window<WT, true>(sU, sBig, W, tid, base, klo, khi, m01, m10, m11);
            }
        }
        __syncthreads();

        // ---- F. one thread per gene: max over the covering windows, odds -> probability
        if (tid >= lo && tid < lo + nout) {
            const int n = c1 - c0;
            float q = 0.0f;
            bool skipped = false;
            if (n >= W) {
                int kmin = tid - (c1 - W);
                kmin = kmin < 0 ? 0 : kmin;
                int kmax = tid - c0;
                kmax = kmax > W - 1 ? W - 1 : kmax;
                const bool interior = kmin == 0 && kmax == W - 1 && step == 1;
                if (WT > 0 && __all_sync(__activemask(), interior)) {
#pragma unroll
                    for (int k = 0; k < (WT > 0 ? WT : 1); ++k) q = fmaxf(q, sBig[k * kThreads + tid - k]);
                } else {
                    for (int k = kmin; k <= kmax; ++k)
                        if (step == 1 || (tid - k - c0) % step == 0) q = fmaxf(q, sBig[k * kThreads + tid - k]);
                }
            } else if (args.pad) {
                q = sBig[(tid - c0 + ((W - n) >> 1)) * kThreads + c0];
            } else {
                skipped = true;  // :228-234 — contig too short and padding disabled
            }
            float p = fminf(fast_div(q, 1.0f + q), 1.0f);  // the approximate reciprocal can land one ulp above
            if (skipped) p = __int_as_float(0x7fc00000);
            const int g = T0 + (tid - lo);
            if (args.out_f32) static_cast<float *>(args.out)[g] = p;
            else static_cast<double *>(args.out)[g] = (double)p;
        }

        // ---- cursor for the next tile: its first staged gene lies inside this tile's contig slice
        {
            const int T1 = T0 + tl.tile_out;
            const int x = T1 - (W - 1) - Gs;  // first staged gene of the next tile, in this tile's coordinates
            for (int k = tid; k <= ngt; k += kThreads)
                if (sCp[k] <= x && x < sCp[k + 1]) sCursor = c_first + k;
        }
        __syncthreads();  // sCursor written, pool and slices free for the next tile
    }
}

template <int WT>
cudaError_t configure(const Tiling &tl, int *ctas_per_sm) {
    cudaError_t err = cudaFuncSetAttribute(windowed_kernel<WT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)tl.bytes());
    if (err != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, windowed_kernel<WT>, kThreads, tl.bytes());
}

}  // namespace

int windowed_max_window(int A) {
    int w = kMaxWindow;
    while (w > 1 && Tiling(A, w).bytes() > 227 * 1024) --w;
    return w;
}

cudaError_t plan_windowed(const WindowedArgs &args, int num_sms, WindowedPlan *plan) {
    if (args.window <= 0 || args.window > kMaxWindow) return cudaErrorInvalidValue;
    const Tiling tl(args.model.A, args.window);
    if (tl.bytes() > 227 * 1024) return cudaErrorInvalidValue;
    int per_sm = 0;
    cudaError_t err;
    switch (args.window) {
        case 20: err = configure<20>(tl, &per_sm); break;
        case 5: err = configure<5>(tl, &per_sm); break;
        default: err = configure<0>(tl, &per_sm); break;
    }
    if (err != cudaSuccess) return err;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    plan->threads = kThreads;
    plan->tile_out = tl.tile_out;
    plan->chunk = kChunk;
    plan->smem_bytes = tl.bytes();
    plan->num_tiles = (args.csr.G + tl.tile_out - 1) / tl.tile_out;
    plan->ctas_per_sm = per_sm;
    int64_t grid = (int64_t)num_sms * per_sm;
    if (grid > plan->num_tiles) grid = plan->num_tiles;
    if (grid < 1) grid = 1;
    plan->grid = (int)grid;
    plan->tiles_per_cta = (int)((plan->num_tiles + grid - 1) / grid);
    return cudaSuccess;
}

cudaError_t launch_windowed(const WindowedArgs &args, const WindowedPlan &plan, cudaStream_t stream,
                            int64_t *launches) {
    if (args.csr.G <= 0) return cudaSuccess;
    const dim3 grid(plan.grid), block(plan.threads);
    switch (args.window) {
        case 20: windowed_kernel<20><<<grid, block, plan.smem_bytes, stream>>>(args, plan.num_tiles, plan.tiles_per_cta); break;
        case 5: windowed_kernel<5><<<grid, block, plan.smem_bytes, stream>>>(args, plan.num_tiles, plan.tiles_per_cta); break;
        default: windowed_kernel<0><<<grid, block, plan.smem_bytes, stream>>>(args, plan.num_tiles, plan.tiles_per_cta); break;
    }
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace gcrf
