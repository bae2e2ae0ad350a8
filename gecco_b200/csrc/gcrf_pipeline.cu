// gcrf_pipeline.cu — the windowed path (W = 20) as two co-resident kernels: a producer that turns attribute ids
// into unary odds and a consumer that runs the window DP, chained with programmatic dependent launch and
// per-run progress flags so that both execute AT THE SAME TIME on every SM.
//
// Why: the fused kernel (gcrf_stream.cu) runs gather -> DP -> pool back to back inside each CTA; its time on
// config 2 (88 us) is the SUM of its phases (HBM stream ~31 us, shared-memory look-ups ~28 us, DP ~31 us):
// the gather saturates the shared-memory pipe while the FP32/MUFU pipes idle, then the DP does the opposite.
// Splitting the two roles into different CTAs gives each its own register budget (gather ~48, DP ~124) and
// lets the hardware overlap them instead of hoping that four identical CTAs drift out of phase.
//
//   unary_kernel  (K1, reference gecco/crf/features.py:13-35 + the tagger's state scores)
//       persistent; CTA c owns a contiguous run of genes, tiles of 240 genes.  attr_idx of a tile arrives by ONE
//       bulk-async copy (TMA) into a 1- or 2-stage ring; thread t walks 52 ids, resolving each through the
//       fixed-point delta table in shared memory and overwriting it with the running prefix sum; a row sum is
//       a difference of prefixes (exact, order independent); u_g = exp(clamp(delta_g)) goes to a global f32
//       array (L2 resident: 4 B/gene).  After each tile: progress[c] = (epoch << 32 | tiles done), released
//       at gpu scope.
//   window_kernel (K2, reference gecco/crf/__init__.py:209-258, _meta.py:124-132)
//       persistent; CTA c owns the same run, tiles of 236 output genes (256 window slots).  Before a tile it
//       acquires progress[c] (and progress[c+1] for the run's last tile), loads the tile's 275 unary odds with
//       L2-only loads, then: contig bookkeeping, two windows per thread in packed f32x2 registers with the
//       forward and backward odds advancing together, max-pool through shared memory, p = q/(1+q).
//
// Ordering / liveness: K1 never waits for K2.  K2 is launched with programmatic stream serialization, so its
// CTAs become resident only after every K1 CTA has started (griddepcontrol.launch_dependents is K1's first
// instruction) — K1 is then resident or finished, and the flags K2 spins on always arrive.  K2 touches the
// batch only after it has seen a flag of the current epoch (K1 sets flags only after griddepcontrol.wait).
// Under a profiler or sanitizer the kernels serialise and the flags are already set.  A spin that lasts two
// seconds traps instead of hanging the device.
#include "gcrf_device.cuh"

#include <climits>
#include <cstdlib>

namespace gcrf {

namespace {

constexpr int kW = 20;            // window size of this path
constexpr int kNT = 128;          // threads per CTA, both kernels
constexpr int kGatherTile = 240;  // genes per gather tile (avg 6,000 ids at 25 domains/gene for 6,656 staged)
constexpr int kCap = kNT * kWalk; // ids staged per gather tile
constexpr int kHalo = kW;         // genes in front of a run that its own gather CTA computes as well

struct WinTiling {
    static constexpr int kSlots = 2 * kNT;
    static constexpr int kPitch = kNT + 16;
    static constexpr int lo = kW;
    static constexpr int tile_out = (kSlots - kW) & ~3;  // 236
    static constexpr int ng = kSlots + kW - 1;           // 275
    static constexpr int off_pool = 0;
    static constexpr int off_u0 = round_up4s((kW + 1) * kPitch);
    static constexpr int off_u1 = off_u0 + round_up4s(ng + 2);
    static constexpr int off_q = off_u1 + round_up4s(ng + 2);
    static constexpr int off_cp = off_q + round_up4s(ng + 2);
    static constexpr int off_stat = off_cp + round_up4s(ng + 4);
    static constexpr int words = off_stat + round_up4s((ng + 8) / 4);
};

struct PipeGeom {
    int grid;        // runs = CTAs of each kernel
    int tpc;         // window tiles per run
    int num_tiles;   // window tiles = ceil(G / 236)
    uint32_t epoch;
};

struct PipeBuffers {
    float *u;                       // [G] unary odds
    float *halo;                    // [grid][kHalo] unary odds of the kHalo genes in front of each run
    unsigned long long *progress;   // [grid]
};

__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Spin until *p >= want (same epoch in the high half); returns the value seen.
__device__ __noinline__ unsigned long long wait_progress(const unsigned long long *p, unsigned long long want) {
    unsigned long long v = ld_acquire(p);
    if (v >= want) return v;
    const unsigned long long t0 = global_ns();
    unsigned ns = 20;
    for (;;) {
        __nanosleep(ns);
        if (ns < 160) ns += ns;
        v = ld_acquire(p);
        if (v >= want) return v;
        if (global_ns() - t0 > 2000000000ull) __trap();  // the producer is gone: fail the launch, do not hang
    }
}

// run c of the gather kernel covers genes [lo1, hi1)
__device__ __forceinline__ void gather_run(const PipeGeom &gm, int c, int G, int *lo1, int *hi1) {
    const long long R = (long long)gm.tpc * WinTiling::tile_out;
    const long long s = (long long)c * R;
    *lo1 = c > 0 ? (int)(s - kHalo) : 0;
    const long long e = s + R;
    *hi1 = e < (long long)G ? (int)e : G;
}

// ---------------------------------------------------------------------------------------------------------
// K1: attribute ids -> unary odds
// ---------------------------------------------------------------------------------------------------------
template <int NSTAGE, typename PtrT>
__global__ void __launch_bounds__(kNT, 4)
unary_kernel(const WindowedArgs args, const PtrT *__restrict__ gene_ptr, const PipeGeom gm, const PipeBuffers pb_) {
    constexpr int NT = kNT;
    const CsrDev &csr = args.csr;
    const int tid = threadIdx.x;
    const uint32_t A = (uint32_t)args.model.A;
    const float clampv = args.model.clamp;
    const float fx_inv = __int_as_float((127 - args.model.fx_bits) << 23);  // 2^-fx_bits
    const int fx_nsafe = args.model.fx_nsafe;
    const int G = (int)csr.G;
    const int c = blockIdx.x;

    extern __shared__ __align__(16) float smem[];
    const int tab_words = round_up4s((int)A + 1);
    int *sTab = reinterpret_cast<int *>(smem);
    int32_t *sIdxBase = reinterpret_cast<int32_t *>(smem) + tab_words;          // NSTAGE buffers of kCap + 4
    int *sP = reinterpret_cast<int *>(smem) + tab_words + NSTAGE * (kCap + 4);  // kGatherTile + 4
    __shared__ __align__(8) uint64_t sBar[NSTAGE], sBarTab;

    int lo1, hi1;
    gather_run(gm, c, G, &lo1, &hi1);
    const int n1 = (hi1 - lo1 + kGatherTile - 1) / kGatherTile;
    const int run_start = c > 0 ? lo1 + kHalo : 0;  // genes below it go to the halo buffer
    const unsigned long long tag = (unsigned long long)gm.epoch << 32;

    asm volatile("griddepcontrol.launch_dependents;");
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&sBar[s], 1);
        mbar_init(&sBarTab, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t tab_bytes = (uint32_t)(4 * tab_words);
        mbar_expect_tx(&sBarTab, tab_bytes);
        tma_load_1d(sTab, args.model.table_fx, tab_bytes, &sBarTab);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");  // everything below reads the batch / writes the scratch

    // tile 0: ranges, row pointers; thread 0 starts the first NSTAGE copies
    int ga = lo1, gb = min(hi1, lo1 + kGatherTile);
    int64_t pa = (int64_t)__ldg(gene_ptr + ga), pb = (int64_t)__ldg(gene_ptr + gb);
    PtrT rowreg[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int t = tid + r * NT;
        rowreg[r] = t <= gb - ga ? __ldg(gene_ptr + ga + t) : 0;
    }
    __syncthreads();  // barriers initialised
    if (tid == 0) {
        stage_ids<kCap>(sIdxBase, csr.attr_idx, pa, pb, &sBar[0]);
        if (NSTAGE > 1 && n1 > 1) {
            const int g1a = lo1 + kGatherTile, g1b = min(hi1, g1a + kGatherTile);
            stage_ids<kCap>(sIdxBase + (kCap + 4), csr.attr_idx, (int64_t)__ldg(gene_ptr + g1a), (int64_t)__ldg(gene_ptr + g1b),
                            &sBar[1]);
        }
    }
    mbar_wait(&sBarTab, 0);  // delta table has landed (every thread observes the barrier itself)
    uint32_t parity = 0;     // bit s = parity of stage s

    for (int i = 0; i < n1; ++i) {
        const int s = NSTAGE > 1 ? (i % NSTAGE) : 0;
        int32_t *sIdx = sIdxBase + s * (kCap + 4);
        const int nn = gb - ga;
        const int64_t a0 = pa & ~(int64_t)3;
        const int64_t total64 = pb - a0;
        const bool staged = total64 <= kCap;  // CTA-uniform
        const int total = staged ? (int)total64 : 0;

        // ---- loads consumed later: the ranges of the tile whose copy is issued at the end of this one (thread 0),
        //      next tile's ranges and row pointers (everyone)
        int64_t fpa = 0, fpb = 0;
        const bool has_far = i + NSTAGE < n1;
        if (tid == 0 && has_far) {
            const int fga = lo1 + (i + NSTAGE) * kGatherTile, fgb = min(hi1, fga + kGatherTile);
            fpa = (int64_t)__ldg(gene_ptr + fga);
            fpb = (int64_t)__ldg(gene_ptr + fgb);
        }
        int nga = 0, ngb = 0;
        int64_t npa = 0, npb = 0;
        PtrT nrow[2] = {0, 0};
        if (i + 1 < n1) {
            nga = ga + kGatherTile;
            ngb = min(hi1, nga + kGatherTile);
            npa = (int64_t)__ldg(gene_ptr + nga);
            npb = (int64_t)__ldg(gene_ptr + ngb);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int t = tid + r * NT;
                nrow[r] = t <= ngb - nga ? __ldg(gene_ptr + nga + t) : 0;
            }
        }
        // ---- row pointers -> staged-range coordinates.  Row 0 starts at 0 so that the (at most 3) ids in front of
        //      the aligned start fold into it — they are masked to -1 below.
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int t = tid + r * NT;
            if (t <= nn) sP[t] = t == 0 ? 0 : (int)((int64_t)rowreg[r] - a0);
        }

        // ---- walk: staged ids -> running prefix of their fixed-point deltas, in place
        mbar_wait(&sBar[s], (parity >> s) & 1u);
        parity ^= 1u << s;
        if (tid == 0 && staged)
            for (int k = 0; k < (int)(pa - a0); ++k) sIdx[k] = -1;  // thread 0's own walk range: no barrier needed
        {
            const int x0 = tid * kWalk;
            int4 *v = reinterpret_cast<int4 *>(sIdx + x0);
            int run = 0;
            if (x0 + kWalk <= total) {
#pragma unroll
                for (int k = 0; k < kWalk / 4; ++k) {
                    int4 id = v[k];
                    run += lookup(sTab, id.x, A); id.x = run;
                    run += lookup(sTab, id.y, A); id.y = run;
                    run += lookup(sTab, id.z, A); id.z = run;
                    run += lookup(sTab, id.w, A); id.w = run;
                    v[k] = id;
                }
            } else {
#pragma unroll 1
                for (int k = 0; x0 + 4 * k < total; ++k) {
                    int4 id = v[k];
                    run += lookup(sTab, id.x, A); id.x = run;
                    run += lookup(sTab, id.y, A); id.y = run;
                    run += lookup(sTab, id.z, A); id.z = run;
                    run += lookup(sTab, id.w, A); id.w = run;
                    v[k] = id;
                }
            }
        }
        __syncthreads();  // prefixes and sP published

        // ---- row sums -> unary odds
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int t = tid + r * NT;
            if (t < nn) {
                const int b = sP[t], e = sP[t + 1];
                float u;
                if (staged && e - b < fx_nsafe) {
                    int v = 0;
                    if (e > b) {
                        const int q0 = walk_thread(b), q1 = walk_thread(e - 1);
                        v = sIdx[e - 1];
                        if (b != q0 * kWalk) v -= sIdx[b - 1];
                        if (q1 > q0) {
                            v += sIdx[(q0 + 1) * kWalk - 1];
#pragma unroll 1
                            for (int q = q0 + 2; q <= q1; ++q) v += sIdx[q * kWalk - 1];  // rows > 52 ids
                        }
                    }
                    u = exp_fast(fminf(fmaxf((float)v * fx_inv, -clampv), clampv));
                } else {
                    // a row long enough to wrap the int32 sum, or a tile whose ids exceed one staging round
                    u = direct_unary(gene_ptr, csr.attr_idx, args.model.table, A, ga + t, clampv);
                }
                const int g = ga + t;
                if (g >= run_start) pb_.u[g] = u;
                else pb_.halo[c * kHalo + (g - lo1)] = u;
            }
        }
        __syncthreads();  // the stage and sP are free again; every thread's stores precede thread 0's release
        if (tid == 0) {
            __threadfence();
            st_release(pb_.progress + c, tag | (unsigned long long)(i + 1));
            if (has_far) stage_ids<kCap>(sIdx, csr.attr_idx, fpa, fpb, &sBar[s]);
        }
        ga = nga; gb = ngb; pa = npa; pb = npb;
        rowreg[0] = nrow[0];
        rowreg[1] = nrow[1];
    }
}

// ---------------------------------------------------------------------------------------------------------
// K2: unary odds -> windowed marginals, max-pooled
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNT, 4)
window_kernel(const WindowedArgs args, const PipeGeom gm, const PipeBuffers pb_) {
    using T = WinTiling;
    constexpr int W = kW, NT = kNT, kPitch = T::kPitch;
    const CsrDev &csr = args.csr;
    const int tid = threadIdx.x;
    const float m01 = args.model.m01, m10 = args.model.m10, m11 = args.model.m11;
    const int step = args.step;
    const int G = (int)csr.G;
    const int c = blockIdx.x;

    extern __shared__ __align__(16) float smem[];
    float *sPool = smem + T::off_pool;
    float *sU0 = smem + T::off_u0;  // sU0[j] = u of local gene j
    float *sU1 = smem + T::off_u1;  // sU1[j] = u of local gene j + 1 (so odd pairs are 8-byte aligned too)
    float *sQ = smem + T::off_q;    // odds of genes of padded short contigs
    int *sCp = reinterpret_cast<int *>(smem + T::off_cp);  // contig_ptr slice in local gene coordinates
    unsigned char *sStat = reinterpret_cast<unsigned char *>(smem + T::off_stat);
    __shared__ int64_t sCursor;
    __shared__ int sShort;

    const int tile_begin = c * gm.tpc;
    const int tile_end = min(gm.num_tiles, tile_begin + gm.tpc);
    asm volatile("griddepcontrol.launch_dependents;");  // the next call's gather kernel may start its prologue
    if (tile_begin >= tile_end) return;

    int lo1, hi1;
    gather_run(gm, c, G, &lo1, &hi1);
    const unsigned long long tag = (unsigned long long)gm.epoch << 32;
    unsigned long long known = 0;     // thread 0: last progress value seen for this run
    bool next_known = false;          // thread 0: run c+1 has finished its first tile

    // the batch may still be in flight until the producer has passed griddepcontrol.wait: first flag, then cursor
    if (tid == 0) known = wait_progress(pb_.progress + c, tag | 1ull);
    __syncthreads();
    if (tid < 32) {
        const int64_t g0 = max(0, tile_begin * T::tile_out - T::lo);
        const int64_t cc = warp_find_contig(csr.contig_ptr, csr.C, g0, tid);
        if (tid == 0) sCursor = cc;
    }
    __syncthreads();

    for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int Gs = tile * T::tile_out - T::lo;  // global index of local gene 0 (negative only for tile 0)
        const int nout = min(G - (Gs + T::lo), T::tile_out);
        const int jlo = max(0, -Gs);           // first existing local gene
        const int jhi = min(T::ng, G - Gs);    // one past the last existing local gene
        const bool has_next = tile + 1 < tile_end;

        if (tid == 0) sShort = 0;
        const int64_t c_first = sCursor;
        int cp0 = INT_MAX;
        if (c_first + tid <= csr.C) cp0 = __ldg(csr.contig_ptr + c_first + tid) - Gs;
        // ---- acquire the producer's progress for the genes this tile reads
        if (tid == 0) {
            const int hi_need = Gs + jhi;
            const int in_run = min(hi_need, hi1);
            const unsigned long long need = tag | (unsigned long long)((in_run - lo1 + kGatherTile - 1) / kGatherTile);
            if (known < need) known = wait_progress(pb_.progress + c, need);
            if (hi_need > hi1 && !next_known) {
                wait_progress(pb_.progress + c + 1, tag | 1ull);
                next_known = true;
            }
        }
        sCp[tid] = cp0;
        if (tid == NT - 1) sCp[NT] = INT_MAX;  // sentinel unless the rest of the slice gets loaded below
        // a tile that holds more contigs than one slice entry per thread covers (contigs of 1-2 genes): load the rest
        const bool wide_slice = __syncthreads_or(tid == NT - 1 && cp0 < T::ng) != 0;  // also publishes the acquire
        // ---- unary odds of the tile's genes (L2 only: another kernel is still writing neighbouring lines)
        {
            const bool from_halo = tile == tile_begin && c > 0;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int j = tid + r * NT;
                if (j < T::ng + 1) {
                    float u = 1.0f;  // local genes that do not exist (before gene 0 / after gene G-1) are neutral
                    if (j >= jlo && j < jhi) {
                        const float *src = (from_halo && j < kHalo) ? pb_.halo + c * kHalo + j : pb_.u + (Gs + j);
                        u = __ldcg(src);
                    }
                    sU0[j] = u;
                    if (j >= 1) sU1[j - 1] = u;
                }
            }
        }
        if (wide_slice) {
            for (int k = NT + tid; k <= T::ng + 1; k += NT)
                sCp[k] = c_first + k <= csr.C ? __ldg(csr.contig_ptr + c_first + k) - Gs : INT_MAX;
            __syncthreads();
        }
        // ---- does the tile hold a contig shorter than the window (CTA-uniform; rare on long contigs), which
        //      contig holds the next tile's first staged gene, how many contigs start inside the tile
        for (int k = tid; k <= (wide_slice ? T::ng : NT - 1); k += NT) {
            const int a = sCp[k], b = sCp[k + 1];
            if (a < jhi && b > jlo && b != INT_MAX && b - a < W) sShort = 1;
            if (has_next && a <= T::tile_out && T::tile_out < b) sCursor = c_first + k;
        }
        // contig starts are a prefix of the slice: counting them among the first NT entries is exact unless all
        // of those are starts, in which case the searches below stay unbounded.  The barrier also publishes
        // sU0/sU1.
        int kt = __syncthreads_count(tid >= 1 && sCp[tid] < T::ng);
        if (kt >= NT - 1) kt = T::ng;
        const bool has_short = sShort != 0;
        if (has_short) {
            // per staged gene: status (1 = padded short contig, 2 = skipped short contig) and the padded windows
#pragma unroll 1
            for (int j = tid; j < T::ng; j += NT) {
                unsigned char stat = 0;
                if (j >= jlo && j < jhi) {
                    const int k = find_slice_contig(sCp, j, kt);
                    const int c0 = sCp[k], n = sCp[k + 1] - c0;
                    if (n < W) {
                        stat = args.pad ? 1 : 2;  // pad = 0: :228-234, the contig's genes keep "no probability"
                        // owner of the padded window: the contig's first gene — always staged when one of its
                        // genes is an output gene (c0 >= j - (W-2) >= 2 for j >= lo)
                        if (args.pad && j == c0 && j < T::lo + nout) padded_window<W>(sU0, sQ, j, n, m01, m10, m11);
                    }
                }
                sStat[j] = stat;
            }
            __syncthreads();
        }

        // ---- two adjacent windows per thread, packed f32x2
        {
            const int b0 = 2 * tid;
            // contig of slot b0 (or of the first existing gene, for the slots in front of gene 0)
            float va = 0.f, vb = 0.f;
            if (b0 + 1 >= jlo && b0 < jhi) {
                const int js = max(b0, jlo);
                int k = 0;
                if (kt <= 4) {
                    // few contigs start inside the tile (the usual case): count the starts at or before js
#pragma unroll
                    for (int q = 1; q <= 4; ++q) k += (q <= kt && sCp[q] <= js) ? 1 : 0;
                } else {
                    k = find_slice_contig(sCp, js, kt);
                }
                int c0 = sCp[k], c1 = sCp[k + 1];
                if (b0 >= jlo) va = (c1 - c0 >= W && b0 <= c1 - W && (step == 1 || (b0 - c0) % step == 0)) ? 1.f : 0.f;
                const int b1 = b0 + 1;
                if (b1 >= c1) {  // the second slot opens the next contig
                    c0 = c1;
                    c1 = sCp[k + 2];
                }
                if (b1 < jhi) vb = (c1 - c0 >= W && b1 <= c1 - W && (step == 1 || (b1 - c0) % step == 0)) ? 1.f : 0.f;
            }
            if (va + vb > 0.f) {
                auto upair = [&](int k) -> float2 {
                    return (k & 1) ? *reinterpret_cast<const float2 *>(&sU1[b0 + k - 1])
                                   : *reinterpret_cast<const float2 *>(&sU0[b0 + k]);
                };
                const float2 M01 = make_float2(m01 * va, m01 * vb);  // masked: an invalid slot keeps odds == 0
                const float2 M10 = make_float2(m10, m10), M11 = make_float2(m11, m11), ONE = make_float2(1.f, 1.f);
                const float2 B01 = make_float2(m01, m01);
                // The forward chain R_k (odds of alpha) and the backward chain S_k (odds of beta) are independent:
                // run them side by side, first halves stored, second halves combined with the stored other half.
                constexpr int H = W / 2;  // positions [0, H) meet positions [H, W)
                float2 ra[H], sb[H];      // ra[k] = R_k for k < H;  sb[i] = S_{H+i}
                auto fwd = [&](float2 R, int k) -> float2 {
                    const float2 num = __ffma2_rn(R, M11, M01);
                    const float2 den = __ffma2_rn(R, M10, ONE);
                    const float2 inv = make_float2(rcp_fast(den.x), rcp_fast(den.y));
                    return __fmul2_rn(__fmul2_rn(num, upair(k)), inv);
                };
                auto bwd = [&](float2 S, int k) -> float2 {  // S_{k+1} -> S_k
                    const float2 Wv = __fmul2_rn(upair(k + 1), S);
                    const float2 num = __ffma2_rn(Wv, M11, M10);
                    const float2 den = __ffma2_rn(Wv, B01, ONE);
                    const float2 inv = make_float2(rcp_fast(den.x), rcp_fast(den.y));
                    return __fmul2_rn(num, inv);
                };
                float2 R = __fmul2_rn(upair(0), make_float2(va, vb));
                float2 S = ONE;
                ra[0] = R;
                sb[H - 1] = S;
#pragma unroll
                for (int k = 1; k < H; ++k) {
                    R = fwd(R, k);
                    ra[k] = R;
                    S = bwd(S, W - 1 - k);
                    sb[H - 1 - k] = S;
                }
                // second halves: Q_k = R_k S_k upward from H, downward from H-1; m[j] = max(q_a[j], q_b[j-1])
                R = fwd(R, H);
                S = bwd(S, H - 1);
                float2 Qup = __fmul2_rn(R, sb[0]);       // Q_H
                float2 Qdn = __fmul2_rn(ra[H - 1], S);   // Q_{H-1}
                sPool[H * kPitch + tid] = fmaxf(Qup.x, Qdn.y);
#pragma unroll
                for (int q = 1; q < H; ++q) {
                    R = fwd(R, H + q);
                    const float2 Qu = __fmul2_rn(R, sb[q]);           // Q_{H+q}
                    sPool[(H + q) * kPitch + tid] = fmaxf(Qu.x, Qup.y);
                    Qup = Qu;
                    S = bwd(S, H - 1 - q);
                    const float2 Qd = __fmul2_rn(ra[H - 1 - q], S);   // Q_{H-1-q}
                    sPool[(H - q) * kPitch + tid] = fmaxf(Qdn.x, Qd.y);
                    Qdn = Qd;
                }
                sPool[W * kPitch + tid] = Qup.y;  // m[W] = q_b[W-1]
                sPool[tid] = Qdn.x;               // m[0] = q_a[0]
            } else {
#pragma unroll
                for (int k = 0; k <= W; ++k) sPool[k * kPitch + tid] = 0.f;
            }
        }
        __syncthreads();

        // ---- two output genes per thread: max over the covering windows, odds -> probability
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
            const int g = T::lo + tid + rep * NT;  // local gene
            if (g < T::lo + nout) {
                const int stat = has_short ? (int)sStat[g] : 0;
                float q = 0.f;
                if (stat == 0) {
                    // rows j = par, par+2, ... of column (g-j)/2: a constant stride of 2*pitch-1 words
                    const int par = g & 1;
                    const float *col = sPool + par * kPitch + ((g - par) >> 1);
#pragma unroll
                    for (int k = 0; 2 * k < W; ++k) q = fmaxf(q, col[k * (2 * kPitch - 1)]);
                    if (!par) q = fmaxf(q, col[(W / 2) * (2 * kPitch - 1)]);
                } else if (stat == 1) {
                    q = sQ[g];
                }
                float p = q * rcp_fast(1.0f + q);
                if (stat == 2) p = __int_as_float(0x7fc00000);
                const int gg = Gs + g;
                if (args.out_f32) static_cast<float *>(args.out)[gg] = p;
                else static_cast<double *>(args.out)[gg] = (double)p;
            }
        }
    }
}

size_t gather_smem_bytes(int A, int nstage) {
    return sizeof(float) * (size_t)(round_up4s(A + 1) + nstage * (kCap + 4) + round_up4s(kGatherTile + 4));
}

int env_int(const char *name, int fallback, int lo, int hi) {
    const char *v = getenv(name);
    if (!v || !*v) return fallback;
    const int x = atoi(v);
    return x < lo ? lo : (x > hi ? hi : x);
}

}  // namespace

bool pipeline_supported(const WindowedArgs &args) {
    if (args.window != kW) return false;
    if (gather_smem_bytes(args.model.A, 1) > 100 * 1024) return false;
    return args.csr.G < 0x7fff0000;  // tile arithmetic is 32-bit
}

void pipeline_geometry(int64_t G, int num_sms, int *grid, int *tpc, int64_t *num_tiles) {
    const int k = env_int("GCRF_PIPE_CTAS", 2, 1, 8);  // CTAs of each kernel per SM
    const int64_t nt = (G + WinTiling::tile_out - 1) / WinTiling::tile_out;
    int64_t g = (int64_t)num_sms * k;
    if (g > nt) g = nt;
    if (g < 1) g = 1;
    const int64_t per = (nt + g - 1) / g;
    *tpc = (int)per;
    *grid = (int)((nt + per - 1) / per);
    *num_tiles = nt;
}

size_t pipeline_scratch_bytes(int64_t G, int num_sms) {
    // progress[grid_max] | halo[grid_max][kHalo] | u[G]; grid <= 8 * num_sms.  The flags sit at a fixed offset so
    // that no later batch geometry ever reinterprets stale odds as flags.
    const size_t grid_max = (size_t)num_sms * 8;
    return grid_max * 8 + grid_max * kHalo * 4 + (size_t)G * 4 + 256;
}

cudaError_t launch_pipeline(const WindowedArgs &args, void *scratch, uint32_t epoch, int num_sms, cudaStream_t stream,
                            int64_t *launches) {
    if (args.csr.G <= 0) return cudaSuccess;
    PipeGeom gm{};
    int64_t nt = 0;
    pipeline_geometry(args.csr.G, num_sms, &gm.grid, &gm.tpc, &nt);
    gm.num_tiles = (int)nt;
    gm.epoch = epoch;
    const size_t grid_max = (size_t)num_sms * 8;
    PipeBuffers pb{};
    char *base = static_cast<char *>(scratch);
    pb.progress = reinterpret_cast<unsigned long long *>(base);
    base += grid_max * 8;
    pb.halo = reinterpret_cast<float *>(base);
    base += grid_max * kHalo * 4;
    pb.u = reinterpret_cast<float *>(base);

    const int nstage = env_int("GCRF_PIPE_STAGES", 2, 1, 2);
    const size_t smem1 = gather_smem_bytes(args.model.A, nstage);
    const size_t smem2 = sizeof(float) * (size_t)WinTiling::words;
    const bool p64 = args.csr.gene_ptr64 != nullptr;

    // kernel attributes depend on (device, A, stages, pointer width) only: set once per thread
    struct Cached { int device = -1, A = -1, nstage = -1, p64 = -1; };
    static thread_local Cached cache;
    int device = 0;
    cudaGetDevice(&device);
    if (cache.device != device || cache.A != args.model.A || cache.nstage != nstage || cache.p64 != (int)p64) {
        cudaError_t err;
        if (nstage == 2)
            err = p64 ? cudaFuncSetAttribute(unary_kernel<2, int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1)
                      : cudaFuncSetAttribute(unary_kernel<2, int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        else
            err = p64 ? cudaFuncSetAttribute(unary_kernel<1, int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1)
                      : cudaFuncSetAttribute(unary_kernel<1, int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (err != cudaSuccess) return err;
        cache.device = device; cache.A = args.model.A; cache.nstage = nstage; cache.p64 = (int)p64;
    }

    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // pairs with griddepcontrol.* in the kernels
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(gm.grid);
    cfg.blockDim = dim3(kNT);
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.dynamicSmemBytes = smem1;
    cudaError_t err;
    if (nstage == 2)
        err = p64 ? cudaLaunchKernelEx(&cfg, unary_kernel<2, int64_t>, args, args.csr.gene_ptr64, gm, pb)
                  : cudaLaunchKernelEx(&cfg, unary_kernel<2, int32_t>, args, args.csr.gene_ptr32, gm, pb);
    else
        err = p64 ? cudaLaunchKernelEx(&cfg, unary_kernel<1, int64_t>, args, args.csr.gene_ptr64, gm, pb)
                  : cudaLaunchKernelEx(&cfg, unary_kernel<1, int32_t>, args, args.csr.gene_ptr32, gm, pb);
    if (err != cudaSuccess) return err;
    if (launches) *launches += 1;
    // GCRF_PIPE_SERIAL=1: plain stream order between the two kernels (A/B of the overlap; results identical)
    cfg.numAttrs = env_int("GCRF_PIPE_SERIAL", 0, 0, 1) ? 0 : 1;
    cfg.dynamicSmemBytes = smem2;
    err = cudaLaunchKernelEx(&cfg, window_kernel, args, gm, pb);
    if (err != cudaSuccess) return err;
    if (launches) *launches += 1;
    return cudaSuccess;
}

}  // namespace gcrf
