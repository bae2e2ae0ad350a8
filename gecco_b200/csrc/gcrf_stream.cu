// gcrf_stream.cu — the fast fused kernel for compile-time window sizes (W = 20 is GECCO's shipped model).
//
// Same contract as gcrf_windowed.cu (reference loop: gecco/crf/__init__.py:209-258, tagger arithmetic
// SURVEY.md Appendix B) but organised around what the ncu captures and ablation runs showed to bound it on
// B200: shared-memory wavefronts (table look-ups), per-thread bookkeeping instructions and dependent-latency
// chains at 4 warps per scheduler — not DRAM.
//
//   * One persistent CTA walks a contiguous run of tiles.  Tile = 2*NT window slots, `tile_out` output
//     genes.  The unary odds u_g of the 2W-1 genes shared with the previous tile stay in shared memory
//     (a ring; the run's very first halo is gathered in the prologue), so every attribute id is fetched
//     and resolved exactly once per CTA run.
//   * attr_idx of the tile's NEW genes is brought in by ONE bulk-async copy (TMA, cp.async.bulk +
//     mbarrier) issued while the previous tile's dynamic programme runs: no load instructions, no
//     registers, latency off the critical path (measured: < 5 % of CTA time spent waiting).
//   * Segmented sums without any per-id boundary logic (a per-thread branch there diverges in every
//     warp): thread t walks ids [K t, K t + K) of the staged range with 16-byte shared loads (K = 4*odd
//     keeps them bank-conflict free), resolves each id through a FIXED-POINT delta table and overwrites
//     it in place with the thread's running prefix sum.  A row sum is then a difference of prefixes
//     (plus whole-thread totals when a row straddles threads) in wrapping int32 arithmetic: exact, and
//     order-independent.  Rows with >= fx_nsafe ids could wrap and take a float path from global memory;
//     tiles whose ids do not fit one staging round take a direct row-per-thread path.
//   * Dynamic programme on odds ratios, two adjacent windows per thread packed in f32x2 registers
//     (FFMA2/FMUL2 halve the issue slots; MUFU.RCP stays scalar), forward and backward chains advancing
//     together.  Each thread finds the contig of its own two window slots.  Invalid window slots run with
//     m01 and u_0 masked to zero, which pins their odds to exactly 0 — the neutral element of the
//     max-pool — so the pool needs no range logic at all.
//   * Padded short contigs (:216-227) and skipped ones (pad = 0) are handled by a slow path that a tile
//     only enters when it actually holds a contig shorter than the window.
//   * Three CTA barriers per tile.
//
// Build with -DGCRF_TUNING to compile the phase timers (GCRF_PHASE_PROFILE=1) and the ablation switches
// (GCRF_DEBUG_SKIP) in; the production build carries neither.
#include "gcrf_device.cuh"

#include <climits>
#include <cstdlib>

namespace gcrf {

namespace {

// CTA geometry of the production kernel; -DGCRF_STREAM_NT / -DGCRF_STREAM_MINB build A/B variants (tools/build_variant.sh)
#ifndef GCRF_STREAM_NT
#define GCRF_STREAM_NT 128
#endif
#ifndef GCRF_STREAM_MINB
#define GCRF_STREAM_MINB 4
#endif
constexpr int kNT = GCRF_STREAM_NT, kMinB = GCRF_STREAM_MINB;
// Resident CTAs the register budget is sized for: from W = 25 on the pool ((W + 1) rows) pushes a CTA past a quarter of
// the SM's shared memory and three fit whatever the registers (two from W = 51 on) — so those windows get three (two)
// CTAs' worth of registers: 168 (255) instead of 128, no spills up to W = 64.  W = 40 on config 2: 121.5 -> 112.4 us.
template <int W>
constexpr int min_blocks() { return W <= 20 ? kMinB : W <= 50 ? (kMinB < 3 ? kMinB : 3) : (kMinB < 2 ? kMinB : 2); }
constexpr int kStreamSmemCap = 100 * 1024;  // largest dynamic shared memory a streaming CTA may ask for
constexpr int kFewPerLane = 16;  // ids per lane of the one-warp walk used for tiles with <= 512 staged ids

// SLOTS: window slots per tile.  2 * NT (two windows per DP thread) is the production geometry; the same kernel with
// NT or NT / 2 slots (the upper DP threads idle) halves / quarters the genes per tile while the id buffer stays — for
// batches so dense (> ~26 ids per gene on average) that a full tile's ids would not fit one staging round and every tile
// would fall to the row-by-row path (2.7x slower: tools/density_probe.py).
template <int W, int NT, int SLOTS = 2 * NT>
struct StreamTiling {
    static constexpr int kSlots = SLOTS;           // window slots per tile
    static constexpr int kCap = NT * kWalk;        // ids staged per tile
    static constexpr int kPitch = NT + 16;         // pool pitch: odd and even genes land 16 banks apart
    static constexpr int lo = W;                   // local index of the first output gene
    static constexpr int tile_out = (kSlots - W) & ~3;
    static constexpr int ng = kSlots + W - 1;      // genes staged per tile
    static constexpr int keep = ng - tile_out;     // genes carried over from the previous tile
    static_assert(tile_out + 1 <= 2 * NT, "two row pointers per thread must cover a tile's new genes");
    static_assert(keep <= NT, "the ring carry is one gene per thread");
    static_assert(keep + 1 <= NT, "one halo row pointer per thread (with NT >= keep + 33 the last warp holds none and only runs the contig search)");
    int off_idx, off_pool, off_u0, off_q, off_sp, off_cp, off_stat, words;
    __host__ __device__ explicit StreamTiling(int A) {
        int o = round_up4s(A + 1);
        off_idx = o; o += kCap + 4;
        off_pool = o; o += round_up4s((W + 1) * kPitch);
        off_u0 = o; o += round_up4s(ng + 2);
        off_q = o; o += round_up4s(ng + 2);
        off_sp = o; o += round_up4s(tile_out + 3);
        off_cp = o; o += round_up4s((ng > NT ? ng : NT) + 4);  // one slice entry per thread at least
        off_stat = o; o += round_up4s((ng + 8) / 4);
        words = o;
    }
    __host__ __device__ size_t bytes() const { return sizeof(float) * (size_t)words; }
};

#ifdef GCRF_TUNING
#define GCRF_MARK(slot)                                   \
    do {                                                  \
        if (prof_on && tid == 0) {                        \
            const long long now__ = clock64();            \
            prof_acc[slot] += now__ - prof_last;          \
            prof_last = now__;                            \
        }                                                 \
    } while (0)
#define GCRF_SKIP(bit) (args.debug_skip & (bit))
#else
#define GCRF_MARK(slot) do { } while (0)
#define GCRF_SKIP(bit) false
#endif

// PEERS: results also go to the peer output arrays of a contig-sharded batch (gcrf_marginals_windowed_peers); a
// separate instantiation, so that the single-GPU kernel carries none of it.
template <int W, int NT, int MINB, typename PtrT, int SLOTS = 2 * NT, bool PEERS = false>
__global__ void __launch_bounds__(NT, MINB)
stream_kernel(const WindowedArgs args, const PtrT *__restrict__ gene_ptr, const int num_tiles, const int tiles_per_cta) {
    using T = StreamTiling<W, NT, SLOTS>;
    constexpr int kCap = T::kCap, kPitch = T::kPitch;
    const CsrDev &csr = args.csr;
    const T tl(args.model.A);
    const int tid = threadIdx.x;
    const uint32_t A = (uint32_t)args.model.A;
    const float m01 = args.model.m01, m10 = args.model.m10, m11 = args.model.m11;
    const float clampv = args.model.clamp;
    const float fx_inv = __int_as_float((127 - args.model.fx_bits) << 23);  // 2^-fx_bits
    const int fx_nsafe = args.model.fx_nsafe;
    const int step = args.step;
    const int G = (int)csr.G;

    extern __shared__ __align__(16) float smem[];
    int *sTab = reinterpret_cast<int *>(smem);
    int32_t *sIdx = reinterpret_cast<int32_t *>(smem + tl.off_idx);
    float *sPool = smem + tl.off_pool;
    float *sU0 = smem + tl.off_u0;  // sU0[j] = u of local gene j
    float *sQ = smem + tl.off_q;    // odds of genes of padded short contigs
    int *sP = reinterpret_cast<int *>(smem + tl.off_sp);   // staged-range coordinates of the new genes' rows
    int *sCp = reinterpret_cast<int *>(smem + tl.off_cp);  // contig_ptr slice in local gene coordinates
    unsigned char *sStat = reinterpret_cast<unsigned char *>(smem + tl.off_stat);
    __shared__ __align__(8) uint64_t sBar, sBarTab;
    __shared__ int64_t sCursor;
    __shared__ int64_t sNextPb;  // end of the next tile's id range, published by the thread that holds that row pointer
    __shared__ int sShort;

    // Consecutive tiles per CTA, spread evenly: the first num_tiles % grid CTAs take one more than the others (with
    // everybody at the rounded-up count the last CTAs of the grid had nothing to do — 26 of config 2's 592 — and the
    // SMs they would have shared finished no sooner for it)
    (void)tiles_per_cta;
    const int tiles_base = num_tiles / (int)gridDim.x, tiles_rem = num_tiles % (int)gridDim.x;
    const int tile_begin = (int)blockIdx.x * tiles_base + min((int)blockIdx.x, tiles_rem);
    const int tile_end = tile_begin + tiles_base + ((int)blockIdx.x < tiles_rem ? 1 : 0);
    if (tile_begin >= tile_end) return;
#ifdef GCRF_TUNING
    const bool prof_on = args.prof != nullptr;
    long long prof_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_last = prof_on ? clock64() : 0;
#endif

    // ---- prologue -------------------------------------------------------------------------------------
    // Programmatic dependent launch: let the next kernel in the stream start its own prologue while this grid
    // drains, and do not touch the batch (which the previous kernel may still be producing) before it completed.
    asm volatile("griddepcontrol.launch_dependents;");
    if (tid == 0) {
        mbar_init(&sBar, 1);
        mbar_init(&sBarTab, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // delta table -> shared in one bulk copy (a per-thread copy loop costs ~20 dependent global round trips);
        // the device table is padded to a multiple of 16 bytes
        const uint32_t tab_bytes = (uint32_t)(4 * round_up4s((int)A + 1));
        mbar_expect_tx(&sBarTab, tab_bytes);
        tma_load_1d(sTab, args.model.table_fx, tab_bytes, &sBarTab);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");  // everything below reads the batch
    if (tid >= NT - 32) {
        // the last warp: it holds none of the halo's row pointers, so its three dependent global round trips run
        // beside the other warps' loads instead of in front of them
        const int64_t g0 = max(0, tile_begin * T::tile_out - T::lo);
        const int64_t c = warp_find_contig(csr.contig_ptr, csr.C, g0 + csr.gene_base, tid & 31);
        if (tid == NT - 32) sCursor = c;
    }
    // New genes of a tile = local genes [keep, ng).  The run's first halo [0, keep) is gathered right here into the
    // slots the first tile's ring carry will read: its ids are one contiguous range, every thread resolves a
    // strided share of it and adds the fixed-point deltas to its row's accumulator (integer atomics: exact).
    const int Gs0 = tile_begin * T::tile_out - T::lo;
    {
        int *hrow = sP;                                   // keep + 1 row pointers, relative to the first one
        int *hacc = reinterpret_cast<int *>(sPool);       // keep accumulators
        const int h0 = max(0, min(G, Gs0)), h1 = max(0, min(G, Gs0 + T::keep));  // existing halo genes [h0, h1)
        const int64_t hp0 = (int64_t)__ldg(gene_ptr + h0);
        if (tid <= h1 - h0) hrow[tid] = (int)((int64_t)__ldg(gene_ptr + h0 + tid) - hp0);
        if (tid < T::keep) hacc[tid] = 0;
        __syncthreads();  // mbarriers initialised, row pointers, zeroed accumulators
        mbar_wait(&sBarTab, 0);  // delta table has landed
        const int hn = h1 - h0, hids = hn > 0 ? hrow[hn] : 0;
        for (int x = tid; x < hids; x += NT) {
            int row = 0, hi = hn;  // largest row with hrow[row] <= x
            while (hi - row > 1) {
                const int mid = (row + hi) >> 1;
                if (hrow[mid] <= x) row = mid; else hi = mid;
            }
            atomicAdd(&hacc[row], lookup(sTab, __ldg(csr.attr_idx + hp0 + x), A));
        }
        __syncthreads();
        if (tid < T::keep) {
            const int g = Gs0 + tid;
            float u = 1.0f;  // genes before gene 0 / after gene G-1 are neutral
            if (g >= h0 && g < h1) {
                const int r = g - h0;
                u = hrow[r + 1] - hrow[r] < fx_nsafe
                        ? exp_fast(fminf(fmaxf((float)hacc[r] * fx_inv, -clampv), clampv))
                        : direct_unary(gene_ptr, csr.attr_idx, args.model.table, A, g, clampv);
            }
            sU0[T::tile_out + tid] = u;
        }
    }
    int ga = max(0, min(G, Gs0 + T::keep)), gb = max(0, min(G, Gs0 + T::ng));
    // Only the first tile loads its id range [pa, pb) directly.  Later tiles start where the previous one ended
    // (pa' = pb) and take pb' from the row pointers that are prefetched a tile ahead anyway: the thread holding the
    // last one publishes it through shared memory after the walk barrier.  A direct warp-uniform load of pb' gets
    // moved to a uniform register by the compiler right away, which stalled every warp for a full global-memory
    // latency at the top of each tile (ncu: 7-8 % of all stall samples).
    int64_t pa = (int64_t)__ldg(gene_ptr + ga), pb = (int64_t)__ldg(gene_ptr + gb);
    PtrT rowreg[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int t = tid + r * NT;
        rowreg[r] = t <= gb - ga ? __ldg(gene_ptr + ga + t) : 0;
    }
    __syncthreads();  // mbarrier initialised, delta table staged, cursor and halo written
    if (tid == 0) stage_ids<kCap>(sIdx, csr.attr_idx, pa, pb, &sBar);
    uint32_t bar_parity = 0;

    for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int Gs = tile * T::tile_out - T::lo;  // global index of local gene 0 (negative only for tile 0)
        const int nout = min(G - (Gs + T::lo), T::tile_out);
        const int nn = gb - ga;     // new genes this tile
        const int jn0 = ga - Gs;    // local index of the first new gene (= keep, except at the batch edges)
        const int jlo = max(0, -Gs);           // first existing local gene
        const int jhi = min(T::ng, G - Gs);    // one past the last existing local gene
        if (tile > tile_begin) {
            pa = pb;
            pb = sNextPb;  // written before the previous tile's unary barrier
        }
        const int64_t a0 = pa & ~(int64_t)3;
        const int64_t total64 = pb - a0;       // staged-range length (ids) from the aligned start
        const bool staged = total64 <= kCap;   // CTA-uniform: the usual case
        const int total = staged ? (int)total64 : 0;
        const bool has_next = tile + 1 < tile_end;
        const bool few_ids = staged && total <= 32 * kFewPerLane;  // CTA-uniform

        GCRF_MARK(8);
        // Three CTA barriers per tile: after the walk, after the unary odds (a counting barrier) and after the
        // DP.  Everything written before the first one (sP, the ring carry, sShort, sCp) was last read before
        // the previous tile's DP barrier or is read only by its writer.
        if (tid == 0) sShort = 0;
        // ---- A. row pointers of the new genes -> staged-range coordinates.  Row 0 starts at 0 so that the
        //         (at most 3) ids in front of the aligned start fold into it — they are masked to -1 below.
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int t = tid + r * NT;
            if (t <= nn) sP[t] = t == 0 ? 0 : (int)((int64_t)rowreg[r] - a0);
        }
        // ---- B. loads consumed later in the tile: contig slice, next tile's ranges and row pointers
        const int64_t c_first = sCursor;
        const bool cp_ok = c_first + tid <= csr.C && !GCRF_SKIP(64);
        int cp_raw = 0;  // rebased to local gene coordinates only after the walk
        if (cp_ok) cp_raw = __ldg(csr.contig_ptr + c_first + tid);
        int nga = 0, ngb = 0;
        PtrT nrow[2] = {0, 0};
        if (has_next && !GCRF_SKIP(128)) {
            nga = max(0, min(G, Gs + T::tile_out + T::keep));
            ngb = max(0, min(G, Gs + T::tile_out + T::ng));
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int t = tid + r * NT;
                nrow[r] = t <= ngb - nga ? __ldg(gene_ptr + nga + t) : 0;
            }
        }
        // ring: carry the odds of the genes shared with the previous tile; source [tile_out, tile_out+keep) and
        // destination [0, keep) do not overlap
        if (tid < T::keep) {
            const float c0 = sU0[tid + T::tile_out];
            sU0[tid] = c0;
        }
        GCRF_MARK(0);

        // ---- C. gather: staged ids -> per-row fixed-point sums
        if (!GCRF_SKIP(32)) mbar_wait(&sBar, bar_parity);
        bar_parity ^= 1;
        // ids in front of the first new row sit in thread 0's own walk range: no barrier needed
        if (tid == 0 && staged)
            for (int i = 0; i < (int)(pa - a0); ++i) sIdx[i] = -1;
        GCRF_MARK(1);
        if (!GCRF_SKIP(1)) {
            // walk: ids -> running prefix of their fixed-point deltas, in place
            const int x0 = tid * kWalk;
            int4 *v = reinterpret_cast<int4 *>(sIdx + x0);
            int run = 0;
            // A thread whose range is only partly staged walks all of it as well: the words beyond `total` are
            // leftovers of earlier tiles (any bit pattern resolves to a valid table slot) and no row ends there.
            // A separate rolled loop for that one thread was the slowest path into the barrier below.
            // (Splitting this into three register passes — 13 id loads, 52 look-ups, prefix — so that all shared-memory
            // requests are in flight at once moved the time into the other phases: 79.9 vs 78.1 us on config 2.  The
            // shared-memory pipe, not this thread's latency, is what the walk runs against.)
            if (few_ids) {
                // Real annotation density (1-2 domains per gene): a few hundred ids per tile.  Thirteen dependent
                // load -> look-up rounds in a handful of threads would be the tile's longest latency chain; instead
                // the first warp takes 16 ids per lane, all loads and look-ups in flight together, and a shuffle scan
                // turns the per-lane sums into ONE prefix over the whole staged range (no per-thread segments).
                if (tid < 32 && tid * kFewPerLane < total) {
                    int4 *w = reinterpret_cast<int4 *>(sIdx + tid * kFewPerLane);
                    int4 id[kFewPerLane / 4];
#pragma unroll
                    for (int i = 0; i < kFewPerLane / 4; ++i) id[i] = w[i];
#pragma unroll
                    for (int i = 0; i < kFewPerLane / 4; ++i) {
                        id[i].x = lookup(sTab, id[i].x, A);
                        id[i].y = lookup(sTab, id[i].y, A);
                        id[i].z = lookup(sTab, id[i].z, A);
                        id[i].w = lookup(sTab, id[i].w, A);
                    }
#pragma unroll
                    for (int i = 0; i < kFewPerLane / 4; ++i) {
                        id[i].x += run;
                        id[i].y += id[i].x;
                        id[i].z += id[i].y;
                        id[i].w += id[i].z;
                        run = id[i].w;
                    }
                    int inc = run;  // lanes without ids were filtered above: shuffle only among the active ones
                    const unsigned active = __activemask();
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int y = __shfl_up_sync(active, inc, d);
                        if (tid >= d) inc += y;
                    }
                    const int base = inc - run;
#pragma unroll
                    for (int i = 0; i < kFewPerLane / 4; ++i) {
                        id[i].x += base; id[i].y += base; id[i].z += base; id[i].w += base;
                        w[i] = id[i];
                    }
                }
            } else if (x0 < total) {
#pragma unroll
                for (int i = 0; i < kWalk / 4; ++i) {
                    int4 id = v[i];
                    run += lookup(sTab, id.x, A); id.x = run;
                    run += lookup(sTab, id.y, A); id.y = run;
                    run += lookup(sTab, id.z, A); id.z = run;
                    run += lookup(sTab, id.w, A); id.w = run;
                    v[i] = id;
                }
            }
        }
        // contig slice -> shared (its load was issued at the top of the tile; visible after the barrier)
        const int GsA = Gs + (int)csr.gene_base;  // local gene 0 in the units of contig_ptr's values
        const int cp0 = cp_ok ? cp_raw - GsA : INT_MAX;
        sCp[tid] = cp0;
        if (tid == NT - 1) sCp[NT] = INT_MAX;  // sentinel unless the rest of the slice gets loaded below
        // a tile that holds more contigs than one slice entry per thread covers (contigs of 1-2 genes): load the rest
        const bool wide_slice = __syncthreads_or(tid == NT - 1 && cp0 < T::ng) != 0;
        // the next tile's last row pointer (loaded at the top of this tile, long since arrived) -> shared; read by
        // thread 0 after the unary barrier below and by everyone at the top of the next tile
        if (has_next) {
            const int last = ngb - nga;
            if (tid == last % NT) sNextPb = (int64_t)(last >= NT ? nrow[1] : nrow[0]);
        }
        if (wide_slice) {
            for (int k = NT + tid; k <= T::ng + 1; k += NT)
                sCp[k] = c_first + k <= csr.C ? __ldg(csr.contig_ptr + c_first + k) - GsA : INT_MAX;
            __syncthreads();
        }
        GCRF_MARK(2);

        // ---- D. row sums -> unary odds of the new genes
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int t = tid + r * NT;
            if (t < nn && !GCRF_SKIP(8)) {
                const int s = sP[t], e = sP[t + 1];
                float u;
                if (staged && e - s < fx_nsafe) {
                    // prefix difference, plus the totals of the threads the row runs through
                    int v = 0;
                    if (few_ids) {
                        if (e > s) v = sIdx[e - 1] - (s > 0 ? sIdx[s - 1] : 0);  // one prefix over the whole range
                    } else if (e > s) {
                        const int q0 = walk_thread(s), q1 = walk_thread(e - 1);
                        v = sIdx[e - 1];
                        if (s != q0 * kWalk) v -= sIdx[s - 1];
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
                const int j = jn0 + t;
                sU0[j] = u;
            }
        }
        if (jlo > 0 || jhi < T::ng) {
            // local genes that do not exist (before gene 0 / after gene G-1) are neutral: u = 1
            for (int j = tid; j < T::ng + 1; j += NT) {
                if (j < jlo || j >= jhi) {
                    sU0[j] = 1.0f;
                }
            }
        }
        GCRF_MARK(3);

        // ---- does the tile hold a contig shorter than the window (CTA-uniform; rare on long contigs), which
        //      contig holds the next tile's first staged gene, how many contigs start inside the tile
        for (int k = tid; k <= (wide_slice ? T::ng : NT - 1); k += NT) {
            const int a = sCp[k], b = sCp[k + 1];
            if (a < jhi && b > jlo && b != INT_MAX && b - a < W) sShort = 1;
            if (has_next && a <= T::tile_out && T::tile_out < b) sCursor = c_first + k;
        }
        // contig starts are a prefix of the slice: counting them among the first NT entries is exact unless all
        // of those are starts, in which case the searches below stay unbounded.  The barrier also publishes
        // sU0 and retires the last readers of sIdx.
        int kt = __syncthreads_count(tid >= 1 && sCp[tid] < T::ng);
        if (kt >= NT - 1) kt = T::ng;
        const bool has_short = sShort != 0;
        // ---- stage the next tile's ids while this tile's dynamic programme runs
        if (has_next && tid == 0 && !GCRF_SKIP(32)) stage_ids<kCap>(sIdx, csr.attr_idx, pb, sNextPb, &sBar);
        GCRF_MARK(4);
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
        GCRF_MARK(5);

        // ---- E. two adjacent windows per thread, packed f32x2
        {
            const int b0 = 2 * tid;
            // contig of slot b0 (or of the first existing gene, for the slots in front of gene 0)
            float va = 0.f, vb = 0.f;
            if (b0 + 1 >= jlo && b0 < jhi && b0 < T::kSlots) {
                const int js = max(b0, jlo);
                int k = 0;
                if (kt <= 4) {
                    // few contigs start inside the tile (the usual case): count the starts at or before js
#pragma unroll
                    for (int i = 1; i <= 4; ++i) k += (i <= kt && sCp[i] <= js) ? 1 : 0;
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
            if (GCRF_SKIP(2)) va = vb = 0.f;
            if (va + vb > 0.f) {
                // The 21 unary odds the two windows touch, loaded ONCE (ten 8-byte loads and one 4-byte load) and kept
                // in registers for both chains: per-step operand loads were 320 of a tile's ~1,740 shared-memory
                // wavefronts.  Odd steps pair registers of two different loads — two scalar multiplies, no load.
                float uu[W + 2];
#pragma unroll
                for (int i = 0; i < (W + 1) / 2; ++i) {
                    const float2 p = *reinterpret_cast<const float2 *>(&sU0[b0 + 2 * i]);
                    uu[2 * i] = p.x;
                    uu[2 * i + 1] = p.y;
                }
                if ((W & 1) == 0) uu[W] = sU0[b0 + W];
                auto upair = [&](int k) -> float2 { return make_float2(uu[k], uu[k + 1]); };
                const float2 M01 = make_float2(m01 * va, m01 * vb);  // masked: an invalid slot keeps odds == 0
                const float2 M10 = make_float2(m10, m10), M11 = make_float2(m11, m11), ONE = make_float2(1.f, 1.f);
                const float2 B01 = make_float2(m01, m01);
                // The forward chain R_k (odds of alpha) and the backward chain S_k (odds of beta) are independent:
                // run them side by side, first halves stored, second halves combined with the stored other half.
                constexpr int H = W / 2;         // the chains meet between positions H-1 and H (even W) or at H (odd W)
                constexpr int ODD = W & 1;
                static_assert(W >= 2, "packed DP needs two positions");
                float2 ra[H], sb[H];      // ra[k] = R_k for k < H;  sb[i] = S_{H+ODD+i}
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
                // second halves: Q_k = R_k S_k upward and downward from the meeting point; m[j] = max(q_a[j], q_b[j-1])
                float2 Qup, Qdn;
                if (ODD) {
                    R = fwd(R, H);
                    S = bwd(S, H);
                    Qup = Qdn = __fmul2_rn(R, S);  // Q_H, the middle position: both directions start from it
                } else {
                    R = fwd(R, H);
                    S = bwd(S, H - 1);
                    Qup = __fmul2_rn(R, sb[0]);       // Q_H
                    Qdn = __fmul2_rn(ra[H - 1], S);   // Q_{H-1}
                    sPool[H * kPitch + tid] = fmaxf(Qup.x, Qdn.y);
                }
#pragma unroll
                for (int i = 1 - ODD; i < H; ++i) {
                    R = fwd(R, H + ODD + i);
                    const float2 Qu = __fmul2_rn(R, sb[i]);                 // Q_{H+ODD+i}
                    sPool[(H + ODD + i) * kPitch + tid] = fmaxf(Qu.x, Qup.y);
                    Qup = Qu;
                    S = bwd(S, H - 1 - i);
                    const float2 Qd = __fmul2_rn(ra[H - 1 - i], S);         // Q_{H-1-i}
                    sPool[(H - i) * kPitch + tid] = fmaxf(Qdn.x, Qd.y);
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
        GCRF_MARK(6);

        // ---- F. two output genes per thread: max over the covering windows, odds -> probability
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
            const int g = T::lo + tid + rep * NT;  // local gene
            if (g < T::lo + nout && !GCRF_SKIP(4)) {
                const int stat = has_short ? (int)sStat[g] : 0;
                float q = 0.f;
                if (stat == 0) {
                    // rows j = par, par+2, ... of column (g-j)/2: a constant stride of 2*pitch-1 words
                    const int par = g & 1;
                    const float *col = sPool + par * kPitch + ((g - par) >> 1);
#pragma unroll
                    for (int i = 0; 2 * i < W; ++i) q = fmaxf(q, col[i * (2 * kPitch - 1)]);
                    if (!par || (W & 1)) q = fmaxf(q, col[((W - par) / 2) * (2 * kPitch - 1)]);
                } else if (stat == 1) {
                    q = sQ[g];
                }
                // the approximate reciprocal can land one ulp above q/(1+q): a probability must not exceed 1
                float p = fminf(q * rcp_fast(1.0f + q), 1.0f);
                if (stat == 2) p = __int_as_float(0x7fc00000);
                if constexpr (PEERS) {
                    store_result(args, Gs + g, p);
                } else {
                    if (args.out_f32) static_cast<float *>(args.out)[Gs + g] = p;
                    else static_cast<double *>(args.out)[Gs + g] = (double)p;
                }
            }
        }
        GCRF_MARK(7);

        // rotate the prefetched ranges
        ga = nga; gb = ngb;
        rowreg[0] = nrow[0];
        rowreg[1] = nrow[1];
    }
#ifdef GCRF_TUNING
    if (prof_on && tid == 0) {
#pragma unroll
        for (int k = 0; k < 10; ++k) atomicAdd(args.prof + k, (unsigned long long)prof_acc[k]);
        atomicAdd(args.prof + 15, 1ull);
    }
#endif
}

template <int W, int NT, int MINB, typename PtrT, int SLOTS>
cudaError_t configure_stream(int A, int *ctas_per_sm, size_t *bytes) {
    const StreamTiling<W, NT, SLOTS> tl(A);
    *bytes = tl.bytes();
    auto kernel = stream_kernel<W, NT, MINB, PtrT, SLOTS>;
    // The attribute is per kernel and device, not per launch: always raise it to the cap stream_supported() enforces, so
    // that host threads with models of different sizes cannot lower it under each other's cached plans.
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmemCap);
    if (err != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, kernel, NT, tl.bytes());
}


// Window sizes with a compiled streaming kernel: GECCO's shipped model (20), the default of `gecco train` (5,
// gecco/cli/commands/_parser.py:364-372) and a spread of sizes around them (the reference takes any window_size >= 1,
// gecco/crf/__init__.py:134-137); every other size runs the generic kernel (or, past 128, the f64 path).
#define GCRF_STREAM_WINDOWS(X) X(5) X(10) X(15) X(20) X(25) X(30) X(40) X(50) X(64)

// Windows that also get the half- and quarter-tile variants for dense batches: the shipped model's and `gecco train`'s.
#define GCRF_STREAM_DENSE_WINDOWS(X) X(5) X(20)
template <int W> constexpr bool has_dense_variants() {
    bool yes = false;
#define X(V) yes = yes || W == V;
    GCRF_STREAM_DENSE_WINDOWS(X)
#undef X
    return yes;
}

template <int W, int SLOTS>
cudaError_t configure_slots(const WindowedArgs &args, int *per_sm, size_t *bytes, int *tile_out) {
    *tile_out = StreamTiling<W, kNT, SLOTS>::tile_out;
    return args.csr.gene_ptr64 ? configure_stream<W, kNT, min_blocks<W>(), int64_t, SLOTS>(args.model.A, per_sm, bytes)
                               : configure_stream<W, kNT, min_blocks<W>(), int32_t, SLOTS>(args.model.A, per_sm, bytes);
}
template <int W>
cudaError_t configure_window(const WindowedArgs &args, int slots, int *per_sm, size_t *bytes, int *tile_out) {
    if constexpr (has_dense_variants<W>()) {
        if (slots == kNT) return configure_slots<W, kNT>(args, per_sm, bytes, tile_out);
        if (slots == kNT / 2) return configure_slots<W, kNT / 2>(args, per_sm, bytes, tile_out);
    }
    return configure_slots<W, 2 * kNT>(args, per_sm, bytes, tile_out);
}

template <int W, int SLOTS>
cudaError_t launch_slots(const cudaLaunchConfig_t &cfg, const WindowedArgs &args, int num_tiles, int tiles_per_cta) {
    return args.csr.gene_ptr64
               ? cudaLaunchKernelEx(&cfg, stream_kernel<W, kNT, min_blocks<W>(), int64_t, SLOTS>, args, args.csr.gene_ptr64, num_tiles, tiles_per_cta)
               : cudaLaunchKernelEx(&cfg, stream_kernel<W, kNT, min_blocks<W>(), int32_t, SLOTS>, args, args.csr.gene_ptr32, num_tiles, tiles_per_cta);
}
// the peer-store instantiation (full tiles only; the windows that have dense variants have this one too)
template <int W, typename PtrT>
cudaError_t launch_peers(const cudaLaunchConfig_t &cfg, const WindowedArgs &args, const PtrT *gene_ptr, int num_tiles, int tiles_per_cta) {
    auto kernel = stream_kernel<W, kNT, min_blocks<W>(), PtrT, 2 * kNT, true>;
    static thread_local int configured_device = -1;
    int device = 0;
    cudaGetDevice(&device);
    if (configured_device != device) {
        const cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmemCap);
        if (err != cudaSuccess) return err;
        configured_device = device;
    }
    return cudaLaunchKernelEx(&cfg, kernel, args, gene_ptr, num_tiles, tiles_per_cta);
}

template <int W>
cudaError_t launch_window(const cudaLaunchConfig_t &cfg, const WindowedArgs &args, int slots, int num_tiles, int tiles_per_cta) {
    if constexpr (has_dense_variants<W>()) {
        if (args.n_peer_out > 0)
            return args.csr.gene_ptr64 ? launch_peers<W>(cfg, args, args.csr.gene_ptr64, num_tiles, tiles_per_cta)
                                       : launch_peers<W>(cfg, args, args.csr.gene_ptr32, num_tiles, tiles_per_cta);
        if (slots == kNT) return launch_slots<W, kNT>(cfg, args, num_tiles, tiles_per_cta);
        if (slots == kNT / 2) return launch_slots<W, kNT / 2>(cfg, args, num_tiles, tiles_per_cta);
    }
    return launch_slots<W, 2 * kNT>(cfg, args, num_tiles, tiles_per_cta);
}

// Window slots per tile for a batch of this density: the most that keep an average tile's ids (plus 4 % headroom for
// the spread between tiles) inside one staging round.
template <int W>
int slots_for_density(double ids_per_gene) {
    if constexpr (has_dense_variants<W>()) {
        constexpr int kCap = StreamTiling<W, kNT>::kCap;
        if (StreamTiling<W, kNT, 2 * kNT>::tile_out * ids_per_gene * 1.04 <= kCap) return 2 * kNT;
        if (StreamTiling<W, kNT, kNT>::tile_out * ids_per_gene * 1.04 <= kCap) return kNT;
        return kNT / 2;
    }
    return 2 * kNT;
}

}  // namespace

bool stream_supported(const WindowedArgs &args) {
    if (args.n_peer_out > 0) {  // peer output arrays: the windows with a peer-store instantiation
        bool ok = false;
#define X(W) ok = ok || args.window == W;
        GCRF_STREAM_DENSE_WINDOWS(X)
#undef X
        if (!ok) return false;
    }
    size_t bytes = 0;
    switch (args.window) {
#define X(W) case W: bytes = StreamTiling<W, kNT>(args.model.A).bytes(); break;
        GCRF_STREAM_WINDOWS(X)
#undef X
        default: return false;
    }
    if (bytes > (size_t)kStreamSmemCap) return false;
    // tile arithmetic is 32-bit: G + one tile of slack must fit
    return args.csr.G < 0x7fff0000;
}

cudaError_t plan_stream(const WindowedArgs &args, int num_sms, WindowedPlan *plan) {
    // the kernel attribute / occupancy query depend on (device, A, pointer width, window) only: cache them per thread
    struct Cached { int device = -1, A = -1, p64 = -1, window = -1, slots = -1, per_sm = 0, tile_out = 0; size_t bytes = 0; };
    static thread_local Cached cache;
    int device = 0;
    cudaGetDevice(&device);
    const bool p64 = args.csr.gene_ptr64 != nullptr;
    const double density = args.csr.G > 0 ? (double)(args.csr.slice_ids > 0 ? args.csr.slice_ids : args.csr.nnz) / (double)args.csr.G : 0.0;
    int slots = 2 * kNT;
    switch (args.window) {
#define X(W) case W: slots = slots_for_density<W>(density); break;
        GCRF_STREAM_WINDOWS(X)
#undef X
    }
    if (args.n_peer_out > 0) slots = 2 * kNT;  // the peer-store kernel exists for full tiles only
    if (const char *env = getenv("GCRF_STREAM_SLOTS")) {  // A/B: 256, 128 or 64 window slots per tile
        const int want = atoi(env);
        if (want == 2 * kNT || want == kNT || want == kNT / 2) slots = want;
    }
    if (cache.device != device || cache.A != args.model.A || cache.p64 != (int)p64 || cache.window != args.window || cache.slots != slots) {
        int q = 0, tile_out = 0;
        size_t b = 0;
        cudaError_t err = cudaErrorInvalidValue;
        switch (args.window) {
#define X(W) case W: err = configure_window<W>(args, slots, &q, &b, &tile_out); break;
            GCRF_STREAM_WINDOWS(X)
#undef X
        }
        if (err != cudaSuccess) return err;
        cache.device = device; cache.A = args.model.A; cache.p64 = (int)p64; cache.window = args.window; cache.slots = slots;
        cache.per_sm = q; cache.bytes = b; cache.tile_out = tile_out;
    }
    int per_sm = cache.per_sm;
    const size_t bytes = cache.bytes;
#ifdef GCRF_TUNING
    if (const char *cap = getenv("GCRF_CTAS_PER_SM")) {  // occupancy experiments: fewer resident CTAs than fit
        const int want = atoi(cap);
        if (want >= 1 && want < per_sm) per_sm = want;
    }
#endif
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    plan->threads = kNT;
    plan->slots = slots;
    plan->tile_out = cache.tile_out;
    plan->chunk = kNT * kWalk;
    plan->smem_bytes = bytes;
    plan->num_tiles = (args.csr.G + plan->tile_out - 1) / plan->tile_out;
    plan->ctas_per_sm = per_sm;
    int64_t grid = (int64_t)num_sms * per_sm;
    if (grid > plan->num_tiles) grid = plan->num_tiles;
    if (grid < 1) grid = 1;
    plan->tiles_per_cta = (int)((plan->num_tiles + grid - 1) / grid);  // the most any CTA gets (see the kernel)
    plan->grid = (int)grid;
    return cudaSuccess;
}

cudaError_t launch_stream(const WindowedArgs &args, const WindowedPlan &plan, cudaStream_t stream, int64_t *launches) {
    if (args.csr.G <= 0) return cudaSuccess;
    const int nt_ = (int)plan.num_tiles, tpc = plan.tiles_per_cta;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(plan.grid);
    cfg.blockDim = dim3(kNT);
    cfg.dynamicSmemBytes = plan.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // pairs with griddepcontrol.* in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t err = cudaErrorInvalidValue;
    switch (args.window) {
#define X(W) case W: err = launch_window<W>(cfg, args, plan.slots, nt_, tpc); break;
        GCRF_STREAM_WINDOWS(X)
#undef X
    }
    if (launches) *launches += 1;
    return err;
}

}  // namespace gcrf
