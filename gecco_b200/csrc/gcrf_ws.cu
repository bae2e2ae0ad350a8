// gcrf_ws.cu — the W = 20 windowed path with WARP-SPECIALISED roles inside one persistent CTA.
//
// Same contract and arithmetic as gcrf_stream.cu (reference loop gecco/crf/__init__.py:209-258, tagger arithmetic
// SURVEY.md Appendix B); what changes is who does what.  In the fused kernel the four warps of a CTA run
// gather -> row sums -> DP -> pool one after the other, and a tile's latency chain (~10k cycles) is hidden only by
// the three other CTAs of the SM: the shared-memory pipe, the resource the gather saturates, sits idle a third of
// the time.  Here a CTA has two warpgroups:
//
//   gather warpgroup (threads 0-127, ~40 registers after setmaxnreg.dec)
//       TMA-staged attribute ids -> prefix walk through the fixed-point delta table -> row sums -> unary odds of
//       the tile's 275 genes into one of two shared-memory buffers, `full[buf]` mbarrier arrives.
//   window warpgroup (threads 128-255, ~120 registers after setmaxnreg.inc)
//       waits on `full[buf]`, does the contig bookkeeping, runs the packed two-window forward/backward odds
//       recursion from registers, max-pools through shared memory, writes the marginals, arrives on `empty[buf]`.
//
// The two roles work on consecutive tiles at the same time and synchronise only through two mbarrier pairs and
// their own named barriers (128 threads each), so gather work (shared-memory pipe) and DP work (FMA / MUFU pipes)
// overlap by construction; the register file is split 40 / 120 instead of 124 / 124.
//
// One CTA per SM holds THREE such pairs (768 threads, 12 + 12 warps), each streaming its own run of tiles.  Sharing
// the CTA lets the three pairs share one copy of the delta table, which is what pays for double-buffered id
// staging: the gather role has nothing to hide a ~1.5 us bulk copy behind, so tile n+2's ids are requested as soon
// as tile n's buffer is free.
#include "gcrf_device.cuh"

#include <climits>
#include <cstdlib>

namespace gcrf {

namespace {

constexpr int kW = 20;
constexpr int kRole = 128;  // threads per role
constexpr int kFew = 16;    // ids per lane of the one-warp walk used for tiles with <= 512 staged ids
constexpr int kPairs = 3;   // gather/window pairs per CTA

struct WsTiling {
    static constexpr int kSlots = 2 * kRole;
    static constexpr int kCap = kRole * kWalk;
    static constexpr int kPitch = kRole + 16;
    static constexpr int lo = kW;
    static constexpr int tile_out = (kSlots - kW) & ~3;  // 236
    static constexpr int ng = kSlots + kW - 1;           // 275
    static constexpr int keep = ng - tile_out;           // 39
    static constexpr int u_words = round_up4s(ng + 2);
    static constexpr int idx_words = kCap + 4;
    // per pair, relative to the pair's base
    int off_idx, off_pool, off_u, off_q, off_sp, off_cp, off_stat, pair_words;
    int tab_words, words;
    __host__ __device__ explicit WsTiling(int A) {
        tab_words = round_up4s(A + 1);
        int o = 0;
        off_idx = o; o += 2 * idx_words;  // two id stages
        off_pool = o; o += round_up4s((kW + 1) * kPitch);
        off_u = o; o += 2 * u_words;
        off_q = o; o += round_up4s(ng + 2);
        off_sp = o; o += round_up4s(tile_out + 3);
        off_cp = o; o += round_up4s(ng + 4);
        off_stat = o; o += round_up4s((ng + 8) / 4);
        pair_words = o;
        words = tab_words + kPairs * pair_words;
    }
    __host__ __device__ size_t bytes() const { return sizeof(float) * (size_t)words; }
};

__device__ __forceinline__ void role_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kRole) : "memory"); }
__device__ __forceinline__ bool role_or(int id, bool pred) {
    int r;
    asm volatile(
        "{\n.reg .pred p, q;\nsetp.ne.u32 p, %2, 0;\nbar.red.or.pred q, %1, %3, p;\nselp.u32 %0, 1, 0, q;\n}\n"
        : "=r"(r)
        : "r"(id), "r"((int)pred), "n"(kRole)
        : "memory");
    return r != 0;
}
__device__ __forceinline__ int role_count(int id, bool pred) {
    int r;
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\nbar.red.popc.u32 %0, %1, %3, p;\n}\n"
                 : "=r"(r)
                 : "r"(id), "r"((int)pred), "n"(kRole)
                 : "memory");
    return r;
}

// mbarrier wait that cannot hang the device: a phase that does not complete within ~2 s traps
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 26); ++spin) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}

#ifdef GCRF_TUNING
#define WS_MARK(slot)                                      \
    do {                                                   \
        if (prof_on && t == 0) {                           \
            const long long now__ = clock64();             \
            prof_acc[slot] += now__ - prof_last;           \
            prof_last = now__;                             \
        }                                                  \
    } while (0)
#define WS_FLUSH(base, n)                                                                                  \
    do {                                                                                                   \
        if (prof_on && t == 0) {                                                                           \
            for (int k__ = 0; k__ < (n); ++k__) atomicAdd(args.prof + (base) + k__, (unsigned long long)prof_acc[k__]); \
            atomicAdd(args.prof + 15, 1ull);                                                               \
        }                                                                                                  \
    } while (0)
#else
#define WS_MARK(slot) do { } while (0)
#define WS_FLUSH(base, n) do { } while (0)
#endif

template <typename PtrT>
__global__ void __launch_bounds__(2 * kRole * kPairs, 1)
ws_kernel(const WindowedArgs args, const PtrT *__restrict__ gene_ptr, const int num_tiles, const int tiles_per_pair) {
    using T = WsTiling;
    constexpr int W = kW, NT = kRole, kCap = T::kCap, kPitch = T::kPitch;
    const CsrDev &csr = args.csr;
    const T tl(args.model.A);
    const int tid = threadIdx.x;
    const int pair = tid >> 8;         // which gather/window pair of the CTA
    const int role = (tid >> 7) & 1;   // 0 = gather, 1 = window
    const int t = tid & (NT - 1);
    const int bar_id = 1 + 2 * pair + role;  // the role's own named barrier
    const uint32_t A = (uint32_t)args.model.A;
    const int G = (int)csr.G;

    extern __shared__ __align__(16) float smem[];
    int *sTab = reinterpret_cast<int *>(smem);
    float *pbase = smem + tl.tab_words + pair * tl.pair_words;
    int32_t *sIdxBase = reinterpret_cast<int32_t *>(pbase + tl.off_idx);  // two stages of idx_words
    float *sPool = pbase + tl.off_pool;
    float *sUbuf = pbase + tl.off_u;  // two buffers of u_words: unary odds of local genes 0 .. ng
    float *sQ = pbase + tl.off_q;
    int *sP = reinterpret_cast<int *>(pbase + tl.off_sp);
    int *sCp = reinterpret_cast<int *>(pbase + tl.off_cp);
    unsigned char *sStat = reinterpret_cast<unsigned char *>(pbase + tl.off_stat);
    __shared__ __align__(8) uint64_t sBarTab, sBarIdsAll[kPairs][2], sFullAll[kPairs][2], sEmptyAll[kPairs][2];
    __shared__ int64_t sCursorAll[kPairs];
    __shared__ int64_t sNextPbAll[kPairs];
    __shared__ int sShortAll[kPairs];
    uint64_t *sBarIds = sBarIdsAll[pair], *sFull = sFullAll[pair], *sEmpty = sEmptyAll[pair];
    int64_t &sCursor = sCursorAll[pair];
    int64_t &sNextPb = sNextPbAll[pair];
    int &sShort = sShortAll[pair];

    if ((int64_t)blockIdx.x * kPairs * tiles_per_pair >= num_tiles) return;  // CTA-uniform
    const int tile_begin = min(num_tiles, (blockIdx.x * kPairs + pair) * tiles_per_pair);
    const int tile_end = min(num_tiles, tile_begin + tiles_per_pair);
    const bool active = tile_begin < tile_end;  // a pair without tiles still takes part in the CTA-wide barriers

    // ---- prologue (all 256 threads) ---------------------------------------------------------------------------
    asm volatile("griddepcontrol.launch_dependents;");
    if (tid == 0) {
        mbar_init(&sBarTab, 1);
        for (int p = 0; p < kPairs; ++p)
            for (int b = 0; b < 2; ++b) {
                mbar_init(&sBarIdsAll[p][b], 1);
                mbar_init(&sFullAll[p][b], 1);
                mbar_init(&sEmptyAll[p][b], NT);
            }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t tab_bytes = (uint32_t)(4 * round_up4s((int)A + 1));
        mbar_expect_tx(&sBarTab, tab_bytes);
        tma_load_1d(sTab, args.model.table_fx, tab_bytes, &sBarTab);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");  // everything below reads the batch
    const int Gs0 = tile_begin * T::tile_out - T::lo;
    if (active && role == 1 && t >= NT - 32) {  // the window warpgroup's last warp: contig of the run's first staged gene
        const int64_t g0 = max(0, Gs0);
        const int64_t c = warp_find_contig(csr.contig_ptr, csr.C, g0 + csr.gene_base, tid & 31);
        if ((tid & 31) == 0) sCursor = c;
    }
    // the run's first halo [0, keep): gathered by the gather warpgroup with shared-memory integer atomics into the
    // slots the first tile's ring carry reads (buffer 1, since the first tile fills buffer 0)
    {
        int *hrow = sP;
        int *hacc = reinterpret_cast<int *>(sPool);
        const int h0 = max(0, min(G, Gs0)), h1 = max(0, min(G, Gs0 + T::keep));
        int64_t hp0 = 0;
        if (role == 0 && active) {
            hp0 = (int64_t)__ldg(gene_ptr + h0);
            if (t <= h1 - h0) hrow[t] = (int)((int64_t)__ldg(gene_ptr + h0 + t) - hp0);
            if (t < T::keep) hacc[t] = 0;
        }
        __syncthreads();  // mbarriers initialised, row pointers, zeroed accumulators
        if (role == 0 && active) {
            mbar_wait(&sBarTab, 0);
            const int hn = h1 - h0, hids = hn > 0 ? hrow[hn] : 0;
            for (int x = t; x < hids; x += NT) {
                int row = 0, hi = hn;
                while (hi - row > 1) {
                    const int mid = (row + hi) >> 1;
                    if (hrow[mid] <= x) row = mid; else hi = mid;
                }
                atomicAdd(&hacc[row], lookup(sTab, __ldg(csr.attr_idx + hp0 + x), A));
            }
        }
        __syncthreads();
        if (role == 0 && active && t < T::keep) {
            const float clampv = args.model.clamp;
            const float fx_inv = __int_as_float((127 - args.model.fx_bits) << 23);
            const int g = Gs0 + t;
            float u = 1.0f;
            if (g >= h0 && g < h1) {
                const int r = g - h0;
                u = hrow[r + 1] - hrow[r] < args.model.fx_nsafe
                        ? exp_fast(fminf(fmaxf((float)hacc[r] * fx_inv, -clampv), clampv))
                        : direct_unary_inl(gene_ptr, csr.attr_idx, args.model.table, A, g, clampv);
            }
            sUbuf[T::u_words + T::tile_out + t] = u;
        }
        __syncthreads();  // halo, cursor; sP and sPool are free again
    }

    if (role == 0) {
        // =============================== gather warpgroup ===========================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;\n");
        if (!active) return;
#ifdef GCRF_TUNING
        const bool prof_on = args.prof != nullptr;
        long long prof_acc[6] = {0, 0, 0, 0, 0, 0};
        long long prof_last = prof_on ? clock64() : 0;
#endif
        const float clampv = args.model.clamp;
        const float fx_inv = __int_as_float((127 - args.model.fx_bits) << 23);  // 2^-fx_bits
        const int fx_nsafe = args.model.fx_nsafe;
        int ga = max(0, min(G, Gs0 + T::keep)), gb = max(0, min(G, Gs0 + T::ng));
        int64_t pa = (int64_t)__ldg(gene_ptr + ga), pb = (int64_t)__ldg(gene_ptr + gb);
        PtrT rowreg[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int k = t + r * NT;
            rowreg[r] = k <= gb - ga ? __ldg(gene_ptr + ga + k) : 0;
        }
        // Two id stages.  Thread 0 runs the copies two tiles ahead: tile n+2's range end is loaded (one scalar) at
        // the top of tile n and its copy issued when tile n's stage is free, a whole tile before it is waited for.
        auto range_end = [&](int tile_index) -> PtrT {  // end of tile_index's new-id range
            const int g = max(0, min(G, tile_index * T::tile_out - T::lo + T::ng));
            return __ldg(gene_ptr + g);
        };
        PtrT issued_end = (PtrT)pb;  // thread 0: end of the last range whose copy has been issued
        if (t == 0) {
            stage_ids<kCap>(sIdxBase, csr.attr_idx, pa, pb, &sBarIds[0]);
            if (tile_begin + 1 < tile_end) {
                const PtrT e1 = range_end(tile_begin + 1);
                stage_ids<kCap>(sIdxBase + T::idx_words, csr.attr_idx, pb, (int64_t)e1, &sBarIds[1]);
                issued_end = e1;
            }
        }

        for (int tile = tile_begin; tile < tile_end; ++tile) {
            const int it = tile - tile_begin;
            const int buf = it & 1;
            int32_t *sIdx = sIdxBase + buf * T::idx_words;
            float *sU = sUbuf + buf * T::u_words;
            const float *sUprev = sUbuf + (buf ^ 1) * T::u_words;
            const bool has_far = tile + 2 < tile_end;
            PtrT far_end = 0;
            if (t == 0 && has_far) far_end = range_end(tile + 2);
            const int Gs = tile * T::tile_out - T::lo;
            const int nn = gb - ga;
            const int jn0 = ga - Gs;
            const int jlo = max(0, -Gs);
            const int jhi = min(T::ng, G - Gs);
            if (it > 0) {
                pa = pb;
                pb = sNextPb;
            }
            const int64_t a0 = pa & ~(int64_t)3;
            const int64_t total64 = pb - a0;
            const bool staged = total64 <= kCap;
            const int total = staged ? (int)total64 : 0;
            const bool has_next = tile + 1 < tile_end;
            const bool few_ids = staged && total <= 32 * kFew;

            // row pointers of the new genes -> staged-range coordinates (row 0 starts at 0: the ids in front of the
            // aligned start fold into it and are masked to -1 below)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int k = t + r * NT;
                if (k <= nn) sP[k] = k == 0 ? 0 : (int)((int64_t)rowreg[r] - a0);
            }
            // next tile's row pointers (consumed one tile later)
            int nga = 0, ngb = 0;
            PtrT nrow[2] = {0, 0};
            if (has_next) {
                nga = max(0, min(G, Gs + T::tile_out + T::keep));
                ngb = max(0, min(G, Gs + T::tile_out + T::ng));
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int k = t + r * NT;
                    nrow[r] = k <= ngb - nga ? __ldg(gene_ptr + nga + k) : 0;
                }
            }

            // ---- walk: staged ids -> prefix sums of their fixed-point deltas, in place
            WS_MARK(0);  // top of tile: pointers, prefetch
            mbar_wait_bounded(&sBarIds[buf], (uint32_t)((it >> 1) & 1));
            WS_MARK(1);  // wait for the ids
            if (t == 0 && staged)
                for (int i = 0; i < (int)(pa - a0); ++i) sIdx[i] = -1;
            {
                int run = 0;
                if (few_ids) {
                    if (t < 32 && t * kFew < total) {
                        int4 *w = reinterpret_cast<int4 *>(sIdx + t * kFew);
                        int4 id[kFew / 4];
#pragma unroll
                        for (int i = 0; i < kFew / 4; ++i) id[i] = w[i];
#pragma unroll
                        for (int i = 0; i < kFew / 4; ++i) {
                            id[i].x = lookup(sTab, id[i].x, A);
                            id[i].y = lookup(sTab, id[i].y, A);
                            id[i].z = lookup(sTab, id[i].z, A);
                            id[i].w = lookup(sTab, id[i].w, A);
                        }
#pragma unroll
                        for (int i = 0; i < kFew / 4; ++i) {
                            id[i].x += run;
                            id[i].y += id[i].x;
                            id[i].z += id[i].y;
                            id[i].w += id[i].z;
                            run = id[i].w;
                        }
                        int inc = run;
                        const unsigned active = __activemask();
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const int y = __shfl_up_sync(active, inc, d);
                            if (t >= d) inc += y;
                        }
                        const int base = inc - run;
#pragma unroll
                        for (int i = 0; i < kFew / 4; ++i) {
                            id[i].x += base; id[i].y += base; id[i].z += base; id[i].w += base;
                            w[i] = id[i];
                        }
                    }
                } else if (t * kWalk < total) {
                    int4 *v = reinterpret_cast<int4 *>(sIdx + t * kWalk);
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
            role_sync(bar_id);  // prefixes and sP published
            WS_MARK(2);  // walk + barrier

            if (has_next) {
                const int last = ngb - nga;
                if (t == last % NT) sNextPb = (int64_t)(last >= NT ? nrow[1] : nrow[0]);
            }
            // the buffer's previous contents (tile it - 2) must have been consumed
            if (it >= 2) mbar_wait_bounded(&sEmpty[buf], (uint32_t)(((it >> 1) - 1) & 1));
            WS_MARK(3);  // wait for the window role to release the buffer
            // ring: the odds of the genes shared with the previous tile live in the other buffer
            if (t < T::keep) sU[t] = sUprev[t + T::tile_out];
            // ---- row sums -> unary odds of the new genes
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int k = t + r * NT;
                if (k < nn) {
                    const int s = sP[k], e = sP[k + 1];
                    float u;
                    if (staged && e - s < fx_nsafe) {
                        int v = 0;
                        if (few_ids) {
                            if (e > s) v = sIdx[e - 1] - (s > 0 ? sIdx[s - 1] : 0);
                        } else if (e > s) {
                            const int q0 = walk_thread(s), q1 = walk_thread(e - 1);
                            v = sIdx[e - 1];
                            if (s != q0 * kWalk) v -= sIdx[s - 1];
                            if (q1 > q0) {
                                v += sIdx[(q0 + 1) * kWalk - 1];
#pragma unroll 1
                                for (int q = q0 + 2; q <= q1; ++q) v += sIdx[q * kWalk - 1];
                            }
                        }
                        u = exp_fast(fminf(fmaxf((float)v * fx_inv, -clampv), clampv));
                    } else {
                        u = direct_unary_inl(gene_ptr, csr.attr_idx, args.model.table, A, ga + k, clampv);
                    }
                    sU[jn0 + k] = u;
                }
            }
            if (jlo > 0 || jhi < T::ng) {
                for (int j = t; j < T::ng + 1; j += NT)
                    if (j < jlo || j >= jhi) sU[j] = 1.0f;  // genes that do not exist are neutral
            }
            role_sync(bar_id);  // the buffer is complete, this id stage is free
            WS_MARK(4);  // row sums + barrier
            if (t == 0) {
                if (has_far) {
                    stage_ids<kCap>(sIdx, csr.attr_idx, (int64_t)issued_end, (int64_t)far_end, &sBarIds[buf]);
                    issued_end = far_end;
                }
                mbar_arrive(&sFull[buf]);
            }
            ga = nga; gb = ngb;
            rowreg[0] = nrow[0];
            rowreg[1] = nrow[1];
        }
        WS_FLUSH(0, 5);
    } else {
        // =============================== window warpgroup ===========================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;\n");
        if (!active) return;
#ifdef GCRF_TUNING
        const bool prof_on = args.prof != nullptr;
        long long prof_acc[6] = {0, 0, 0, 0, 0, 0};
        long long prof_last = prof_on ? clock64() : 0;
#endif
        const float m01 = args.model.m01, m10 = args.model.m10, m11 = args.model.m11;
        const int step = args.step;
        // the contig slice of a tile is loaded one tile ahead (as soon as the scan has found the next cursor), so
        // that its global-memory latency hides behind the DP
        int64_t c_first = sCursor;
        int cp_raw = 0;
        if (c_first + t <= csr.C) cp_raw = __ldg(csr.contig_ptr + c_first + t);

        for (int tile = tile_begin; tile < tile_end; ++tile) {
            const int it = tile - tile_begin;
            const int buf = it & 1;
            const float *sU0 = sUbuf + buf * T::u_words;
            const int Gs = tile * T::tile_out - T::lo;
            const int nout = min(G - (Gs + T::lo), T::tile_out);
            const int jlo = max(0, -Gs);
            const int jhi = min(T::ng, G - Gs);
            const bool has_next = tile + 1 < tile_end;

            if (t == 0) sShort = 0;
            const bool cp_ok = c_first + t <= csr.C;
            const int GsA = Gs + (int)csr.gene_base;
            const int cp0 = cp_ok ? cp_raw - GsA : INT_MAX;
            sCp[t] = cp0;
            if (t == NT - 1) sCp[NT] = INT_MAX;
            const bool wide_slice = role_or(bar_id, t == NT - 1 && cp0 < T::ng);
            if (wide_slice) {
                for (int k = NT + t; k <= T::ng + 1; k += NT)
                    sCp[k] = c_first + k <= csr.C ? __ldg(csr.contig_ptr + c_first + k) - GsA : INT_MAX;
                role_sync(bar_id);
            }
            for (int k = t; k <= (wide_slice ? T::ng : NT - 1); k += NT) {
                const int a = sCp[k], b = sCp[k + 1];
                if (a < jhi && b > jlo && b != INT_MAX && b - a < W) sShort = 1;
                if (has_next && a <= T::tile_out && T::tile_out < b) sCursor = c_first + k;
            }
            int kt = role_count(bar_id, t >= 1 && sCp[t] < T::ng);
            if (kt >= NT - 1) kt = T::ng;
            const bool has_short = sShort != 0;
            if (has_next) {
                c_first = sCursor;  // written by the scan above, published by the counting barrier
                cp_raw = c_first + t <= csr.C ? __ldg(csr.contig_ptr + c_first + t) : 0;
            }

            WS_MARK(0);  // contig bookkeeping
            // ---- the tile's unary odds
            mbar_wait_bounded(&sFull[buf], (uint32_t)((it >> 1) & 1));
            WS_MARK(1);  // wait for the gather role

            if (has_short) {
#pragma unroll 1
                for (int j = t; j < T::ng; j += NT) {
                    unsigned char stat = 0;
                    if (j >= jlo && j < jhi) {
                        const int k = find_slice_contig(sCp, j, kt);
                        const int c0 = sCp[k], n = sCp[k + 1] - c0;
                        if (n < W) {
                            stat = args.pad ? 1 : 2;
                            if (args.pad && j == c0 && j < T::lo + nout) padded_window_inl<W>(sU0, sQ, j, n, m01, m10, m11);
                        }
                    }
                    sStat[j] = stat;
                }
                role_sync(bar_id);
            }

            // ---- two adjacent windows per thread, packed f32x2
            {
                const int b0 = 2 * t;
                float va = 0.f, vb = 0.f;
                if (b0 + 1 >= jlo && b0 < jhi) {
                    const int js = max(b0, jlo);
                    int k = 0;
                    if (kt <= 4) {
#pragma unroll
                        for (int i = 1; i <= 4; ++i) k += (i <= kt && sCp[i] <= js) ? 1 : 0;
                    } else {
                        k = find_slice_contig(sCp, js, kt);
                    }
                    int c0 = sCp[k], c1 = sCp[k + 1];
                    if (b0 >= jlo) va = (c1 - c0 >= W && b0 <= c1 - W && (step == 1 || (b0 - c0) % step == 0)) ? 1.f : 0.f;
                    const int b1 = b0 + 1;
                    if (b1 >= c1) {
                        c0 = c1;
                        c1 = sCp[k + 2];
                    }
                    if (b1 < jhi) vb = (c1 - c0 >= W && b1 <= c1 - W && (step == 1 || (b1 - c0) % step == 0)) ? 1.f : 0.f;
                }
                if (va + vb > 0.f) {
                    float uu[W + 1];
#pragma unroll
                    for (int i = 0; i < W / 2; ++i) {
                        const float2 p = *reinterpret_cast<const float2 *>(&sU0[b0 + 2 * i]);
                        uu[2 * i] = p.x;
                        uu[2 * i + 1] = p.y;
                    }
                    uu[W] = sU0[b0 + W];
                    auto upair = [&](int k) -> float2 { return make_float2(uu[k], uu[k + 1]); };
                    const float2 M01 = make_float2(m01 * va, m01 * vb);
                    const float2 M10 = make_float2(m10, m10), M11 = make_float2(m11, m11), ONE = make_float2(1.f, 1.f);
                    const float2 B01 = make_float2(m01, m01);
                    constexpr int H = W / 2;
                    float2 ra[H], sb[H];
                    auto fwd = [&](float2 R, int k) -> float2 {
                        const float2 num = __ffma2_rn(R, M11, M01);
                        const float2 den = __ffma2_rn(R, M10, ONE);
                        const float2 inv = make_float2(rcp_fast(den.x), rcp_fast(den.y));
                        return __fmul2_rn(__fmul2_rn(num, upair(k)), inv);
                    };
                    auto bwd = [&](float2 S, int k) -> float2 {
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
                    R = fwd(R, H);
                    S = bwd(S, H - 1);
                    float2 Qup = __fmul2_rn(R, sb[0]);
                    float2 Qdn = __fmul2_rn(ra[H - 1], S);
                    sPool[H * kPitch + t] = fmaxf(Qup.x, Qdn.y);
#pragma unroll
                    for (int i = 1; i < H; ++i) {
                        R = fwd(R, H + i);
                        const float2 Qu = __fmul2_rn(R, sb[i]);
                        sPool[(H + i) * kPitch + t] = fmaxf(Qu.x, Qup.y);
                        Qup = Qu;
                        S = bwd(S, H - 1 - i);
                        const float2 Qd = __fmul2_rn(ra[H - 1 - i], S);
                        sPool[(H - i) * kPitch + t] = fmaxf(Qdn.x, Qd.y);
                        Qdn = Qd;
                    }
                    sPool[W * kPitch + t] = Qup.y;
                    sPool[t] = Qdn.x;
                } else {
#pragma unroll
                    for (int k = 0; k <= W; ++k) sPool[k * kPitch + t] = 0.f;
                }
            }
            mbar_arrive(&sEmpty[buf]);  // this thread is done with the buffer's odds
            role_sync(bar_id);
            WS_MARK(2);  // short contigs + DP + barrier

            // ---- two output genes per thread
#pragma unroll
            for (int rep = 0; rep < 2; ++rep) {
                const int g = T::lo + t + rep * NT;
                if (g < T::lo + nout) {
                    const int stat = has_short ? (int)sStat[g] : 0;
                    float q = 0.f;
                    if (stat == 0) {
                        const int par = g & 1;
                        const float *col = sPool + par * kPitch + ((g - par) >> 1);
#pragma unroll
                        for (int i = 0; 2 * i < W; ++i) q = fmaxf(q, col[i * (2 * kPitch - 1)]);
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
            WS_MARK(3);  // pool + output
        }
        WS_FLUSH(8, 4);
    }
}

template <typename PtrT>
cudaError_t configure_ws(int A, int *ctas_per_sm, size_t *bytes) {
    const WsTiling tl(A);
    *bytes = tl.bytes();
    auto kernel = ws_kernel<PtrT>;
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl.bytes());
    if (err != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, kernel, 2 * kRole * kPairs, tl.bytes());
}

}  // namespace

bool ws_supported(const WindowedArgs &args) {
    if (args.window != kW) return false;
    const WsTiling tl(args.model.A);
    if (tl.bytes() > 226 * 1024) return false;  // one CTA per SM: table + three pairs of buffers
    return args.csr.G < 0x7fff0000;  // tile arithmetic is 32-bit
}

cudaError_t plan_ws(const WindowedArgs &args, int num_sms, WindowedPlan *plan) {
    struct Cached { int device = -1, A = -1, p64 = -1, per_sm = 0; size_t bytes = 0; };
    static thread_local Cached cache;
    int device = 0;
    cudaGetDevice(&device);
    const bool p64 = args.csr.gene_ptr64 != nullptr;
    if (cache.device != device || cache.A != args.model.A || cache.p64 != (int)p64) {
        int q = 0;
        size_t b = 0;
        cudaError_t err = p64 ? configure_ws<int64_t>(args.model.A, &q, &b) : configure_ws<int32_t>(args.model.A, &q, &b);
        if (err != cudaSuccess) return err;
        cache.device = device; cache.A = args.model.A; cache.p64 = (int)p64; cache.per_sm = q; cache.bytes = b;
    }
    if (cache.per_sm < 1) return cudaErrorInvalidConfiguration;
    plan->threads = 2 * kRole * kPairs;
    plan->tile_out = WsTiling::tile_out;
    plan->chunk = WsTiling::kCap;
    plan->smem_bytes = cache.bytes;
    plan->num_tiles = (args.csr.G + plan->tile_out - 1) / plan->tile_out;
    plan->ctas_per_sm = cache.per_sm;
    // runs = gather/window pairs; tiles_per_cta holds the tiles per PAIR, a CTA covers kPairs consecutive runs
    int64_t runs = (int64_t)num_sms * cache.per_sm * kPairs;
    if (runs > plan->num_tiles) runs = plan->num_tiles;
    if (runs < 1) runs = 1;
    plan->tiles_per_cta = (int)((plan->num_tiles + runs - 1) / runs);
    const int64_t used_runs = (plan->num_tiles + plan->tiles_per_cta - 1) / plan->tiles_per_cta;
    plan->grid = (int)((used_runs + kPairs - 1) / kPairs);
    return cudaSuccess;
}

cudaError_t launch_ws(const WindowedArgs &args, const WindowedPlan &plan, cudaStream_t stream, int64_t *launches) {
    if (args.csr.G <= 0) return cudaSuccess;
    const int nt_ = (int)plan.num_tiles, tpc = plan.tiles_per_cta;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(plan.grid);
    cfg.blockDim = dim3(2 * kRole * kPairs);
    cfg.dynamicSmemBytes = plan.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t err;
    if (args.csr.gene_ptr64) err = cudaLaunchKernelEx(&cfg, ws_kernel<int64_t>, args, args.csr.gene_ptr64, nt_, tpc);
    else err = cudaLaunchKernelEx(&cfg, ws_kernel<int32_t>, args, args.csr.gene_ptr32, nt_, tpc);
    if (launches) *launches += 1;
    return err;
}

}  // namespace gcrf
