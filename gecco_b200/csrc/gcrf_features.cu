// gcrf_features.cu — feature extraction on device: integer domain accessions -> attribute ids of the CRF model,
// with the set semantics of gecco/crf/features.py:13-35 (a gene's features are a dict keyed by domain name, so a
// domain that occurs twice in a gene counts once) and the tagger's handling of unknown attributes (dropped).
// Row pointers are left untouched: dropped entries become -1, which the marginal kernels ignore.
#include "gcrf_kernels.cuh"

#include <cstdlib>

namespace gcrf {

namespace {

constexpr int kThreads = 256;

// Exact for any vocabulary size, and the path for blocks the bitmap kernel cannot take: one warp per gene, lane i resolves
// row i (+32, +64, ...) and looks for an earlier equal accession among the rows of its gene (features.py:32 — the first
// occurrence keeps the key).
__device__ __forceinline__ void gene_rows_simple(const int32_t *__restrict__ accession, int64_t rb, int64_t re,
                                                 const int32_t *__restrict__ lut, int32_t lut_size,
                                                 int32_t *__restrict__ out, int lane) {
    for (int64_t p = rb + lane; p < re; p += 32) {
        const int32_t acc = __ldg(accession + p);
        int32_t id = (acc >= 0 && acc < lut_size) ? __ldg(lut + acc) : -1;
        if (id >= 0) {
            for (int64_t q = rb; q < p; ++q) {
                if (__ldg(accession + q) == acc) {
                    id = -1;
                    break;
                }
            }
        }
        out[p] = id;
    }
}

template <typename PtrT>
__global__ void __launch_bounds__(kThreads)
features_simple_kernel(const int32_t *__restrict__ accession, const PtrT *__restrict__ gene_ptr, int64_t G,
                       const int32_t *__restrict__ lut, int32_t lut_size, int32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t g = warp0; g < G; g += nwarps)
        gene_rows_simple(accession, (int64_t)__ldg(gene_ptr + g), (int64_t)__ldg(gene_ptr + g + 1), lut, lut_size, out, lane);
}

// old value of the word when `id` is an attribute id (not 0xFFFF), else 0 and no access: predicated, no branch
__device__ __forceinline__ uint32_t atom_or_shared_if_id(uint32_t addr, uint32_t bits, uint32_t id) {
    uint32_t old;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0xFFFF;\n\tmov.u32 %0, 0;\n\t@p atom.shared.or.b32 %0, [%1], %2;\n\t}"
        : "=r"(old)
        : "r"(addr), "r"(bits), "r"(id)
        : "memory");
    return old;
}
// A value every lane holds, handed to the compiler as provably warp-uniform (REDUX writes a uniform register): branches
// on it need no reconvergence bookkeeping and the shuffles behind them no divergence check.
__device__ __forceinline__ int uniform(int v) { return __reduce_max_sync(0xffffffffu, v); }
// out[imm] = v when row < end (streaming store, predicated: no branch, the offset an immediate)
template <int kByteOffset>
__device__ __forceinline__ void store_row_if(int32_t *base, int32_t v, int row, int end) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %2, %3;\n\t@p st.global.cs.s32 [%0 + %4], %1;\n\t}" ::"l"(base), "r"(v), "r"(row),
                 "r"(end), "n"(kByteOffset)
                 : "memory");
}
__device__ __forceinline__ void store_shared_zero4(uint32_t addr) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
}
__device__ __forceinline__ void store_shared_zero(uint32_t addr) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(0u) : "memory");
}
__device__ __forceinline__ uint32_t load_shared_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t load_shared_u16(uint32_t addr) {
    uint16_t v;
    asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}

template <int U, int u = 0>
__device__ __forceinline__ void store_rows(int32_t *out_p, const uint32_t (&id)[U], int row, int end) {
    if constexpr (u < U) {
        store_row_if<u * 128>(out_p, id[u] >= 0xFFFFu ? -1 : (int32_t)id[u], row + u * 32, end);
        store_rows<U, u + 1>(out_p, id, row, end);
    }
}

// The streaming kernel.  A warp takes a block of 32 consecutive genes (33 row pointers: one coalesced load, kept in
// registers as offsets from the block's first row, the next block's already in flight) and cuts it into groups of GG
// genes.  A group's rows are one contiguous range, read 32 rows per step with lane = row — every load and store is a
// full 128-byte line whatever the genes' sizes — in batches of U steps.  The accession loads of the NEXT batch (next
// group, or the next block's first group) are issued before the current batch is worked on, so the HBM round trip
// overlaps the table lookups and the bitmap work instead of preceding them.  A row finds its gene inside the group
// with log2(GG) shuffles over the pointer registers (upper bound, so empty genes are skipped).  Repeats: per warp a
// bitmap over (attribute id, gene of the group) in shared memory — bit id * GG + gene, so the word follows from the id
// alone — set with atomicOr; a bit found set means another row of this gene has the id.  Only a batch that really
// holds a repeat — rare in real tables, never in the synthetic ones — pays for the exact answer: its rows compare
// their accession with the earlier rows of their gene, so the FIRST of equal rows keeps the key (features.py:32).
// Lanes clear the words they set when the group fitted one batch; a longer group is cleared wholesale.
//
// The kernel is bound by instruction issue, not by memory (ncu: 80 % issue-active at 23 % of the DRAM peak), hence the
// raw shared-memory addresses, the clamped table index with a sentinel entry, and ids kept as 0xFFFF = "none" until
// the store.
//
// SLUT: the accession -> id table is staged in shared memory as uint16 (0xFFFF = unknown).  A gather of 32 scattered
// words through L1 costs one wavefront per distinct line — 32 cycles of the SM's load pipe per 32 rows — while the
// same gather from shared memory costs its bank conflicts (~3.5 wavefronts).
template <typename PtrT, int GG, int U, int THREADS, bool SLUT>
__global__ void __launch_bounds__(THREADS, SLUT ? 2 : GG == 32 ? 5 : 1)
features_kernel(const int32_t *__restrict__ accession, const PtrT *__restrict__ gene_ptr, int64_t G,
                const int32_t *__restrict__ lut, int32_t lut_size, int32_t words, int32_t *__restrict__ out) {
    constexpr unsigned kFull = 0xffffffffu;
    constexpr uint32_t kNone = 0xFFFFu;
    constexpr int kIdsPerWord = 32 / GG;  // GG = 8: four ids per bitmap word
    constexpr int kIdShift = GG == 8 ? 2 : GG == 16 ? 1 : 0;
    static_assert(GG == 8 || GG == 16 || GG == 32, "group sizes with a whole number of ids per word");
    // Dense groups wipe their bitmap with 16-byte stores (6 per lane for 2,659 attributes) — fewer instructions than
    // every row clearing its own word; sparse groups (a few rows against 5 KB of bitmap) clear what they set.
    constexpr bool kOwnClear = GG == 16;
    // GG = 32 (sparse tables: the whole 32-gene block is one group, usually one batch): 64 bits per gene slot, keyed by
    // id & 63 — 256 bytes of bitmap per warp instead of 10 KB.  Two different ids of a gene that share a bit only send the
    // batch through the exact comparison; with the one or two rows a gene has there, that is a few percent of the batches.
    constexpr bool kHashed = GG == 32;
    extern __shared__ __align__(16) uint32_t sBitmap[];  // `words` (a multiple of 4) per warp, then the uint16 table (SLUT)
    const int lane = threadIdx.x & 31;
    static_assert(U <= 8, "eight start-mask words per warp");
    uint32_t *bm = sBitmap + (threadIdx.x >> 5) * (words + 8);  // the warp's bitmap, then its start masks
    for (int i = lane; i < words; i += 32) bm[i] = 0;
    const uint32_t bm_addr = (uint32_t)__cvta_generic_to_shared(bm);
    const uint32_t mask_addr = bm_addr + words * 4;
    const uint32_t lanemask_le = 0xffffffffu >> (31 - lane);
    const uint32_t lut_addr = (uint32_t)__cvta_generic_to_shared(sBitmap + (THREADS / 32) * (words + 8));
    if (SLUT) {
        uint16_t *dst = reinterpret_cast<uint16_t *>(sBitmap + (THREADS / 32) * (words + 8));
        for (int i = threadIdx.x; i < lut_size; i += THREADS) dst[i] = (uint16_t)__ldg(lut + i);  // -1 -> 0xFFFF
        if (threadIdx.x == 0) dst[lut_size] = (uint16_t)kNone;  // where out-of-range accessions are clamped to
        __syncthreads();
    }
    __syncwarp();
    const int64_t warp0 = (int64_t)blockIdx.x * (THREADS / 32) + uniform((int)(threadIdx.x >> 5));
    const int64_t stride = (int64_t)gridDim.x * THREADS;  // genes per sweep of the grid: 32 per warp
    int64_t base = warp0 * 32;
    if (base >= G) return;
    int64_t p_lo = (int64_t)__ldg(gene_ptr + min(base + lane, G));  // first row of gene base + lane
    int64_t p_end = (int64_t)__ldg(gene_ptr + min(base + 32, G));   // end of the block's rows
    // ... and of the warp's next block: pointers run two blocks ahead, so that the next block's first accession loads
    // can be issued from this block's last batch without waiting for them
    int64_t p_lo1 = 0, p_end1 = 0;
    if (base + stride < G) {
        p_lo1 = (int64_t)__ldg(gene_ptr + min(base + stride + lane, G));
        p_end1 = (int64_t)__ldg(gene_ptr + min(base + stride + 32, G));
    }
    int32_t acc_n[U];       // accessions of the pending batch
    int pend_off = -1, pend_r1 = -1;  // ... which covers rows [pend_off, pend_r1) of the block it belongs to
    for (; base < G; base += stride) {
        const int ng = (int)min((int64_t)32, G - base);
        const int64_t P0 = __shfl_sync(kFull, p_lo, 0);
        const int64_t block_rows = p_end - P0;
        const int rel = (int)min(p_lo - P0, (int64_t)0x7fffffff);  // lanes past the last gene hold the block's end
        const int64_t next = base + stride;
        p_lo = p_lo1;  // from here on: the next block's pointers
        p_end = p_end1;
        if (next + stride < G) {
            p_lo1 = (int64_t)__ldg(gene_ptr + min(next + stride + lane, G));
            p_end1 = (int64_t)__ldg(gene_ptr + min(next + stride + 32, G));
        }
        if (block_rows > 0x7fffffff - 32 * U) {  // offsets (plus one batch) would not fit 32 bits: gene by gene
            for (int g = 0; g < ng; ++g)
                gene_rows_simple(accession, (int64_t)__ldg(gene_ptr + base + g), (int64_t)__ldg(gene_ptr + base + g + 1),
                                 lut, lut_size, out, lane);
            pend_off = -1;
            continue;
        }
        const int32_t *__restrict__ acc_b = accession + P0;
        int32_t *__restrict__ out_b = out + P0;
        for (int sg = 0; sg < ng; sg += GG) {
            const int r0 = uniform(__shfl_sync(kFull, rel, sg));
            const int r1 = uniform(sg + GG >= ng ? (int)block_rows : __shfl_sync(kFull, rel, (sg + GG) & 31));
            bool dirty = false;
            int slot = -1;  // gene slot of the last row seen so far (warp-uniform)
            for (int off = r0; off < r1; off += 32 * U) {
                if (pend_off != off || pend_r1 != r1) {  // nothing in flight for this batch (warp-uniform)
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int r = off + u * 32 + lane;
                        acc_n[u] = r < r1 ? __ldcs(acc_b + r) : -1;
                    }
                }
                uint32_t id[U];  // attribute id or kNone
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (SLUT) {
                        id[u] = load_shared_u16(lut_addr + 2u * min((uint32_t)acc_n[u], (uint32_t)lut_size));
                    } else {
                        const int v = (unsigned)acc_n[u] < (unsigned)lut_size ? __ldg(lut + acc_n[u]) : -1;
                        id[u] = v < 0 ? kNone : (uint32_t)v;
                    }
                }
                // the loads of the batch behind this one
                pend_off = -1;
                if (off + 32 * U < r1) {
                    pend_off = off + 32 * U;
                    pend_r1 = r1;
                } else if (sg + GG < ng) {
                    pend_off = r1;  // the next group's rows follow this group's
                    pend_r1 = uniform(sg + 2 * GG >= ng ? (int)block_rows : __shfl_sync(kFull, rel, (sg + 2 * GG) & 31));
                }
                if (pend_off >= 0) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int r = pend_off + u * 32 + lane;
                        acc_n[u] = r < pend_r1 ? __ldcs(acc_b + r) : -1;
                    }
                } else if (next < G) {  // the first group of the warp's next block (its pointers have arrived by now)
                    const int64_t P0n = __shfl_sync(kFull, p_lo, 0);
                    const int ngn = (int)min((int64_t)32, G - next);
                    const int64_t r1n = (GG >= ngn ? p_end : __shfl_sync(kFull, p_lo, GG & 31)) - P0n;
                    if (p_end - P0n <= 0x7fffffff - 32 * U) {
                        pend_off = 0;
                        pend_r1 = uniform((int)r1n);
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int r = u * 32 + lane;
                            acc_n[u] = r < pend_r1 ? __ldcs(accession + P0n + r) : -1;
                        }
                    }
                }
                // The group's gene starts inside this batch, one 32-bit mask per step (a handful of lanes set them): a
                // row's gene slot is the number of distinct starts at or before it, minus one — one popc instead of a
                // search.  Genes without rows share their successor's start, so slots stay below GG.
                if (lane < U) store_shared_zero(mask_addr + lane * 4);
                __syncwarp();
                {
                    const int d = rel - off;
                    if (lane >= sg && lane < sg + GG && d >= 0 && d < 32 * U && rel < r1)
                        atom_or_shared_if_id(mask_addr + ((d >> 5) << 2), 1u << (d & 31), 0u);
                }
                __syncwarp();
                uint32_t seen = 0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t starts = load_shared_u32(mask_addr + u * 4);
                    const int g = slot + __popc(starts & lanemask_le);
                    slot += __popc(starts);
                    const uint32_t bit = kHashed ? 1u << (id[u] & 31) : 1u << ((id[u] & (kIdsPerWord - 1)) * GG + g);
                    const uint32_t word = kHashed ? (uint32_t)(g * 2) + ((id[u] >> 5) & 1u) : id[u] >> kIdShift;
                    seen |= atom_or_shared_if_id(bm_addr + (word << 2), bit, id[u]) & bit;
                }
                if (__any_sync(kFull, seen != 0)) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (off + u * 32 < r1) {
                            const int r = off + u * 32 + lane;
                            int pos = 0;
#pragma unroll
                            for (int s = GG / 2; s >= 1; s >>= 1)
                                if (__shfl_sync(kFull, rel, sg + pos + s) <= r) pos += s;
                            const int gs = __shfl_sync(kFull, rel, sg + pos);
                            if (id[u] != kNone && r < r1) {
                                const int32_t mine = __ldg(acc_b + r);
                                for (int q = gs; q < r; ++q)
                                    if (__ldg(acc_b + q) == mine) {
                                        id[u] |= 0x10000u;  // a repeat: stored as -1, its bitmap word still cleared below
                                        break;
                                    }
                            }
                        }
                    }
                }
                const bool own_clear = kOwnClear && !dirty && off + 32 * U >= r1;
                {
                    int32_t *out_p = out_b + off + lane;
                    store_rows<U>(out_p, id, off + lane, r1);
                }
                __syncwarp();
                if (own_clear) {
#pragma unroll
                    for (int u = 0; u < U; ++u)  // all writers store the same zero
                        if ((id[u] & 0xFFFFu) != kNone) store_shared_zero(bm_addr + (((id[u] & 0xFFFFu) >> kIdShift) << 2));
                } else {
                    dirty = true;
                }
            }
            if (dirty)
                for (int i = lane * 4; i < words; i += 128) store_shared_zero4(bm_addr + i * 4);
            __syncwarp();
        }
    }
}

// Compact ids (GCRF_FLAG_IDX_U16): uint16 -> int32, 0xFFFF -> -1.  Eight ids per thread and step: one 16-byte load,
// two 16-byte stores.  `out` is 16-byte aligned and padded to a multiple of 8 entries; `in` is 16-byte aligned.
__global__ void __launch_bounds__(kThreads) widen_u16_kernel(const uint16_t *__restrict__ in, int32_t *__restrict__ out, int64_t n) {
    const int64_t groups = (n + 7) >> 3;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
        uint4 v;
        if ((g << 3) + 8 <= n) {
            v = __ldg(reinterpret_cast<const uint4 *>(in) + g);
        } else {  // the tail: do not read past the caller's array
            uint16_t t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = (g << 3) + k < n ? in[(g << 3) + k] : (uint16_t)0xFFFF;
            v.x = t[0] | ((uint32_t)t[1] << 16);
            v.y = t[2] | ((uint32_t)t[3] << 16);
            v.z = t[4] | ((uint32_t)t[5] << 16);
            v.w = t[6] | ((uint32_t)t[7] << 16);
        }
        auto lo = [](uint32_t w) -> int { const int x = (int)(w & 0xFFFFu); return x == 0xFFFF ? -1 : x; };
        auto hi = [](uint32_t w) -> int { const int x = (int)(w >> 16); return x == 0xFFFF ? -1 : x; };
        int4 *dst = reinterpret_cast<int4 *>(out) + 2 * g;
        dst[0] = make_int4(lo(v.x), hi(v.x), lo(v.y), hi(v.y));
        dst[1] = make_int4(lo(v.z), hi(v.z), lo(v.w), hi(v.w));
    }
}

}  // namespace

cudaError_t launch_widen_u16(const uint16_t *in, int32_t *out, int64_t n, int num_sms, cudaStream_t stream, int64_t *launches) {
    if (n <= 0) return cudaSuccess;
    int64_t blocks = ((n + 7) / 8 + kThreads - 1) / kThreads;
    if (blocks > (int64_t)num_sms * 8) blocks = (int64_t)num_sms * 8;
    widen_u16_kernel<<<(int)blocks, kThreads, 0, stream>>>(in, out, n);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

namespace {

template <typename PtrT, int GG, int U, int THREADS, bool SLUT>
cudaError_t launch_features_bitmap(const int32_t *accession, const PtrT *gene_ptr, int64_t G, const int32_t *lut,
                                   int32_t lut_size, int32_t words, int32_t *out, int num_sms, cudaStream_t stream) {
    auto kernel = features_kernel<PtrT, GG, U, THREADS, SLUT>;
    const size_t smem = (size_t)(words + 8) * (THREADS / 32) * sizeof(uint32_t) + (SLUT ? ((size_t)lut_size * 2 + 2 + 15) / 16 * 16 : 0);
    cudaError_t err = cudaSuccess;
    if (smem > 48 * 1024) {
        err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
    }
    int per_sm = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, smem);
    if (err != cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    int64_t blocks = ((G + 31) / 32 + THREADS / 32 - 1) / (THREADS / 32);  // one warp per 32 genes
    if (blocks > (int64_t)num_sms * per_sm) blocks = (int64_t)num_sms * per_sm;  // one resident wave, grid-stride
    kernel<<<(int)blocks, THREADS, smem, stream>>>(accession, gene_ptr, G, lut, lut_size, words, out);
    return cudaGetLastError();
}

template <typename PtrT>
cudaError_t launch_features_t(const int32_t *accession, const PtrT *gene_ptr, int64_t G, int64_t nnz, const int32_t *lut,
                              int32_t lut_size, int32_t num_attrs, int32_t *out, int num_sms, cudaStream_t stream) {
    // sparse tables (real annotation: 1.4 rows per gene) take the 32-gene block as one group with hashed 64-bit bitmaps
    // (4 = exact bitmaps over groups of 16 genes instead), dense ones groups of 8 genes with exact bitmaps
    const bool sparse = nnz < 6 * G;
    const char *force = getenv("GCRF_FEATURES_SIMPLE");  // A/B and tests
    const bool sparse_exact = sparse && force && force[0] == '4';
    const int32_t words = sparse && !sparse_exact ? 64 : (int32_t)(((int64_t)num_attrs * (sparse ? 16 : 8) + 127) / 128 * 4);  // bitmap words per warp (16-byte units)
    const size_t bitmaps = (size_t)(words + 8) * (kThreads / 32) * sizeof(uint32_t);
    // 1 = gene by gene, 2 = table never / 3 = always in shared memory
    if (words == 0 || num_attrs >= 0xFFFF || bitmaps > 96 * 1024 || (force && force[0] == '1')) {  // bitmaps do not fit: the exact gene-by-gene kernel
        int64_t blocks = (G * 32 + kThreads - 1) / kThreads;
        if (blocks > (int64_t)num_sms * 16) blocks = (int64_t)num_sms * 16;
        features_simple_kernel<PtrT><<<(int)blocks, kThreads, 0, stream>>>(accession, gene_ptr, G, lut, lut_size, out);
        return cudaGetLastError();
    }
    if (sparse_exact) return launch_features_bitmap<PtrT, 16, 2, kThreads, false>(accession, gene_ptr, G, lut, lut_size, words, out, num_sms, stream);
    if (sparse) return launch_features_bitmap<PtrT, 32, 2, kThreads, false>(accession, gene_ptr, G, lut, lut_size, words, out, num_sms, stream);
    // dense tables with enough rows to pay for staging the table (82 KB of L2 reads per CTA): two 512-thread CTAs per SM
    const size_t staged = 2 * bitmaps + (size_t)lut_size * 2 + 16;
    if (staged <= 110 * 1024 && !(force && force[0] == '2') &&
        (nnz >= (int64_t)num_sms * 8192 || (force && force[0] == '3')))
        return launch_features_bitmap<PtrT, 8, 8, 512, true>(accession, gene_ptr, G, lut, lut_size, words, out, num_sms, stream);
    return launch_features_bitmap<PtrT, 8, 8, kThreads, false>(accession, gene_ptr, G, lut, lut_size, words, out, num_sms, stream);
}

}  // namespace

cudaError_t launch_features(const int32_t *accession, const int32_t *gene_ptr32, const int64_t *gene_ptr64, int64_t G,
                            int64_t nnz, const int32_t *lut, int32_t lut_size, int32_t num_attrs, int32_t *attr_idx_out,
                            int num_sms, cudaStream_t stream, int64_t *launches) {
    if (G <= 0) return cudaSuccess;
    cudaError_t err = gene_ptr64 ? launch_features_t<int64_t>(accession, gene_ptr64, G, nnz, lut, lut_size, num_attrs,
                                                               attr_idx_out, num_sms, stream)
                                 : launch_features_t<int32_t>(accession, gene_ptr32, G, nnz, lut, lut_size, num_attrs,
                                                               attr_idx_out, num_sms, stream);
    if (launches) *launches += 1;
    return err;
}

}  // namespace gcrf
