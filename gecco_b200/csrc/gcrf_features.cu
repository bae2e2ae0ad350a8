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

// The streaming kernel.  A warp takes a block of 32 consecutive genes (33 row pointers: one coalesced load, kept in
// registers as offsets from the block's first row, the next block's already in flight) and cuts it into groups of GG
// genes.  A group's rows are one contiguous range, read 32 rows per step with lane = row — every load and store is a
// full 128-byte line whatever the genes' sizes — U steps in flight (accession loads, then the LUT gathers, then the
// stores).  A row finds its gene inside the group with log2(GG) shuffles over the pointer registers (upper bound, so
// empty genes are skipped).  Repeats: a bitmap per gene of the group over the attribute ids in shared memory
// (atomicOr; a set bit means another row of this gene has the id).  Only a batch that really holds a repeat — rare in
// real tables, never in the synthetic ones — pays for the exact answer: its rows compare their accession with the
// earlier rows of their gene, so the FIRST of equal rows keeps the key (features.py:32).  Lanes clear the words they
// set when the group fitted one batch; a longer group is cleared wholesale.
template <typename PtrT, int GG, int U>
__global__ void __launch_bounds__(kThreads)
features_kernel(const int32_t *__restrict__ accession, const PtrT *__restrict__ gene_ptr, int64_t G,
                const int32_t *__restrict__ lut, int32_t lut_size, int32_t words, int32_t *__restrict__ out) {
    constexpr unsigned kFull = 0xffffffffu;
    extern __shared__ uint32_t sBitmap[];  // GG * words per warp
    const int lane = threadIdx.x & 31;
    uint32_t *bm = sBitmap + (threadIdx.x >> 5) * (GG * words);
    for (int i = lane; i < GG * words; i += 32) bm[i] = 0;
    __syncwarp();
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t stride = (((int64_t)gridDim.x * blockDim.x) >> 5) * 32;
    int64_t base = warp0 * 32;
    if (base >= G) return;
    int64_t p_lo = (int64_t)__ldg(gene_ptr + min(base + lane, G));  // first row of gene base + lane
    int64_t p_end = (int64_t)__ldg(gene_ptr + min(base + 32, G));   // end of the block's rows
    for (; base < G; base += stride) {
        const int ng = (int)min((int64_t)32, G - base);
        const int64_t P0 = __shfl_sync(kFull, p_lo, 0);
        const int64_t block_rows = p_end - P0;
        const int rel = (int)min(p_lo - P0, (int64_t)0x7fffffff);  // lanes past the last gene hold the block's end
        const int64_t next = base + stride;
        if (next < G) {
            p_lo = (int64_t)__ldg(gene_ptr + min(next + lane, G));
            p_end = (int64_t)__ldg(gene_ptr + min(next + 32, G));
        }
        if (block_rows > 0x7fffffff) {  // offsets would not fit 32 bits: gene by gene
            for (int g = 0; g < ng; ++g)
                gene_rows_simple(accession, (int64_t)__ldg(gene_ptr + base + g), (int64_t)__ldg(gene_ptr + base + g + 1),
                                 lut, lut_size, out, lane);
            continue;
        }
        const int32_t *__restrict__ acc_b = accession + P0;
        int32_t *__restrict__ out_b = out + P0;
        for (int sg = 0; sg < ng; sg += GG) {
            const int r0 = __shfl_sync(kFull, rel, sg);
            const int r1 = sg + GG >= ng ? (int)block_rows : __shfl_sync(kFull, rel, (sg + GG) & 31);
            bool dirty = false;
            for (int off = r0; off < r1; off += 32 * U) {
                int32_t acc[U], id[U];
                int slot[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int r = off + u * 32 + lane;
                    acc[u] = r < r1 ? __ldg(acc_b + r) : -1;
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    id[u] = (unsigned)acc[u] < (unsigned)lut_size ? __ldg(lut + acc[u]) : -1;
                uint32_t seen = 0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    slot[u] = -1;
                    if (off + u * 32 < r1) {  // warp-uniform
                        const int r = off + u * 32 + lane;
                        int pos = 0;  // the last gene of the group that starts at or before row r
#pragma unroll
                        for (int s = GG / 2; s >= 1; s >>= 1)
                            if (__shfl_sync(kFull, rel, sg + pos + s) <= r) pos += s;
                        if (id[u] >= 0) {
                            slot[u] = pos * words + (id[u] >> 5);
                            const uint32_t bit = 1u << (id[u] & 31);
                            seen |= atomicOr(&bm[slot[u]], bit) & bit;
                        }
                    }
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
                            if (id[u] >= 0)
                                for (int q = gs; q < r; ++q)
                                    if (__ldg(acc_b + q) == acc[u]) {
                                        id[u] = -1;
                                        break;
                                    }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int r = off + u * 32 + lane;
                    if (r < r1) out_b[r] = id[u];
                }
                __syncwarp();
                if (!dirty && off + 32 * U >= r1) {
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (slot[u] >= 0) bm[slot[u]] = 0;  // all writers store the same zero
                } else {
                    dirty = true;
                }
            }
            if (dirty)
                for (int i = lane; i < GG * words; i += 32) bm[i] = 0;
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

template <typename PtrT, int GG, int U>
cudaError_t launch_features_bitmap(const int32_t *accession, const PtrT *gene_ptr, int64_t G, const int32_t *lut,
                                   int32_t lut_size, int32_t words, int32_t *out, int num_sms, cudaStream_t stream) {
    auto kernel = features_kernel<PtrT, GG, U>;
    const size_t smem = (size_t)GG * words * (kThreads / 32) * sizeof(uint32_t);
    cudaError_t err = cudaSuccess;
    if (smem > 48 * 1024) {
        err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
    }
    int per_sm = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem);
    if (err != cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    int64_t blocks = ((G + 31) / 32 + kThreads / 32 - 1) / (kThreads / 32);  // one warp per 32 genes
    if (blocks > (int64_t)num_sms * per_sm) blocks = (int64_t)num_sms * per_sm;  // one resident wave, grid-stride
    kernel<<<(int)blocks, kThreads, smem, stream>>>(accession, gene_ptr, G, lut, lut_size, words, out);
    return cudaGetLastError();
}

template <typename PtrT>
cudaError_t launch_features_t(const int32_t *accession, const PtrT *gene_ptr, int64_t G, int64_t nnz, const int32_t *lut,
                              int32_t lut_size, int32_t num_attrs, int32_t *out, int num_sms, cudaStream_t stream) {
    const int32_t words = (num_attrs + 31) / 32;
    // sparse tables (real annotation: 1.4 rows per gene) are cut into groups of 16 genes, dense ones into groups of 8
    const bool sparse = nnz < 6 * G;
    const size_t smem = (size_t)(sparse ? 16 : 8) * words * (kThreads / 32) * sizeof(uint32_t);
    const char *force = getenv("GCRF_FEATURES_SIMPLE");  // A/B and tests
    if (words == 0 || smem > 96 * 1024 || (force && force[0] == '1')) {  // bitmaps do not fit: the exact gene-by-gene kernel
        int64_t blocks = (G * 32 + kThreads - 1) / kThreads;
        if (blocks > (int64_t)num_sms * 16) blocks = (int64_t)num_sms * 16;
        features_simple_kernel<PtrT><<<(int)blocks, kThreads, 0, stream>>>(accession, gene_ptr, G, lut, lut_size, out);
        return cudaGetLastError();
    }
    if (sparse) return launch_features_bitmap<PtrT, 16, 2>(accession, gene_ptr, G, lut, lut_size, words, out, num_sms, stream);
    return launch_features_bitmap<PtrT, 8, 8>(accession, gene_ptr, G, lut, lut_size, words, out, num_sms, stream);
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
