// gcrf_wire.cu — a compact wire format for host-buffer calls: what gcrf_marginals_windowed moves over PCIe is the
// whole cost of such a call (the kernels are ~2 % of it), so the bytes are the lever.
//
// The CSR batch of include/gecco_crf_b200.h costs 4 bytes per attribute id, 4 per gene (row pointer) and 8 per gene
// (float64 marginal).  A gcrf_wire holds the same batch as
//     contig_ptr[C+1]   int32, unchanged (4 bytes per contig)
//     len_ids[G]        ids of gene g            uint8, or uint16 when a gene has more than 255 of them
//     len_bytes[G]      bytes of gene g's stream (same width)
//     stream[]          per gene: its attribute ids SORTED ascending, unknown ids (outside [0, A)) mapped to A, as
//                       deltas (first id absolute) in LEB128 — 7 value bits per byte, high bit = "more follows"
// in ONE page-locked block, i.e. one host-to-device copy.  Sorting is free for the result: a gene's features are a set
// (gecco/crf/features.py:32) and the device forms row sums in exact integer arithmetic, so the order of a row does not
// matter (rows of >= ModelDev::fx_nsafe ids, which take the float path, may differ in the last bit).  For config 2
// (Poisson(25) ids per gene out of 2,659): 1.3 bytes per id + 2 per gene instead of 4 + 4.
//
// On the device three small kernels rebuild gene_ptr and attr_idx (block sums of the two length arrays, a one-block
// scan of those, scan-within-block + decode), then the marginal kernels run unchanged.
#include "../../include/gecco_crf_b200.h"
#include "gcrf_kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

struct gcrf_wire {
    int64_t C = 0, G = 0, nnz = 0, stream_bytes = 0;
    int32_t A = 0;
    int32_t len_width = 1;  // bytes per entry of len_ids / len_bytes: 1 or 2
    char *block = nullptr;  // page-locked: [contig_ptr | len_ids | len_bytes | stream], every part 16-byte aligned
    size_t off_len_ids = 0, off_len_bytes = 0, off_stream = 0, total = 0;
    // contig-aligned slices of about equal bytes (the call overlaps slice k's way back with slice k+1's way in)
    static constexpr int kMaxSlices = 8;
    int n_slices = 1;
    int64_t s_contig[kMaxSlices + 1] = {}, s_gene[kMaxSlices + 1] = {}, s_id[kMaxSlices + 1] = {}, s_byte[kMaxSlices + 1] = {};
};

namespace gcrf {

namespace {

constexpr int kWireThreads = 256, kWirePerThread = 8, kWireChunk = kWireThreads * kWirePerThread;  // genes per block

inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// LEB128 of one non-negative value
inline void put_varint(std::vector<uint8_t> &out, uint32_t v) {
    while (v >= 128) {
        out.push_back((uint8_t)(v | 128));
        v >>= 7;
    }
    out.push_back((uint8_t)v);
}

// ---- kernel 1: per block of kWireChunk genes, the sums of both length arrays
template <typename LenT>
__global__ void __launch_bounds__(kWireThreads)
wire_block_sums_kernel(const LenT *__restrict__ len_ids, const LenT *__restrict__ len_bytes, int64_t G, int64_t *__restrict__ sums) {
    __shared__ long long sA[kWireThreads / 32], sB[kWireThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kWireChunk;
    long long a = 0, b = 0;
    for (int k = 0; k < kWirePerThread; ++k) {
        const int64_t g = base + threadIdx.x + (int64_t)k * kWireThreads;
        if (g < G) {
            a += len_ids[g];
            b += len_bytes[g];
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, d);
        b += __shfl_down_sync(0xffffffffu, b, d);
    }
    if ((threadIdx.x & 31) == 0) {
        sA[threadIdx.x >> 5] = a;
        sB[threadIdx.x >> 5] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long ta = 0, tb = 0;
        for (int w = 0; w < kWireThreads / 32; ++w) {
            ta += sA[w];
            tb += sB[w];
        }
        sums[2 * (int64_t)blockIdx.x] = ta;
        sums[2 * (int64_t)blockIdx.x + 1] = tb;
    }
}

// ---- kernel 2: exclusive scan of the block sums, in place (one block; n = number of chunks)
__global__ void __launch_bounds__(1024)
wire_scan_sums_kernel(int64_t *__restrict__ sums, int64_t n) {
    __shared__ long long sWarp[2][32];
    __shared__ long long sCarry[2];
    if (threadIdx.x < 2) sCarry[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        long long v[2] = {i < n ? sums[2 * i] : 0, i < n ? sums[2 * i + 1] : 0};
        long long inc[2] = {v[0], v[1]};
        for (int c = 0; c < 2; ++c) {
            for (int d = 1; d < 32; d <<= 1) {
                const long long y = __shfl_up_sync(0xffffffffu, inc[c], d);
                if (lane >= d) inc[c] += y;
            }
            if (lane == 31) sWarp[c][warp] = inc[c];
        }
        __syncthreads();
        if (warp == 0) {
            for (int c = 0; c < 2; ++c) {
                long long w = sWarp[c][lane];
                for (int d = 1; d < 32; d <<= 1) {
                    const long long y = __shfl_up_sync(0xffffffffu, w, d);
                    if (lane >= d) w += y;
                }
                sWarp[c][lane] = w;  // inclusive scan of the warp totals
            }
        }
        __syncthreads();
        for (int c = 0; c < 2; ++c) {
            const long long before = sCarry[c] + (warp > 0 ? sWarp[c][warp - 1] : 0) + inc[c] - v[c];
            if (i < n) sums[2 * i + c] = before;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            sCarry[0] += sWarp[0][31];
            sCarry[1] += sWarp[1][31];
        }
        __syncthreads();
    }
}

// ---- kernel 3: offsets of every gene inside its block (scan), gene_ptr, and the ids themselves
template <typename LenT>
__global__ void __launch_bounds__(kWireThreads)
wire_decode_kernel(const LenT *__restrict__ len_ids, const LenT *__restrict__ len_bytes, const uint8_t *__restrict__ stream,
                   int64_t G, int64_t id_base, const int64_t *__restrict__ sums, int32_t *__restrict__ gene_ptr,
                   int32_t *__restrict__ attr_idx) {
    __shared__ int sIds[kWireChunk], sBytes[kWireChunk];  // lengths, then exclusive offsets within the block
    __shared__ int sWarpA[kWireThreads / 32], sWarpB[kWireThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kWireChunk;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = 0; k < kWirePerThread; ++k) {
        const int j = tid + k * kWireThreads;
        const int64_t g = base + j;
        sIds[j] = g < G ? (int)len_ids[g] : 0;
        sBytes[j] = g < G ? (int)len_bytes[g] : 0;
    }
    __syncthreads();
    // thread t scans entries [8t, 8t + 8), then the thread totals are scanned across the block
    int a[kWirePerThread], b[kWirePerThread], ta = 0, tb = 0;
#pragma unroll
    for (int k = 0; k < kWirePerThread; ++k) {
        a[k] = ta;
        b[k] = tb;
        ta += sIds[tid * kWirePerThread + k];
        tb += sBytes[tid * kWirePerThread + k];
    }
    int ia = ta, ib = tb;
    for (int d = 1; d < 32; d <<= 1) {
        const int ya = __shfl_up_sync(0xffffffffu, ia, d), yb = __shfl_up_sync(0xffffffffu, ib, d);
        if (lane >= d) {
            ia += ya;
            ib += yb;
        }
    }
    if (lane == 31) {
        sWarpA[warp] = ia;
        sWarpB[warp] = ib;
    }
    __syncthreads();
    int wa = 0, wb = 0;
    for (int w = 0; w < warp; ++w) {
        wa += sWarpA[w];
        wb += sWarpB[w];
    }
    __syncthreads();  // everybody has read the lengths of its eight entries
#pragma unroll
    for (int k = 0; k < kWirePerThread; ++k) {
        sIds[tid * kWirePerThread + k] = wa + ia - ta + a[k];
        sBytes[tid * kWirePerThread + k] = wb + ib - tb + b[k];
    }
    __syncthreads();
    const int64_t id0 = id_base + sums[2 * (int64_t)blockIdx.x], byte0 = sums[2 * (int64_t)blockIdx.x + 1];
    for (int k = 0; k < kWirePerThread; ++k) {
        const int j = tid + k * kWireThreads;  // neighbouring threads take neighbouring genes: their bytes share lines
        const int64_t g = base + j;
        if (g >= G) break;
        const int64_t p0 = id0 + sIds[j];
        gene_ptr[g] = (int32_t)p0;
        const int n = (int)len_ids[g];
        const uint8_t *src = stream + byte0 + sBytes[j];
        int32_t prev = 0;
        for (int i = 0; i < n; ++i) {
            uint32_t v = 0, byte;
            int shift = 0;
            do {
                byte = *src++;
                v |= (byte & 127u) << shift;
                shift += 7;
            } while (byte & 128u);
            prev += (int32_t)v;
            attr_idx[p0 + i] = prev;
        }
        if (g == G - 1) gene_ptr[G] = (int32_t)(p0 + n);
    }
}

}  // namespace

int64_t wire_chunks(int64_t G) { return (G + kWireChunk - 1) / kWireChunk; }

cudaError_t launch_wire_decode(const void *len_ids, const void *len_bytes, int32_t len_width, const uint8_t *stream, int64_t G,
                               int64_t id_base, int64_t *sums, int32_t *gene_ptr, int32_t *attr_idx, cudaStream_t cuda_stream,
                               int64_t *launches) {
    if (G <= 0) return cudaSuccess;
    const int64_t nb = wire_chunks(G);
    if (len_width == 1) {
        wire_block_sums_kernel<uint8_t><<<(unsigned)nb, kWireThreads, 0, cuda_stream>>>(static_cast<const uint8_t *>(len_ids),
                                                                                     static_cast<const uint8_t *>(len_bytes), G, sums);
    } else {
        wire_block_sums_kernel<uint16_t><<<(unsigned)nb, kWireThreads, 0, cuda_stream>>>(static_cast<const uint16_t *>(len_ids),
                                                                                      static_cast<const uint16_t *>(len_bytes), G, sums);
    }
    wire_scan_sums_kernel<<<1, 1024, 0, cuda_stream>>>(sums, nb);
    if (len_width == 1) {
        wire_decode_kernel<uint8_t><<<(unsigned)nb, kWireThreads, 0, cuda_stream>>>(static_cast<const uint8_t *>(len_ids),
                                                                                 static_cast<const uint8_t *>(len_bytes), stream, G, id_base,
                                                                                 sums, gene_ptr, attr_idx);
    } else {
        wire_decode_kernel<uint16_t><<<(unsigned)nb, kWireThreads, 0, cuda_stream>>>(static_cast<const uint16_t *>(len_ids),
                                                                                  static_cast<const uint16_t *>(len_bytes), stream, G, id_base,
                                                                                  sums, gene_ptr, attr_idx);
    }
    if (launches) *launches += 3;
    return cudaGetLastError();
}

}  // namespace gcrf

// ---------------------------------------------------------------------------------------------------------------------
// host side: encoder, accessors, host decoder (the CPU test-suite checks the format without a GPU)
// ---------------------------------------------------------------------------------------------------------------------

namespace {

struct EncodedChunk {
    std::vector<uint8_t> stream;
    std::vector<uint32_t> n_ids, n_bytes;
};

template <typename PtrT>
void encode_range(const PtrT *gene_ptr, const int32_t *attr_idx, int64_t g0, int64_t g1, int32_t A, EncodedChunk *out) {
    std::vector<uint32_t> row;
    out->n_ids.reserve((size_t)(g1 - g0));
    out->n_bytes.reserve((size_t)(g1 - g0));
    out->stream.reserve((size_t)((gene_ptr[g1] - gene_ptr[g0]) * 3 / 2 + 16));
    for (int64_t g = g0; g < g1; ++g) {
        row.clear();
        for (int64_t p = (int64_t)gene_ptr[g]; p < (int64_t)gene_ptr[g + 1]; ++p) {
            const uint32_t a = (uint32_t)attr_idx[p];
            row.push_back(a < (uint32_t)A ? a : (uint32_t)A);  // every unknown id becomes the zero slot A
        }
        std::sort(row.begin(), row.end());
        const size_t before = out->stream.size();
        uint32_t prev = 0;
        for (uint32_t a : row) {
            gcrf::put_varint(out->stream, a - prev);
            prev = a;
        }
        out->n_ids.push_back((uint32_t)row.size());
        out->n_bytes.push_back((uint32_t)(out->stream.size() - before));
    }
}

thread_local char g_wire_error[256] = "";

int wire_fail(int code, const char *msg) {
    snprintf(g_wire_error, sizeof g_wire_error, "%s", msg);
    return code;
}

}  // namespace

extern "C" {

const char *gcrf_wire_last_error(void) { return g_wire_error; }

int gcrf_wire_encode(const int32_t *contig_ptr, const void *gene_ptr, const int32_t *attr_idx, int64_t C, int64_t G, int64_t nnz,
                     int32_t num_attrs, uint32_t flags, gcrf_wire **out) {
    if (!out) return wire_fail(GCRF_EINVAL, "out is NULL");
    *out = nullptr;
    if (C < 0 || G < 0 || nnz < 0 || num_attrs < 0) return wire_fail(GCRF_EINVAL, "negative size");
    if (G > 0 && (!contig_ptr || !gene_ptr)) return wire_fail(GCRF_EINVAL, "NULL array");
    if (nnz > 0 && !attr_idx) return wire_fail(GCRF_EINVAL, "attr_idx is NULL");
    if (nnz > 0x7fffffff) return wire_fail(GCRF_EUNSUPPORTED, "the wire format rebuilds 32-bit row pointers: nnz must stay below 2^31");
    const bool ptr64 = (flags & GCRF_FLAG_PTR64) != 0;
    auto row = [&](int64_t g) -> int64_t {
        return ptr64 ? static_cast<const int64_t *>(gene_ptr)[g] : (int64_t)static_cast<const int32_t *>(gene_ptr)[g];
    };
    if (G > 0 && (row(0) != 0 || row(G) != nnz)) return wire_fail(GCRF_EINVAL, "gene_ptr must start at 0 and end at nnz");
    for (int64_t g = 0; g < G; ++g)
        if (row(g + 1) < row(g)) return wire_fail(GCRF_EINVAL, "gene_ptr must be non-decreasing");

    unsigned nthreads = std::thread::hardware_concurrency();
    if (nthreads == 0) nthreads = 1;
    if (nthreads > 16) nthreads = 16;
    if ((int64_t)nthreads > G / 4096 + 1) nthreads = (unsigned)(G / 4096 + 1);
    std::vector<EncodedChunk> chunks(nthreads);
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthreads; ++t) {
        const int64_t g0 = G * t / nthreads, g1 = G * (t + 1) / nthreads;
        auto work = [=, &chunks]() {
            if (ptr64) encode_range(static_cast<const int64_t *>(gene_ptr), attr_idx, g0, g1, num_attrs, &chunks[t]);
            else encode_range(static_cast<const int32_t *>(gene_ptr), attr_idx, g0, g1, num_attrs, &chunks[t]);
        };
        if (t + 1 == nthreads) work();
        else pool.emplace_back(work);
    }
    for (auto &th : pool) th.join();

    uint32_t longest = 0;
    int64_t stream_bytes = 0;
    for (const auto &c : chunks) {
        for (uint32_t v : c.n_ids) longest = std::max(longest, v);
        for (uint32_t v : c.n_bytes) longest = std::max(longest, v);
        stream_bytes += (int64_t)c.stream.size();
    }
    if (longest > 0xFFFF) return wire_fail(GCRF_EUNSUPPORTED, "a gene with more than 65535 ids / stream bytes does not fit the wire format");
    gcrf_wire *w = new (std::nothrow) gcrf_wire();
    if (!w) return wire_fail(GCRF_ENOMEM, "out of host memory");
    w->C = C; w->G = G; w->nnz = nnz; w->A = num_attrs; w->stream_bytes = stream_bytes;
    w->len_width = longest > 0xFF ? 2 : 1;
    w->off_len_ids = gcrf::align16((size_t)(C + 1) * 4);
    w->off_len_bytes = gcrf::align16(w->off_len_ids + (size_t)G * w->len_width);
    w->off_stream = gcrf::align16(w->off_len_bytes + (size_t)G * w->len_width);
    w->total = gcrf::align16(w->off_stream + (size_t)stream_bytes + 16);
    if (cudaHostAlloc(reinterpret_cast<void **>(&w->block), w->total, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        // no CUDA runtime / device on this host: plain memory keeps the encoder usable (the copy is then pageable)
        w->block = static_cast<char *>(malloc(w->total));
        if (!w->block) {
            delete w;
            return wire_fail(GCRF_ENOMEM, "out of host memory");
        }
        w->len_width = -w->len_width;  // negative: block came from malloc
    }
    memset(w->block, 0, w->total);
    if (G > 0) memcpy(w->block, contig_ptr, (size_t)(C + 1) * 4);
    const int lw = w->len_width < 0 ? -w->len_width : w->len_width;
    int64_t g = 0;
    size_t spos = 0;
    for (const auto &c : chunks) {
        for (size_t k = 0; k < c.n_ids.size(); ++k, ++g) {
            if (lw == 1) {
                reinterpret_cast<uint8_t *>(w->block + w->off_len_ids)[g] = (uint8_t)c.n_ids[k];
                reinterpret_cast<uint8_t *>(w->block + w->off_len_bytes)[g] = (uint8_t)c.n_bytes[k];
            } else {
                reinterpret_cast<uint16_t *>(w->block + w->off_len_ids)[g] = (uint16_t)c.n_ids[k];
                reinterpret_cast<uint16_t *>(w->block + w->off_len_bytes)[g] = (uint16_t)c.n_bytes[k];
            }
        }
        if (!c.stream.empty()) memcpy(w->block + w->off_stream + spos, c.stream.data(), c.stream.size());
        spos += c.stream.size();
    }
    // slice table: cut at contig starts into parts of about equal bytes moved (stream + lengths in, 8 per gene out).
    // ONE slice unless GCRF_WIRE_SLICES asks for more: on the PCIe Gen5 hosts measured the copy back of slice k did not
    // overlap the copy in of slice k+1 to any effect (config 2: 965 / 967 / 950 / 856 M genes/s with 1 / 2 / 4 / 8
    // slices, profiles/r2_e2e_wire_slices.txt) and every slice costs ~25 us of extra launches and copies
    {
        std::vector<int64_t> gene_id((size_t)G + 1), gene_byte((size_t)G + 1);
        int64_t ids = 0, bytes = 0, gg = 0;
        for (const auto &c : chunks)
            for (size_t k = 0; k < c.n_ids.size(); ++k, ++gg) {
                gene_id[gg] = ids;
                gene_byte[gg] = bytes;
                ids += c.n_ids[k];
                bytes += c.n_bytes[k];
            }
        gene_id[G] = ids;
        gene_byte[G] = bytes;
        auto cost_at = [&](int64_t c) -> double { const int64_t g = contig_ptr[c]; return (double)gene_byte[g] + (2.0 * lw + 8.0) * (double)g; };
        const double total_cost = G > 0 ? cost_at(C) : 0.0;
        int n = 1;
        if (const char *env = getenv("GCRF_WIRE_SLICES")) n = atoi(env);
        if (n > gcrf_wire::kMaxSlices) n = gcrf_wire::kMaxSlices;
        if (n > C) n = (int)C;
        if (n < 1) n = 1;
        w->n_slices = n;
        for (int k = 0; k <= n; ++k) {
            int64_t c = k == n ? C : 0;
            if (k > 0 && k < n) {
                int64_t lo = w->s_contig[k - 1], hi = C;
                const double want = total_cost * k / n;
                while (lo < hi) {
                    const int64_t mid = (lo + hi) / 2;
                    if (cost_at(mid) < want) lo = mid + 1; else hi = mid;
                }
                c = lo;
            }
            w->s_contig[k] = c;
            w->s_gene[k] = G > 0 ? contig_ptr[c] : 0;
            w->s_id[k] = gene_id[w->s_gene[k]];
            w->s_byte[k] = gene_byte[w->s_gene[k]];
        }
    }
    *out = w;
    return GCRF_OK;
}

void gcrf_wire_destroy(gcrf_wire *w) {
    if (!w) return;
    if (w->block) {
        if (w->len_width < 0) free(w->block);
        else cudaFreeHost(w->block);
    }
    delete w;
}

int64_t gcrf_wire_bytes(const gcrf_wire *w) { return w ? (int64_t)w->total : 0; }
int64_t gcrf_wire_contigs(const gcrf_wire *w) { return w ? w->C : 0; }
int64_t gcrf_wire_genes(const gcrf_wire *w) { return w ? w->G : 0; }
int64_t gcrf_wire_ids(const gcrf_wire *w) { return w ? w->nnz : 0; }

int gcrf_wire_decode_host(const gcrf_wire *w, int32_t *gene_ptr, int32_t *attr_idx) {
    if (!w || (w->G > 0 && !gene_ptr) || (w->nnz > 0 && !attr_idx)) return wire_fail(GCRF_EINVAL, "NULL argument");
    const int lw = w->len_width < 0 ? -w->len_width : w->len_width;
    const uint8_t *src = reinterpret_cast<const uint8_t *>(w->block + w->off_stream);
    int64_t p = 0;
    for (int64_t g = 0; g < w->G; ++g) {
        const uint32_t n = lw == 1 ? reinterpret_cast<const uint8_t *>(w->block + w->off_len_ids)[g]
                                   : reinterpret_cast<const uint16_t *>(w->block + w->off_len_ids)[g];
        const uint32_t nb = lw == 1 ? reinterpret_cast<const uint8_t *>(w->block + w->off_len_bytes)[g]
                                    : reinterpret_cast<const uint16_t *>(w->block + w->off_len_bytes)[g];
        gene_ptr[g] = (int32_t)p;
        const uint8_t *end = src + nb;
        int32_t prev = 0;
        for (uint32_t i = 0; i < n; ++i) {
            uint32_t v = 0, byte;
            int shift = 0;
            do {
                byte = *src++;
                v |= (byte & 127u) << shift;
                shift += 7;
            } while (byte & 128u);
            prev += (int32_t)v;
            attr_idx[p++] = prev;
        }
        if (src != end) return wire_fail(GCRF_EINVAL, "corrupt wire block");
    }
    if (w->G > 0) gene_ptr[w->G] = (int32_t)p;
    return p == w->nnz ? GCRF_OK : wire_fail(GCRF_EINVAL, "corrupt wire block");
}

}  // extern "C"

// accessors for gcrf_abi.cu (the marginal entry point lives there, next to the model handle)
namespace gcrf {
const char *wire_block(const gcrf_wire *w) { return w->block; }
size_t wire_total(const gcrf_wire *w) { return w->total; }
size_t wire_off_len_ids(const gcrf_wire *w) { return w->off_len_ids; }
size_t wire_off_len_bytes(const gcrf_wire *w) { return w->off_len_bytes; }
size_t wire_off_stream(const gcrf_wire *w) { return w->off_stream; }
int32_t wire_len_width(const gcrf_wire *w) { return w->len_width < 0 ? -w->len_width : w->len_width; }
int wire_slices(const gcrf_wire *w) { return w->n_slices; }
void wire_slice(const gcrf_wire *w, int k, int64_t *contig, int64_t *gene, int64_t *id, int64_t *byte) {
    *contig = w->s_contig[k];
    *gene = w->s_gene[k];
    *id = w->s_id[k];
    *byte = w->s_byte[k];
}
int64_t wire_stream_bytes(const gcrf_wire *w) { return w->stream_bytes; }
}  // namespace gcrf
