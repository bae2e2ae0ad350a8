// gcrf_wire.cu — a compact wire format for host-buffer calls: what gcrf_marginals_windowed moves over PCIe is the
// whole cost of such a call (the kernels are ~2 % of it), so the bytes are the lever.
//
// The CSR batch of include/gecco_crf_b200.h costs 4 bytes per attribute id, 4 per gene (row pointer) and 8 per gene
// (float64 marginal).  A gcrf_wire holds the same batch as
//     contig_ptr[C+1]   int32, unchanged (4 bytes per contig)
//     len_ids[G]        ids of gene g            uint8, or uint16 when a gene has more than 255 of them
//     len_bytes[G]      bytes of gene g's stream (same width)
//     stream[]          per gene: its attribute ids SORTED ascending, unknown ids (outside [0, A)) mapped to A, as
//                       deltas (first id absolute) in a Rice code, padded to a whole byte per gene
// in ONE page-locked block (one section per slice, see struct gcrf_wire).  Sorting is free for the result: a gene's
// features are a set (gecco/crf/features.py:32) and the device forms row sums in exact integer arithmetic, so the
// order of a row does not matter (rows of >= ModelDev::fx_nsafe ids, which take the float path, may differ in the last
// bit).
//
// The code of one delta v, bits written LSB first, k chosen per batch from the mean delta (k ~ log2(mean ln 2)):
//     q = v >> k < 8:   q ones, a zero, the k low bits of v             (q + 1 + k bits)
//     otherwise:        eight ones, v in 24 bits                        (32 bits; hence A < 2^24)
// Gaps between sorted uniform ids are geometric, for which this is within a few percent of the entropy: config 2
// (Poisson(25) ids per gene out of 2,659, k = 6) costs 1.03 bytes per id + 2.5 per gene instead of 4 + 4 (LEB128
// deltas, this format's first version: 1.30 per id).  A code never exceeds 32 bits, so a decoder keeps a 64-bit window
// and refills it with aligned 32-bit words.
//
// On the device ONE kernel rebuilds gene_ptr and attr_idx (the encoder wrote the sums in front of every block of
// genes; scan within the block, then one thread per gene decoding in shared memory), then the marginal kernels run
// unchanged.
#include "../../include/gecco_crf_b200.h"
#include "gcrf_kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <thread>
#include <vector>

struct gcrf_wire {
    int64_t C = 0, G = 0, nnz = 0, stream_bytes = 0;
    int32_t A = 0;
    int32_t rice_k = 0;     // parameter of the delta code (header comment)
    int32_t len_width = 1;  // bytes per entry of len_ids / len_bytes: 1 or 2 (negative: the block came from malloc)
    // One page-locked block, every part 16-byte aligned:
    //   [contig_ptr int32[C+1] | chunk sums int64[2 * chunks] | section of slice 0 | section of slice 1 | ...]
    //   section of a slice = [len_ids | len_bytes | its stretch of the delta stream]
    // The head (contig_ptr + sums) and every section travel as ONE copy each.  The chunk sums are what the decoder's
    // blocks start from: ids and stream bytes in front of every kChunk-th gene of the slice, relative to the slice.
    char *block = nullptr;
    size_t off_sums = 0, total = 0;
    // contig-aligned slices (the bulk call pipelines them: copy in / decode + kernels / copy back)
    static constexpr int kMaxSlices = 8;
    int n_slices = 1;
    int64_t s_contig[kMaxSlices + 1] = {}, s_gene[kMaxSlices + 1] = {}, s_id[kMaxSlices + 1] = {}, s_byte[kMaxSlices + 1] = {};
    size_t s_off[kMaxSlices] = {}, s_size[kMaxSlices] = {};
    int64_t s_chunk[kMaxSlices + 1] = {};  // index of the slice's first pair of chunk sums
};

namespace gcrf {

namespace {

constexpr int kWireThreads = 256, kWirePerThread = 2, kWireChunk = kWireThreads * kWirePerThread;  // genes per block

inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// ---- the decoder: offsets of every gene inside its block (scan), gene_ptr, and the ids themselves.  `sums` holds, per
// block, the ids and stream bytes in front of it (written by the encoder: no scan over the whole batch on the device).
// The block's stretch of the byte stream is contiguous, and so are the ids it produces: both go through shared
// memory in rounds (as many genes as fit the two staging areas), so that global memory sees 16-byte loads and
// coalesced 4-byte stores only; the serial walk of one gene per thread runs on shared memory.  (The first
// version walked global memory directly: 0.41 ms for BASELINE config 2, five times this one.)
#ifndef GCRF_WIRE_CAP_IDS
#define GCRF_WIRE_CAP_IDS 8192
#endif
constexpr int kWireCapIds = GCRF_WIRE_CAP_IDS;            // ids staged per round
constexpr int kWireCapBytes = GCRF_WIRE_CAP_IDS * 3 / 2;  // stream bytes staged per round (plus the 16-byte alignment slack)
constexpr int kWireInSlack = 32;       // 16 for the alignment of the first word, 11 + for the look-ahead of decode_gene
constexpr size_t kWireDecodeSmem = (size_t)(2 * (kWireChunk + 1)) * sizeof(int) + (size_t)kWireCapIds * 4 + kWireCapBytes + kWireInSlack + 16;

// n ids of one gene from its stretch of the stream (any address space; reads aligned 32-bit words, up to 11 bytes past
// the gene's last byte — the sections and the staging area carry that slack).  Returns the bits consumed.
__host__ __device__ inline int64_t decode_gene(const uint8_t *src, int n, int k, int32_t *dst) {
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
    const uint32_t *next = reinterpret_cast<const uint32_t *>(src - mis);
    unsigned long long w = ((unsigned long long)next[1] << 32 | next[0]) >> (8 * mis);
    int avail = 64 - 8 * (int)mis;
    next += 2;
    const uint32_t kmask = (1u << k) - 1u;
    int32_t prev = 0;
    int64_t bits = 0;
    for (int i = 0; i < n; ++i) {
        const uint32_t x = (uint32_t)w;
        uint32_t v;
        int len;
        if ((x & 0xffu) != 0xffu) {
#ifdef __CUDA_ARCH__
            const int q = __ffs((int)~x) - 1;  // ones in front of the first zero
#else
            const int q = __builtin_ctz(~x);
#endif
            len = q + 1 + k;
            v = ((uint32_t)q << k) | ((uint32_t)(w >> (q + 1)) & kmask);
        } else {
            len = 32;
            v = x >> 8;
        }
        w >>= len;
        avail -= len;
        bits += len;
        if (avail <= 32) {
            w |= (unsigned long long)*next++ << avail;
            avail += 32;
        }
        prev += (int32_t)v;
        dst[i] = prev;
    }
    return bits;
}

template <typename LenT>
__global__ void __launch_bounds__(kWireThreads)
wire_decode_kernel(const LenT *__restrict__ len_ids, const LenT *__restrict__ len_bytes, const uint8_t *__restrict__ stream,
                   int64_t G, int64_t id_base, int rice_k, const int64_t *__restrict__ sums, int32_t *__restrict__ gene_ptr,
                   int32_t *__restrict__ attr_idx) {
    extern __shared__ __align__(16) unsigned char sDyn[];
    int32_t *sOut = reinterpret_cast<int32_t *>(sDyn);                    // [kWireCapIds]
    uint8_t *sIn = sDyn + (size_t)kWireCapIds * 4;                        // [kWireCapBytes + kWireInSlack]
    int *sIds = reinterpret_cast<int *>(sIn + kWireCapBytes + kWireInSlack);  // [kWireChunk + 1] lengths, then exclusive offsets
    int *sBytes = sIds + kWireChunk + 1;                                  // [kWireChunk + 1]
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
    // thread t scans its kWirePerThread consecutive entries, then the thread totals are scanned across the block
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
    __syncthreads();  // everybody has read the lengths of its entries
#pragma unroll
    for (int k = 0; k < kWirePerThread; ++k) {
        sIds[tid * kWirePerThread + k] = wa + ia - ta + a[k];
        sBytes[tid * kWirePerThread + k] = wb + ib - tb + b[k];
    }
    if (tid == kWireThreads - 1) {  // the sentinel: totals of the block
        sIds[kWireChunk] = wa + ia;
        sBytes[kWireChunk] = wb + ib;
    }
    __syncthreads();
    const int64_t id0 = id_base + sums[2 * (int64_t)blockIdx.x], byte0 = sums[2 * (int64_t)blockIdx.x + 1];
    const int nG = (int)(G - base < kWireChunk ? G - base : kWireChunk);
    for (int j = tid; j < nG; j += kWireThreads) gene_ptr[base + j] = (int32_t)(id0 + sIds[j]);
    if (base + nG == G && tid == 0) gene_ptr[G] = (int32_t)(id0 + sIds[nG]);

    int s = 0;
    while (s < nG) {  // every quantity that steers the loop is the same in all threads
        const int i0 = sIds[s], b0 = sBytes[s];
        // the largest e <= nG whose genes [s, e) fit both staging areas (the offsets are non-decreasing)
        int lo = s, hi = nG;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (sIds[mid] - i0 <= kWireCapIds && sBytes[mid] - b0 <= kWireCapBytes) lo = mid;
            else hi = mid - 1;
        }
        const int e = lo;
        if (e == s) {  // one gene larger than a staging area (thousands of domains): straight from and to global memory
            if (tid == 0) decode_gene(stream + byte0 + b0, sIds[s + 1] - i0, rice_k, attr_idx + id0 + i0);
            s += 1;
            continue;
        }
        const int nbytes = sBytes[e] - b0, nids = sIds[e] - i0;
        const uint8_t *src = stream + byte0 + b0;
        const int mis = (int)(reinterpret_cast<uintptr_t>(src) & 15u);
        const uint4 *src16 = reinterpret_cast<const uint4 *>(src - mis);
        const int n16 = (mis + nbytes + 15) >> 4;
        for (int k = tid; k < n16; k += kWireThreads) reinterpret_cast<uint4 *>(sIn)[k] = __ldg(src16 + k);
        __syncthreads();
        for (int j = s + tid; j < e; j += kWireThreads)
            decode_gene(sIn + mis + (sBytes[j] - b0), sIds[j + 1] - sIds[j], rice_k, sOut + (sIds[j] - i0));
        __syncthreads();
        int32_t *dst = attr_idx + id0 + i0;
        for (int k = tid; k < nids; k += kWireThreads) dst[k] = sOut[k];
        __syncthreads();
        s = e;
    }
}

}  // namespace

int64_t wire_chunks(int64_t G) { return (G + kWireChunk - 1) / kWireChunk; }

cudaError_t launch_wire_decode(const void *len_ids, const void *len_bytes, int32_t len_width, const uint8_t *stream, int64_t G,
                               int64_t id_base, int32_t rice_k, const int64_t *sums, int32_t *gene_ptr, int32_t *attr_idx,
                               cudaStream_t cuda_stream, int64_t *launches) {
    if (G <= 0) return cudaSuccess;
    const int64_t nb = wire_chunks(G);
    auto decode = [&](auto kernel, auto *ids, auto *bytes) -> cudaError_t {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWireDecodeSmem);
        if (err != cudaSuccess) return err;
        kernel<<<(unsigned)nb, kWireThreads, kWireDecodeSmem, cuda_stream>>>(ids, bytes, stream, G, id_base, (int)rice_k, sums, gene_ptr, attr_idx);
        return cudaSuccess;
    };
    cudaError_t err;
    if (len_width == 1) err = decode(wire_decode_kernel<uint8_t>, static_cast<const uint8_t *>(len_ids), static_cast<const uint8_t *>(len_bytes));
    else err = decode(wire_decode_kernel<uint16_t>, static_cast<const uint16_t *>(len_ids), static_cast<const uint16_t *>(len_bytes));
    if (err != cudaSuccess) return err;
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace gcrf

// ---------------------------------------------------------------------------------------------------------------------
// host side: encoder, accessors, host decoder (the CPU test-suite checks the format without a GPU)
// ---------------------------------------------------------------------------------------------------------------------

namespace {

struct EncodedChunk {
    std::unique_ptr<uint8_t[]> stream;  // not value-initialised: only the bytes written are ever touched
    size_t stream_size = 0;
    std::vector<uint32_t> n_ids, n_bytes;
    uint32_t longest = 0;  // largest entry of n_ids / n_bytes
};

// One range of genes -> their stream bytes and the two length arrays.  Written for speed (the encoder runs once per batch
// in front of a call that takes a millisecond): rows that arrive sorted — what the packers emit — skip the sort, the
// code of a delta is assembled in a register and leaves with one 8-byte store.
template <typename PtrT>
void encode_range(const PtrT *gene_ptr, const int32_t *attr_idx, int64_t g0, int64_t g1, int32_t A, int k, EncodedChunk *out) {
    const size_t genes = (size_t)(g1 - g0), ids = (size_t)(gene_ptr[g1] - gene_ptr[g0]);
    out->n_ids.resize(genes);
    out->n_bytes.resize(genes);
    out->stream.reset(new uint8_t[ids * 4 + genes + 16]);  // a code is at most 32 bits, a gene ends with at most one padding byte
    uint8_t *base = out->stream.get();
    uint32_t longest = 0;
    size_t pos = 0;
    std::vector<uint32_t> row;
    const uint32_t kmask = (1u << k) - 1u, top = (uint32_t)A;
    for (int64_t g = g0; g < g1; ++g) {
        const int64_t p0 = (int64_t)gene_ptr[g], n = (int64_t)gene_ptr[g + 1] - p0;
        row.resize((size_t)n);
        bool sorted = true;
        uint32_t last = 0;
        for (int64_t i = 0; i < n; ++i) {
            uint32_t a = (uint32_t)attr_idx[p0 + i];
            a = a < top ? a : top;  // every unknown id becomes the zero slot A
            sorted &= a >= last;
            last = a;
            row[(size_t)i] = a;
        }
        if (!sorted) std::sort(row.begin(), row.end());
        const size_t start = pos;
        unsigned long long acc = 0;
        int nb = 0;  // bits waiting in acc, < 8 between codes
        uint32_t prev = 0;
        for (uint32_t a : row) {
            const uint32_t v = a - prev, q = v >> k;
            prev = a;
            uint32_t code;
            int len;
            if (q < 8) {  // q ones, a zero, the k low bits
                code = ((1u << q) - 1u) | ((v & kmask) << (q + 1));
                len = (int)q + 1 + k;
            } else {      // eight ones, 24 raw bits
                code = 0xffu | (v << 8);
                len = 32;
            }
            acc |= (unsigned long long)code << nb;
            nb += len;
            memcpy(base + pos, &acc, 8);  // little-endian hosts; the bytes behind the valid ones are rewritten later
            const int whole = nb >> 3;
            pos += (size_t)whole;
            acc >>= 8 * whole;
            nb &= 7;
        }
        if (nb) base[pos++] = (uint8_t)(acc & 0xff);
        out->n_ids[(size_t)(g - g0)] = (uint32_t)n;
        out->n_bytes[(size_t)(g - g0)] = (uint32_t)(pos - start);
        longest = std::max(longest, std::max((uint32_t)n, (uint32_t)(pos - start)));
    }
    out->stream_size = pos;
    out->longest = longest;
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
    if (num_attrs >= (1 << 24)) return wire_fail(GCRF_EUNSUPPORTED, "the wire format codes ids in at most 24 bits: fewer than 2^24 attributes");
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
    // the Rice parameter from the mean delta = (sum over genes of their largest id) / ids
    std::vector<double> top_sum(nthreads, 0.0);
    auto each_range = [&](auto body) {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nthreads; ++t) {
            const int64_t g0 = G * t / nthreads, g1 = G * (t + 1) / nthreads;
            if (t + 1 == nthreads) body(t, g0, g1);
            else pool.emplace_back(body, t, g0, g1);
        }
        for (auto &th : pool) th.join();
    };
    each_range([&](unsigned t, int64_t g0, int64_t g1) {
        double sum = 0;
        for (int64_t g = g0; g < g1; ++g) {
            uint32_t top = 0;
            for (int64_t p = row(g); p < row(g + 1); ++p) {
                const uint32_t a = (uint32_t)attr_idx[p];
                top = std::max(top, a < (uint32_t)num_attrs ? a : (uint32_t)num_attrs);
            }
            sum += top;
        }
        top_sum[t] = sum;
    });
    double mean_delta = 0;
    for (double v : top_sum) mean_delta += v;
    mean_delta = nnz > 0 ? mean_delta / (double)nnz : 0.0;
    int rice_k = 0;
    while (rice_k < 20 && (double)(1u << (rice_k + 1)) <= mean_delta * 0.6931 * 1.4142) ++rice_k;  // round(log2(mean ln 2))
    std::vector<EncodedChunk> chunks(nthreads);
    each_range([&](unsigned t, int64_t g0, int64_t g1) {
        if (ptr64) encode_range(static_cast<const int64_t *>(gene_ptr), attr_idx, g0, g1, num_attrs, rice_k, &chunks[t]);
        else encode_range(static_cast<const int32_t *>(gene_ptr), attr_idx, g0, g1, num_attrs, rice_k, &chunks[t]);
    });

    uint32_t longest = 0;
    int64_t stream_bytes = 0;
    for (const auto &c : chunks) {
        longest = std::max(longest, c.longest);
        stream_bytes += (int64_t)c.stream_size;
    }
    if (longest > 0xFFFF) return wire_fail(GCRF_EUNSUPPORTED, "a gene with more than 65535 ids / stream bytes does not fit the wire format");
    gcrf_wire *w = new (std::nothrow) gcrf_wire();
    if (!w) return wire_fail(GCRF_ENOMEM, "out of host memory");
    w->C = C; w->G = G; w->nnz = nnz; w->A = num_attrs; w->stream_bytes = stream_bytes;
    w->rice_k = rice_k;
    const int lw = longest > 0xFF ? 2 : 1;
    w->len_width = lw;

    // ids and stream bytes in front of every gene; where every encoded range starts in the stream
    std::vector<int64_t> gene_id((size_t)G + 1), gene_byte((size_t)G + 1), range_byte(chunks.size() + 1), range_id(chunks.size() + 1);
    {
        int64_t ids = 0, bytes = 0;
        for (size_t t = 0; t < chunks.size(); ++t) {
            range_id[t] = ids;
            range_byte[t] = bytes;
            ids += (int64_t)row(G * (int64_t)(t + 1) / nthreads) - (int64_t)row(G * (int64_t)t / nthreads);
            bytes += (int64_t)chunks[t].stream_size;
        }
        range_id[chunks.size()] = ids;
        range_byte[chunks.size()] = bytes;
        gene_id[G] = ids;
        gene_byte[G] = bytes;
        each_range([&](unsigned t, int64_t g0, int64_t g1) {
            int64_t i = range_id[t], b = range_byte[t];
            for (int64_t g = g0; g < g1; ++g) {
                gene_id[g] = i;
                gene_byte[g] = b;
                i += chunks[t].n_ids[(size_t)(g - g0)];
                b += chunks[t].n_bytes[(size_t)(g - g0)];
            }
        });
    }
    // Slice table: cut at contig starts.  A bulk call runs the slices as a three-stage pipeline (copy in / decode +
    // kernels / copy back, gcrf_marginals_windowed_wire); the copy in is the long pole, so what the pipeline adds to it
    // is the LAST slice's kernels and copy back: the slices shrink geometrically (8 : 4 : 2 : 1 of the bytes moved —
    // stream + lengths in, 8 per gene out), each still long enough to cover its predecessor's kernels (~0.2 of its copy).
    // Numbers: profiles/r2_e2e_wire_slices.txt.
    {
        auto cost_at = [&](int64_t c) -> double { const int64_t g = contig_ptr[c]; return (double)gene_byte[g] + (2.0 * lw + 8.0) * (double)g; };
        const double total_cost = G > 0 ? cost_at(C) : 0.0;
        int n = total_cost >= 24e6 ? 4 : total_cost >= 8e6 ? 2 : 1;
        if (const char *env = getenv("GCRF_WIRE_SLICES")) n = atoi(env);
        if (n > gcrf_wire::kMaxSlices) n = gcrf_wire::kMaxSlices;
        if (n > C) n = (int)C;
        if (n < 1) n = 1;
        w->n_slices = n;
        for (int k = 0; k <= n; ++k) {
            int64_t c = k == n ? C : 0;
            if (k > 0 && k < n) {
                int64_t lo = w->s_contig[k - 1], hi = C;
                const double want = total_cost * (1.0 - (double)((1 << (n - k)) - 1) / (double)((1 << n) - 1));
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
    // layout
    int64_t n_chunks = 0;
    for (int k = 0; k < w->n_slices; ++k) {
        w->s_chunk[k] = n_chunks;
        n_chunks += gcrf::wire_chunks(w->s_gene[k + 1] - w->s_gene[k]);
    }
    w->s_chunk[w->n_slices] = n_chunks;
    w->off_sums = gcrf::align16((size_t)(C + 1) * 4);
    size_t pos = gcrf::align16(w->off_sums + (size_t)n_chunks * 2 * sizeof(int64_t));
    for (int k = 0; k < w->n_slices; ++k) {
        const size_t genes = (size_t)(w->s_gene[k + 1] - w->s_gene[k]), bytes = (size_t)(w->s_byte[k + 1] - w->s_byte[k]);
        w->s_off[k] = pos;
        w->s_size[k] = 2 * gcrf::align16(genes * lw) + gcrf::align16(bytes + 16);  // the decoder reads whole 16-byte words
        pos += w->s_size[k];
    }
    w->total = pos;
    if (cudaHostAlloc(reinterpret_cast<void **>(&w->block), w->total, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        // no CUDA runtime / device on this host: plain memory keeps the encoder usable (the copy is then pageable)
        w->block = static_cast<char *>(malloc(w->total));
        if (!w->block) {
            delete w;
            return wire_fail(GCRF_ENOMEM, "out of host memory");
        }
        w->len_width = -lw;  // negative: block came from malloc
    }
    // fill: head, then every encoded range writes its genes' lengths and its stretch of the stream into the sections of
    // the slices it overlaps (all threads); the few padding bytes between the parts are zeroed explicitly
    if (G > 0) memcpy(w->block, contig_ptr, (size_t)(C + 1) * 4);
    memset(w->block + (size_t)(C + 1) * 4, 0, w->off_sums - (size_t)(C + 1) * 4);
    int64_t *sums = reinterpret_cast<int64_t *>(w->block + w->off_sums);
    for (int k = 0; k < w->n_slices; ++k) {
        const int64_t g0 = w->s_gene[k], g1 = w->s_gene[k + 1];
        for (int64_t j = 0, g = g0; g < g1; ++j, g += gcrf::kWireChunk) {
            sums[2 * (w->s_chunk[k] + j)] = gene_id[g] - gene_id[g0];
            sums[2 * (w->s_chunk[k] + j) + 1] = gene_byte[g] - gene_byte[g0];
        }
        // padding of the section: behind each length array and behind the stream
        const size_t part = gcrf::align16((size_t)(g1 - g0) * lw), used = (size_t)(g1 - g0) * lw;
        char *sec = w->block + w->s_off[k];
        memset(sec + used, 0, part - used);
        memset(sec + part + used, 0, part - used);
        const size_t sbytes = (size_t)(w->s_byte[k + 1] - w->s_byte[k]);
        memset(sec + 2 * part + sbytes, 0, w->s_size[k] - 2 * part - sbytes);
    }
    {
        const size_t head_end = w->off_sums + (size_t)n_chunks * 2 * sizeof(int64_t);
        const size_t first = w->n_slices > 0 ? w->s_off[0] : w->total;
        if (first > head_end) memset(w->block + head_end, 0, first - head_end);
    }
    each_range([&](unsigned t, int64_t ga, int64_t gb) {
        int k = 0;
        for (int64_t g = ga; g < gb;) {
            while (g >= w->s_gene[k + 1]) ++k;  // the slice of gene g
            const int64_t g0 = w->s_gene[k], g1 = w->s_gene[k + 1], ge = gb < g1 ? gb : g1;
            const size_t part = gcrf::align16((size_t)(g1 - g0) * lw);
            char *sec = w->block + w->s_off[k];
            if (lw == 1) {
                uint8_t *li = reinterpret_cast<uint8_t *>(sec), *lb = reinterpret_cast<uint8_t *>(sec + part);
                for (int64_t x = g; x < ge; ++x) {
                    li[x - g0] = (uint8_t)chunks[t].n_ids[(size_t)(x - ga)];
                    lb[x - g0] = (uint8_t)chunks[t].n_bytes[(size_t)(x - ga)];
                }
            } else {
                uint16_t *li = reinterpret_cast<uint16_t *>(sec), *lb = reinterpret_cast<uint16_t *>(sec + part);
                for (int64_t x = g; x < ge; ++x) {
                    li[x - g0] = (uint16_t)chunks[t].n_ids[(size_t)(x - ga)];
                    lb[x - g0] = (uint16_t)chunks[t].n_bytes[(size_t)(x - ga)];
                }
            }
            // this range's bytes of genes [g, ge) are contiguous in its stream and in the section
            const int64_t b_lo = gene_byte[g], b_hi = gene_byte[ge];
            if (b_hi > b_lo)
                memcpy(sec + 2 * part + (size_t)(b_lo - w->s_byte[k]), chunks[t].stream.get() + (size_t)(b_lo - range_byte[t]), (size_t)(b_hi - b_lo));
            g = ge;
        }
    });
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
    int64_t p = 0;
    for (int k = 0; k < w->n_slices; ++k) {
        const int64_t g0 = w->s_gene[k], g1 = w->s_gene[k + 1];
        const char *sec = w->block + w->s_off[k];
        const char *sec_bytes = sec + gcrf::align16((size_t)(g1 - g0) * lw);
        const uint8_t *src = reinterpret_cast<const uint8_t *>(sec_bytes + gcrf::align16((size_t)(g1 - g0) * lw));
        if (p != w->s_id[k]) return wire_fail(GCRF_EINVAL, "corrupt wire block");
        for (int64_t g = g0; g < g1; ++g) {
            const uint32_t n = lw == 1 ? reinterpret_cast<const uint8_t *>(sec)[g - g0] : reinterpret_cast<const uint16_t *>(sec)[g - g0];
            const uint32_t nb = lw == 1 ? reinterpret_cast<const uint8_t *>(sec_bytes)[g - g0] : reinterpret_cast<const uint16_t *>(sec_bytes)[g - g0];
            gene_ptr[g] = (int32_t)p;
            if (p + (int64_t)n > w->nnz) return wire_fail(GCRF_EINVAL, "corrupt wire block");
            const int64_t bits = gcrf::decode_gene(src, (int)n, w->rice_k, attr_idx + p);
            p += n;
            if ((bits + 7) / 8 != (int64_t)nb) return wire_fail(GCRF_EINVAL, "corrupt wire block");
            src += nb;
        }
    }
    if (w->G > 0) gene_ptr[w->G] = (int32_t)p;
    return p == w->nnz ? GCRF_OK : wire_fail(GCRF_EINVAL, "corrupt wire block");
}

}  // extern "C"

// accessors for gcrf_abi.cu (the marginal entry point lives there, next to the model handle)
namespace gcrf {
const char *wire_block(const gcrf_wire *w) { return w->block; }
size_t wire_total(const gcrf_wire *w) { return w->total; }
size_t wire_head_bytes(const gcrf_wire *w) { return w->n_slices > 0 ? w->s_off[0] : w->total; }  // contig_ptr + chunk sums
size_t wire_off_sums(const gcrf_wire *w) { return w->off_sums; }
void wire_section(const gcrf_wire *w, int k, size_t *off, size_t *size, size_t *rel_len_bytes, size_t *rel_stream, int64_t *first_chunk) {
    const int lw = w->len_width < 0 ? -w->len_width : w->len_width;
    const size_t part = align16((size_t)(w->s_gene[k + 1] - w->s_gene[k]) * lw);
    *off = w->s_off[k];
    *size = w->s_size[k];
    *rel_len_bytes = part;
    *rel_stream = 2 * part;
    *first_chunk = w->s_chunk[k];
}
int32_t wire_len_width(const gcrf_wire *w) { return w->len_width < 0 ? -w->len_width : w->len_width; }
int wire_slices(const gcrf_wire *w) { return w->n_slices; }
int32_t wire_rice_k(const gcrf_wire *w) { return w->rice_k; }
void wire_slice(const gcrf_wire *w, int k, int64_t *contig, int64_t *gene, int64_t *id, int64_t *byte) {
    *contig = w->s_contig[k];
    *gene = w->s_gene[k];
    *id = w->s_id[k];
    *byte = w->s_byte[k];
}
}  // namespace gcrf
