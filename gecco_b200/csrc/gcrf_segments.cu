// gcrf_segments.cu — threshold + contiguous-segment extraction on the marginal kernel's output (next row N1,
// SURVEY.md §8(f)): the array form of gecco.refine.ClusterRefiner with criterion "gecco"
// (gecco/refine.py:51-64 GeneGrouper, :183-200 _iter_clusters, :167-180 _trim_cluster, :139-165 _validate_cluster).
//
// The reference walks the genes in order with one piece of state (GeneGrouper.in_cluster, which survives contig
// boundaries because the grouper object is created once, :190).  Here the walk is three chunked scans over the
// gene axis, each two launches of a persistent grid (chunk summaries, then chunk replay with the carry-in folded
// from the summaries):
//
//   round 1  in-cluster flag        max-scan of (position, p > threshold) over the genes that HAVE a probability
//                                   (a gene without one — NaN — inherits the state of the last gene that has)
//   round 2  annotated-gene count   sum-scan (exclusive prefix + compaction array of annotated positions),
//            raw-run ordinal        sum-scan of run ends (the reference numbers clusters per contig BEFORE validation),
//            run start              max-scan of "latest breaker": an out-of-cluster gene or a contig start,
//            contig index           sum-scan of contig starts
//            -> one RECORD per run (last gene, first gene, contig), written where the run ends, and the number of
//               runs that ended before each contig
//   round 3  valid-cluster count    sum-scan over the RUNS (not the genes) -> clusters are written in the reference's
//            order, no atomics, no sort
//
// Every run evaluates trim + validation in O(1) from its record and the annotated-gene prefix.
// HBM-bound integer/byte work: 9 B/gene read in round 1, ~6 B/gene of scratch written in round 2, the rest is per run.
#include "gcrf_kernels.cuh"

namespace gcrf {

namespace {

constexpr int kThreads = 256;
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;

struct SumMax {
    int32_t ann, ends, brk, cst;  // annotated genes, run ends, latest breaker position (max), contig starts
};
__device__ __forceinline__ SumMax combine(const SumMax &a, const SumMax &b) {
    return SumMax{a.ann + b.ann, a.ends + b.ends, max(a.brk, b.brk), a.cst + b.cst};
}
__device__ __forceinline__ uint32_t combine(uint32_t a, uint32_t b) { return max(a, b); }
__device__ __forceinline__ int32_t combine(int32_t a, int32_t b) { return a + b; }

__device__ __forceinline__ SumMax shfl_up(const SumMax &v, int d) {
    return SumMax{__shfl_up_sync(0xffffffffu, v.ann, d), __shfl_up_sync(0xffffffffu, v.ends, d),
                  __shfl_up_sync(0xffffffffu, v.brk, d), __shfl_up_sync(0xffffffffu, v.cst, d)};
}
__device__ __forceinline__ uint32_t shfl_up(uint32_t v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ int32_t shfl_up(int32_t v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }

// Exclusive scan of one value per thread across the CTA; *total = combination of all of them.  `ident` is the
// identity of combine() for T.  sWarp: kThreads/32 + 1 elements of shared scratch.
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T ident, T *sWarp, T *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const T o = shfl_up(inc, d);
        if (lane >= d) inc = combine(o, inc);
    }
    __syncthreads();  // sWarp may still be read from the previous call
    if (lane == 31) sWarp[warp] = inc;
    __syncthreads();
    T base = ident;
    T all = ident;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
        const T t = sWarp[w];
        if (w < warp) base = combine(base, t);
        all = combine(all, t);
    }
    *total = all;
    T exc = shfl_up(inc, 1);
    if (lane == 0) exc = ident;
    return combine(base, exc);
}

// Combination of partial[0 .. n) (n <= a few thousand chunk summaries), by the whole CTA.
template <typename T>
__device__ __forceinline__ T block_fold(const T *partial, int n, T ident, T *sWarp) {
    T acc = ident;
    // order matters only for non-commutative operators; max and + commute
    for (int i = threadIdx.x; i < n; i += kThreads) acc = combine(acc, partial[i]);
    T total;
    block_exclusive_scan(acc, ident, sWarp, &total);
    return total;
}

struct Geometry {
    int64_t G;
    int64_t chunk;  // genes per CTA, a multiple of kTile
};

__device__ __forceinline__ double load_prob(const void *prob, int f32, int64_t g) {
    return f32 ? (double)__ldg(static_cast<const float *>(prob) + g) : __ldg(static_cast<const double *>(prob) + g);
}

// ---- round 1: key of gene g = 2 (g + 1) + (p > threshold) if p is not NaN, else 0; flag[g] = low bit of the max
//      over [0, g] (0 when no gene so far had a probability: GeneGrouper starts with in_cluster = False, :56)
//      With reset_per_contig a gene without probability at a contig start counts as "out of cluster": one
//      iter_clusters call — a fresh GeneGrouper — per contig, as the pipeline does (_common.py:616-618); the
//      genes behind it inherit that.
__device__ __forceinline__ uint32_t state_key(const SegmentsArgs &a, int64_t g) {
    const double p = load_prob(a.prob, a.prob_f32, g);
    if (p != p) {
        if (!a.reset_per_contig) return 0u;
        // nearest defined gene or contig start to the left decides; a start is a defined "False"
        return a.cmark[g] ? (uint32_t)(2 * (g + 1)) : 0u;
    }
    return (uint32_t)(2 * (g + 1)) + (p > a.threshold ? 1u : 0u);
}

template <bool kReplay>
__global__ void __launch_bounds__(kThreads)
state_kernel(const SegmentsArgs a, const Geometry geo, uint32_t *partial) {
    __shared__ uint32_t sWarp[kThreads / 32 + 1];
    const int64_t c0 = (int64_t)blockIdx.x * geo.chunk, c1 = min(geo.G, c0 + geo.chunk);
    uint32_t carry = 0;
    if (kReplay) carry = block_fold(partial, (int)blockIdx.x, 0u, sWarp);
    for (int64_t t0 = c0; t0 < c1; t0 += kTile) {
        const int64_t g0 = t0 + (int64_t)threadIdx.x * kItems;
        uint32_t key[kItems];
        uint32_t mine = 0;
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            key[i] = g0 + i < c1 ? state_key(a, g0 + i) : 0u;
            mine = max(mine, key[i]);
        }
        if (!kReplay) {
            carry = max(carry, mine);  // thread-local; folded once after the loop
        } else {
            uint32_t total;
            uint32_t run = max(carry, block_exclusive_scan(mine, 0u, sWarp, &total));
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
                run = max(run, key[i]);
                if (g0 + i < c1) a.flag[g0 + i] = (uint8_t)(run & 1u);
            }
            carry = max(carry, total);
        }
    }
    if (!kReplay) {
        uint32_t total;
        block_exclusive_scan(carry, 0u, sWarp, &total);
        if (threadIdx.x == 0) partial[blockIdx.x] = total;
    }
}

// contig starts -> byte marks (cmark is zeroed beforehand; cmark[G] = 1 closes the last contig)
__global__ void __launch_bounds__(kThreads) mark_kernel(const int32_t *__restrict__ contig_ptr, int64_t C, uint8_t *cmark) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i <= C) cmark[__ldg(contig_ptr + i)] = 1;
}

// ---- round 2
template <bool kReplay>
__global__ void __launch_bounds__(kThreads)
runs_kernel(const SegmentsArgs a, const Geometry geo, SumMax *partial) {
    __shared__ SumMax sWarp[kThreads / 32 + 1];
    const SumMax ident{0, 0, 0, 0};
    const int64_t c0 = (int64_t)blockIdx.x * geo.chunk, c1 = min(geo.G, c0 + geo.chunk);
    SumMax carry = ident;
    if (kReplay) carry = block_fold(partial, (int)blockIdx.x, ident, sWarp);
    for (int64_t t0 = c0; t0 < c1; t0 += kTile) {
        const int64_t g0 = t0 + (int64_t)threadIdx.x * kItems;
        SumMax item[kItems];
        SumMax mine = ident;
        // eight genes per thread: their flag / annotation / contig-start bytes come as ONE 8-byte load per array
        // (g0 is a multiple of 8 and the arrays are 16-byte aligned; the scratch is padded past G), plus the two
        // bytes of the gene behind them
        static_assert(kItems == 8, "byte arrays are read eight at a time");
        uint64_t fl8 = 0, an8 = 0, cm8 = 0;
        unsigned next_fl = 0, next_cm = 1;
        if (g0 < c1) {
            fl8 = *reinterpret_cast<const uint64_t *>(a.flag + g0);
            cm8 = *reinterpret_cast<const uint64_t *>(a.cmark + g0);
            if (g0 + 8 <= geo.G && (reinterpret_cast<uintptr_t>(a.annotated) & 7) == 0) {  // a caller's device array
                an8 = __ldg(reinterpret_cast<const unsigned long long *>(a.annotated + g0));
            } else {
                for (int i = 0; i < 8 && g0 + i < geo.G; ++i) an8 |= (uint64_t)(a.annotated[g0 + i] != 0) << (8 * i);
            }
            if (g0 + 8 < geo.G) {
                next_fl = a.flag[g0 + 8];
                next_cm = a.cmark[g0 + 8];
            }
        }
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            const int64_t g = g0 + i;
            item[i] = ident;
            if (g < c1) {
                const bool fl = ((fl8 >> (8 * i)) & 0xff) != 0;
                const bool cm = ((cm8 >> (8 * i)) & 0xff) != 0;
                const bool fl_next = i + 1 < kItems ? ((fl8 >> (8 * (i + 1))) & 0xff) != 0 : next_fl != 0;
                const bool cm_next = i + 1 < kItems ? ((cm8 >> (8 * (i + 1))) & 0xff) != 0 : next_cm != 0;
                item[i].ann = ((an8 >> (8 * i)) & 0xff) ? 1 : 0;
                // is_run_end: in a cluster, and the next gene is not, or opens another contig, or does not exist
                item[i].ends = (fl && (g + 1 == geo.G || !fl_next || cm_next)) ? 1 : 0;
                // latest place a run can have started: right after an out-of-cluster gene, or at a contig start
                item[i].brk = !fl ? (int32_t)(g + 1) : (cm ? (int32_t)g : 0);
                item[i].cst = cm ? 1 : 0;
            }
            mine = combine(mine, item[i]);
        }
        if (!kReplay) {
            carry = combine(carry, mine);
        } else {
            SumMax total;
            SumMax run = combine(carry, block_exclusive_scan(mine, ident, sWarp, &total));
            int32_t pre[kItems];
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
                const int64_t g = g0 + i;
                pre[i] = run.ann;  // annotated genes in [0, g)
                if (g < c1) {
                    if (item[i].ann) a.ann_pos[run.ann] = (int32_t)g;
                    if (item[i].cst) a.contig_first_run[run.cst] = run.ends;  // runs that ended before this contig
                }
                const int32_t r = run.ends;  // runs that ended before g: the index of a run ending at g
                run = combine(run, item[i]);
                if (g < c1 && item[i].ends) {
                    a.run_end[r] = (int32_t)g;
                    a.run_start[r] = run.brk;
                    a.run_contig[r] = run.cst - 1;
                }
            }
            if (g0 < c1) {  // two 16-byte stores (the array is padded past G; slots beyond c1 are rewritten by their owner)
                int4 *dst = reinterpret_cast<int4 *>(a.ann_prefix + g0);
                if (g0 + kItems <= c1) {
                    dst[0] = make_int4(pre[0], pre[1], pre[2], pre[3]);
                    dst[1] = make_int4(pre[4], pre[5], pre[6], pre[7]);
                } else {
                    for (int i = 0; g0 + i < c1; ++i) a.ann_prefix[g0 + i] = pre[i];
                }
            }
            carry = combine(carry, total);
            if (t0 + kTile >= c1 && c1 == geo.G && threadIdx.x == 0) {
                a.ann_prefix[geo.G] = carry.ann;
                *a.n_runs = carry.ends;
            }
        }
    }
    if (!kReplay) {
        SumMax total;
        block_exclusive_scan(carry, ident, sWarp, &total);
        if (threadIdx.x == 0) partial[blockIdx.x] = total;
    }
}

// ---- round 3: trim + validate the run that ends at gene g (refine.py:139-180)
struct Segment {
    int32_t contig, begin, end, ordinal;
    bool valid;
};

__device__ __forceinline__ int32_t overlap(int32_t a0, int32_t a1, int32_t b0, int32_t b1) {
    return max(0, min(a1, b1) - max(a0, b0));
}

__device__ Segment evaluate_run(const SegmentsArgs &a, int32_t r) {
    Segment s;
    const int32_t rs = a.run_start[r], re = a.run_end[r] + 1, c = a.run_contig[r];
    int32_t b = rs, t = re;
    int32_t k0 = a.ann_prefix[b], k1 = a.ann_prefix[t];
    if (a.trim) {  // _trim_cluster: genes without domains are dropped from both ends
        if (k1 > k0) {
            b = a.ann_pos[k0];
            t = a.ann_pos[k1 - 1] + 1;
        } else {
            t = b;  // nothing annotated: the cluster is emptied
        }
    }
    // _validate_cluster, criterion "gecco": annotated genes of the cluster >= n_cds, and genes of the cluster that
    // are not among the contig's first / last `edge_distance` ANNOTATED genes >= n_cds
    const int32_t n_annot = k1 - k0;
    int32_t n_edge = 0;
    if (a.edge_distance > 0) {
        const int32_t cs = __ldg(a.contig_ptr + c), ce = __ldg(a.contig_ptr + c + 1);
        const int32_t A0 = a.ann_prefix[cs], nA = a.ann_prefix[ce] - A0;
        const int32_t r0 = k0 - A0, r1 = k1 - A0;  // ranks (within the contig) of the cluster's annotated genes
        const int32_t low1 = min(a.edge_distance, nA), high0 = max(0, nA - a.edge_distance);
        n_edge = overlap(r0, r1, 0, low1) + overlap(r0, r1, high0, nA) - overlap(r0, r1, high0, low1);
    }
    s.contig = c;
    s.begin = b;
    s.end = t;
    s.ordinal = r - a.contig_first_run[c] + 1;  // enumerate(clusters) per contig, :199-200
    s.valid = n_annot >= a.n_cds && (t - b) - n_edge >= a.n_cds;
    return s;
}

// The scan of round 3 runs over the run records; the number of runs is known on the device only, so both launches
// derive the same chunking from it.
template <bool kReplay>
__global__ void __launch_bounds__(kThreads)
emit_kernel(const SegmentsArgs a, int32_t *partial) {
    // runs are few (one per cluster candidate) and every evaluation is a chain of dependent loads: two runs per thread
    // spread them over as many CTAs as possible
    constexpr int kItems = 2, kTile = kThreads * kItems;
    __shared__ int32_t sWarp[kThreads / 32 + 1];
    const int64_t R = *a.n_runs;
    const int64_t tiles = (R + kTile - 1) / kTile;
    const int64_t chunk = ((tiles + gridDim.x - 1) / gridDim.x) * kTile;
    const int64_t c0 = min(R, (int64_t)blockIdx.x * chunk), c1 = min(R, c0 + chunk);
    int32_t carry = 0;
    if (kReplay) carry = block_fold(partial, (int)blockIdx.x, 0, sWarp);
    for (int64_t t0 = c0; t0 < c1; t0 += kTile) {
        const int64_t r0 = t0 + (int64_t)threadIdx.x * kItems;
        Segment seg[kItems];
        int32_t mine = 0;
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            seg[i].valid = false;
            if (r0 + i < c1) seg[i] = evaluate_run(a, (int32_t)(r0 + i));
            mine += seg[i].valid ? 1 : 0;
        }
        if (!kReplay) {
            carry += mine;
        } else {
            int32_t total;
            int32_t slot = carry + block_exclusive_scan(mine, 0, sWarp, &total);
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
                if (seg[i].valid) {
                    if (slot < a.capacity) {
                        a.seg_contig[slot] = seg[i].contig;
                        a.seg_begin[slot] = seg[i].begin;
                        a.seg_end[slot] = seg[i].end;
                        a.seg_ordinal[slot] = seg[i].ordinal;
                    }
                    ++slot;
                }
            }
            carry += total;
        }
    }
    if (!kReplay) {
        int32_t total;
        block_exclusive_scan(carry, 0, sWarp, &total);
        if (threadIdx.x == 0) partial[blockIdx.x] = total;
    } else if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        *a.count = (int64_t)carry;
    }
}

// per-cluster mean / max of the gene probabilities (gecco/model.py:443-454: genes without one are left out);
// one warp per cluster, NaN when no gene of the cluster has a probability
__global__ void __launch_bounds__(kThreads) stats_kernel(const SegmentsArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t n = min(*a.count, a.capacity);
    for (int64_t s = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5; s < n; s += ((int64_t)gridDim.x * kThreads) >> 5) {
        const int32_t b = a.seg_begin[s], e = a.seg_end[s];
        double sum = 0.0, mx = -1.0;
        int32_t cnt = 0;
        for (int32_t g = b + lane; g < e; g += 32) {
            const double p = load_prob(a.prob, a.prob_f32, g);
            if (p == p) {
                sum += p;
                mx = fmax(mx, p);
                ++cnt;
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, d);
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
            cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        }
        if (lane == 0) {
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            a.seg_avg_p[s] = cnt ? sum / cnt : nan;
            a.seg_max_p[s] = cnt ? mx : nan;
        }
    }
}

}  // namespace

size_t segments_scratch_bytes(int64_t G, int num_sms) {
    // flag[G] cmark[G+1] (bytes, padded) + ann_prefix[G+1] ann_pos[G] run_end/run_start/run_contig[<= G]
    // contig_first_run[<= G+1] + the run count + chunk summaries
    const size_t bytes8 = (((size_t)G + 16) & ~(size_t)15) * 2 + 32;
    return bytes8 + 6 * ((size_t)G + 4) * 4 + 16 + (size_t)num_sms * 8 * sizeof(SumMax) + 64;
}

cudaError_t launch_segments(SegmentsArgs args, void *scratch, int num_sms, cudaStream_t stream, int64_t *launches) {
    const int64_t G = args.G;
    cudaError_t err = cudaMemsetAsync(args.count, 0, sizeof(int64_t), stream);
    if (err != cudaSuccess || G <= 0) return err;
    // carve the scratch block
    char *p = static_cast<char *>(scratch);
    const size_t bytes8 = ((size_t)G + 16) & ~(size_t)15;
    args.flag = reinterpret_cast<uint8_t *>(p); p += bytes8;
    args.cmark = reinterpret_cast<uint8_t *>(p); p += bytes8 + 32;
    args.ann_prefix = reinterpret_cast<int32_t *>(p); p += ((size_t)G + 4) * 4;
    args.ann_pos = reinterpret_cast<int32_t *>(p); p += ((size_t)G + 4) * 4;
    args.run_end = reinterpret_cast<int32_t *>(p); p += ((size_t)G + 4) * 4;
    args.run_start = reinterpret_cast<int32_t *>(p); p += ((size_t)G + 4) * 4;
    args.run_contig = reinterpret_cast<int32_t *>(p); p += ((size_t)G + 4) * 4;
    args.contig_first_run = reinterpret_cast<int32_t *>(p); p += ((size_t)G + 4) * 4;
    args.n_runs = reinterpret_cast<int32_t *>(p); p += 16;
    void *partial = p;

    const int64_t tiles = (G + kTile - 1) / kTile;
    int64_t grid = (int64_t)num_sms * 8;
    if (grid > tiles) grid = tiles;
    Geometry geo;
    geo.G = G;
    geo.chunk = ((tiles + grid - 1) / grid) * kTile;
    grid = (G + geo.chunk - 1) / geo.chunk;
    const int n = (int)grid;

    if ((err = cudaMemsetAsync(args.cmark, 0, (size_t)G + 1, stream)) != cudaSuccess) return err;
    mark_kernel<<<(int)((args.C + 1 + kThreads - 1) / kThreads), kThreads, 0, stream>>>(args.contig_ptr, args.C, args.cmark);
    state_kernel<false><<<n, kThreads, 0, stream>>>(args, geo, static_cast<uint32_t *>(partial));
    state_kernel<true><<<n, kThreads, 0, stream>>>(args, geo, static_cast<uint32_t *>(partial));
    runs_kernel<false><<<n, kThreads, 0, stream>>>(args, geo, static_cast<SumMax *>(partial));
    runs_kernel<true><<<n, kThreads, 0, stream>>>(args, geo, static_cast<SumMax *>(partial));
    emit_kernel<false><<<n, kThreads, 0, stream>>>(args, static_cast<int32_t *>(partial));
    emit_kernel<true><<<n, kThreads, 0, stream>>>(args, static_cast<int32_t *>(partial));
    stats_kernel<<<num_sms * 16, kThreads, 0, stream>>>(args);  // the cluster count lives on the device: a wide grid-stride
    if (launches) *launches += 8;
    return cudaGetLastError();
}

}  // namespace gcrf
