// gcrf_segments.cu — threshold + contiguous-segment extraction on the marginal kernel's output (next row N1,
// SURVEY.md §8(f)): the array form of gecco.refine.ClusterRefiner with criterion "gecco"
// (gecco/refine.py:51-64 GeneGrouper, :183-200 _iter_clusters, :167-180 _trim_cluster, :139-165 _validate_cluster).
//
// The reference walks the genes in order with one piece of state (GeneGrouper.in_cluster, which survives contig
// boundaries because the grouper object is created once, :190).  Here the walk is TWO kernels, each a single-pass scan
// over a grid whose CTAs are all resident at once: a CTA reduces its chunk to a summary and publishes it (release store
// of a status word), waits for the summaries of the CTAs in front of it (acquire loads — no grid-wide barrier, no
// chain of dependent tiles: everybody publishes at about the same time and then folds what it needs), and replays its
// chunk from registers with the exclusive prefix.  Batches larger than one grid-full go round by round; the last CTA
// of a round publishes the round's inclusive prefix for the next one.
//
//   1. scan_kernel, over the genes.  The walker's state after a range of genes depends on the state it came in with
//      only through the genes in front of the first one that HAS a probability (a gene without one — NaN — inherits),
//      so a range is summarised as a FUNCTION of the incoming in-cluster bit: for both values, the outgoing bit, the
//      number of runs that ended inside, and the latest place a run could have started (with the count of annotated
//      genes in front of it); plus the plain counts of annotated genes and contig starts.  Function composition is
//      associative, which is all a scan needs.  With its exclusive prefix a thread replays its genes and writes one
//      RECORD per run where the run ends (first gene, end, contig, annotated-gene ranks of both), the compaction array
//      of annotated positions, and per contig the annotated-gene rank of its start and the index of its first record.
//   2. emit_kernel, over the run records: trim + validation in O(1) per run from its record (refine.py:139-180), the
//      per-contig ordinal from the index of the contig's first record (clusters are numbered per contig BEFORE
//      validation, :199-200), an ordered compaction of the valid ones (sum scan), and their mean / max probability by
//      eight lanes per cluster (gecco/model.py:443-454).
//
// HBM traffic: 9 bytes per gene read once (probability + annotation mark; the genes stay in registers as bit masks
// between the two phases), 4 bytes written per annotated gene, the rest is per run / per contig.  Contig starts are
// found per chunk from contig_ptr (no per-gene mark array).
#include "gcrf_kernels.cuh"

#include <cstdio>
#include <cstdlib>

namespace gcrf {

namespace {

constexpr int kThreads = 256;
constexpr int kItems = 32;
constexpr int kTile = kThreads * kItems;  // genes per sub-tile: 32 consecutive genes per thread, as bit masks
constexpr int kSub = 2;                   // sub-tiles a CTA keeps in registers per round
constexpr int kWarps = kThreads / 32;
constexpr int kEmitPer = 4;                // runs per thread and round in emit_kernel
constexpr unsigned kFull = 0xffffffffu;
static_assert(kSub * kWarps <= 32, "one warp folds the (sub-tile, warp) totals of a CTA");

// ---- the walker's concrete state after a prefix of the genes ------------------------------------------------------
struct Prefix {
    int32_t state;    // in-cluster bit of the last gene (0 in front of gene 0: GeneGrouper starts "out", :56)
    int32_t ends;     // runs that ended so far
    int32_t brk1;     // 1 + the latest position a run can have started at (0 = none yet)
    int32_t brk_ann;  // annotated genes in front of that position
    int32_t ann;      // annotated genes so far
    int32_t cst;      // contig starts so far
};

// ---- a range of genes as a function of the incoming in-cluster bit s --------------------------------------------
struct Summary {
    int32_t state;          // bit s: outgoing bit for incoming s
    int32_t ends[2];        // runs ended inside the range
    int32_t brk1[2];        // 1 + latest run-start position inside the range (absolute gene index), 0 = none
    int32_t brk_ann[2];     // annotated genes of the range in front of it
    int32_t ann, cst;
};

__device__ __forceinline__ Summary identity() { return Summary{2, {0, 0}, {0, 0}, {0, 0}, 0, 0}; }  // state: s -> s

// a first, then b
__device__ __forceinline__ Summary compose(const Summary &a, const Summary &b) {
    Summary r;
    r.state = 0;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const bool m = ((a.state >> s) & 1) != 0;  // selects, not indices: the arrays stay in registers
        r.state |= ((b.state >> (m ? 1 : 0)) & 1) << s;
        r.ends[s] = a.ends[s] + (m ? b.ends[1] : b.ends[0]);
        const int32_t bb = m ? b.brk1[1] : b.brk1[0], ba = m ? b.brk_ann[1] : b.brk_ann[0];
        const bool later = bb != 0;
        r.brk1[s] = later ? bb : a.brk1[s];
        r.brk_ann[s] = later ? a.ann + ba : a.brk_ann[s];
    }
    r.ann = a.ann + b.ann;
    r.cst = a.cst + b.cst;
    return r;
}

__device__ __forceinline__ Prefix apply(const Prefix &p, const Summary &f) {
    const bool s = p.state != 0;
    Prefix r;
    r.state = (f.state >> (s ? 1 : 0)) & 1;
    r.ends = p.ends + (s ? f.ends[1] : f.ends[0]);
    const int32_t fb = s ? f.brk1[1] : f.brk1[0], fa = s ? f.brk_ann[1] : f.brk_ann[0];
    const bool later = fb != 0;
    r.brk1 = later ? fb : p.brk1;
    r.brk_ann = later ? p.ann + fa : p.brk_ann;
    r.ann = p.ann + f.ann;
    r.cst = p.cst + f.cst;
    return r;
}

__device__ __forceinline__ Summary shfl(const Summary &v, int src_lane_delta_down) {
    Summary r;
    r.state = __shfl_down_sync(kFull, v.state, src_lane_delta_down);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        r.ends[s] = __shfl_down_sync(kFull, v.ends[s], src_lane_delta_down);
        r.brk1[s] = __shfl_down_sync(kFull, v.brk1[s], src_lane_delta_down);
        r.brk_ann[s] = __shfl_down_sync(kFull, v.brk_ann[s], src_lane_delta_down);
    }
    r.ann = __shfl_down_sync(kFull, v.ann, src_lane_delta_down);
    r.cst = __shfl_down_sync(kFull, v.cst, src_lane_delta_down);
    return r;
}
__device__ __forceinline__ Summary shfl_up1(const Summary &v, int d) {
    Summary r;
    r.state = __shfl_up_sync(kFull, v.state, d);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        r.ends[s] = __shfl_up_sync(kFull, v.ends[s], d);
        r.brk1[s] = __shfl_up_sync(kFull, v.brk1[s], d);
        r.brk_ann[s] = __shfl_up_sync(kFull, v.brk_ann[s], d);
    }
    r.ann = __shfl_up_sync(kFull, v.ann, d);
    r.cst = __shfl_up_sync(kFull, v.cst, d);
    return r;
}

// one gene, concrete state.  Returns true when a run ended right in front of gene x (the caller emits its record
// from the state BEFORE this call).
struct Gene {
    bool defined, above, cm, ann;
};
__device__ __forceinline__ bool step(Prefix &p, int32_t x, const Gene &g, bool reset_per_contig) {
    // a gene without probability inherits; with one iter_clusters call per contig (a fresh grouper, _common.py:616-618)
    // the state it would inherit at a contig start is "out"
    const bool flag = g.defined ? g.above : ((g.cm && reset_per_contig) ? false : p.state != 0);
    const bool ended = p.state != 0 && (!flag || g.cm);
    p.ends += ended ? 1 : 0;
    if (!flag) {  // a run can start behind this gene
        p.brk1 = x + 2;
        p.brk_ann = p.ann + (g.ann ? 1 : 0);
    } else if (g.cm) {  // or with it, when it opens a contig
        p.brk1 = x + 1;
        p.brk_ann = p.ann;
    }
    p.ann += g.ann ? 1 : 0;
    p.cst += g.cm ? 1 : 0;
    p.state = flag ? 1 : 0;
    return ended;
}

// ---- descriptors: what a CTA publishes per round ---------------------------------------------------------------------
constexpr int kNotReady = 0, kReady = 1;
struct alignas(16) Descriptor {
    int32_t status;
    int32_t pad[3];
    Summary summary;
};
struct alignas(16) RoundPrefix {
    int32_t status;     // the round's inclusive prefix has been written
    int32_t published;  // CTAs of the round whose summary is out
    Prefix inclusive;
};

__device__ __forceinline__ int32_t load_acquire(const int32_t *p) {
    int32_t v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void store_release(int32_t *p, int32_t v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <typename T>
__device__ __forceinline__ T load_cg(const T *p) {  // payloads are read after an acquire of the status word: skip L1
    T v;
    const int *src = reinterpret_cast<const int *>(p);
    int *dst = reinterpret_cast<int *>(&v);
#pragma unroll
    for (size_t i = 0; i < sizeof(T) / 4; ++i) dst[i] = __ldcg(src + i);
    return v;
}

// Ordered reduction of one Summary per lane to lane 0 (lane 0 = earliest).
__device__ __forceinline__ Summary warp_fold(Summary f) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Summary later = shfl(f, d);  // from lane + d
        if (lane + d < 32) f = compose(f, later);
    }
    return f;
}

__device__ __forceinline__ double load_prob(const void *prob, int f32, int64_t g) {
    return f32 ? (double)__ldg(static_cast<const float *>(prob) + g) : __ldg(static_cast<const double *>(prob) + g);
}

struct ScanScratch {
    Descriptor *desc;       // [rounds][grid]
    RoundPrefix *round;     // [rounds]
    int32_t *run_start, *run_end, *run_contig, *run_a0, *run_a1;  // one record per run, in order
    int32_t *ann_pos;       // [G] positions of the annotated genes, in order
    int32_t *contig_ann;    // [C+1] annotated genes in front of every contig (and in total)
    int32_t *contig_first_run;  // [C] index of the contig's first run record
    int32_t *n_runs;
};

// The genes of one thread in one sub-tile as bit masks (bit i = gene g0 + i; genes past the end of the batch: 0).
struct Masks {
    uint32_t def, abv, cm, ann;  // has a probability / above the threshold / opens a contig / annotated
    uint32_t valid;              // the gene exists
};
__device__ __forceinline__ Gene gene_at(const Masks &m, int i) {
    return Gene{((m.def >> i) & 1u) != 0, ((m.abv >> i) & 1u) != 0, ((m.cm >> i) & 1u) != 0, ((m.ann >> i) & 1u) != 0};
}
__device__ __forceinline__ uint32_t below(int i) { return i >= 32 ? kFull : (1u << i) - 1u; }  // bits [0, i)

// Masks of a sub-tile, built by ballots: in iteration j the warp reads the 32 genes of ITS thread j with coalesced
// loads (one gene per lane), votes, and lane j keeps the result — every lane ends up with the masks of its own 32
// consecutive genes without a byte going through shared memory.
template <typename P>
__device__ __forceinline__ Masks load_genes_t(const P *__restrict__ prob, const SegmentsArgs &a, int64_t tile0, int64_t g_end,
                                              const uint8_t *sCmSub) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Masks m{0u, 0u, 0u, 0u, 0u};
    const int64_t w0 = tile0 + (int64_t)warp * 32 * kItems;  // first gene of the warp's 32 x 32 block
    const P thr = (P)a.threshold;
    // the comparison runs in the array's own type: for float probabilities `p > (float)threshold` equals the reference's
    // `p > threshold` on the widened value unless the threshold itself is not a float — compare widened then
    const bool exact = (double)thr == a.threshold;
#pragma unroll 16
    for (int j = 0; j < 32; ++j) {
        const int64_t g = w0 + (int64_t)j * kItems + lane;
        const bool ok = g < g_end;
        const P p = ok ? __ldg(prob + g) : (P)0;
        const bool an = ok && __ldg(a.annotated + g) != 0;
        const bool cm = sCmSub[(warp * 32 + j) * kItems + lane] != 0;
        const bool above = exact ? p > thr : (double)p > a.threshold;
        const uint32_t v_def = __ballot_sync(kFull, ok && p == p), v_abv = __ballot_sync(kFull, ok && above);
        const uint32_t v_cm = __ballot_sync(kFull, cm), v_ann = __ballot_sync(kFull, an), v_ok = __ballot_sync(kFull, ok);
        if (lane == j) m = Masks{v_def, v_abv, v_cm, v_ann, v_ok};
    }
    return m;
}
__device__ __forceinline__ Masks load_genes(const SegmentsArgs &a, int64_t tile0, int64_t g_end, const uint8_t *sCmSub) {
    return a.prob_f32 ? load_genes_t(static_cast<const float *>(a.prob), a, tile0, g_end, sCmSub)
                      : load_genes_t(static_cast<const double *>(a.prob), a, tile0, g_end, sCmSub);
}

// A thread whose 32 genes all exist and all DETERMINE the walker's state themselves (they have a probability, or open a
// contig when every contig starts afresh): then everything is bit-parallel — the in-cluster bits are known without
// walking, a run ends in front of gene i when bit i-1 is set and gene i is out or opens a contig, and so on.  Anything
// else (a gene without probability, the ragged last thread) takes the gene-by-gene walk.
__device__ __forceinline__ bool is_fast(const Masks &m, bool reset_per_contig) {
    return m.valid == kFull && (m.def | (reset_per_contig ? m.cm : 0u)) == kFull;
}

struct FastBits {
    uint32_t flags, brk;  // in-cluster bits; "a run can (re)start here": the gene is out, or opens a contig
};
__device__ __forceinline__ FastBits fast_bits(const Masks &m) {
    FastBits f;
    f.flags = m.abv & m.def;  // a gene without probability is only here as a fresh contig start: out
    f.brk = ~f.flags | m.cm;
    return f;
}
// latest run-start position among the breakers in `mask` (relative to the thread's first gene); -1 when there is none
__device__ __forceinline__ int latest_start(const FastBits &f, uint32_t mask) {
    if (mask == 0) return -1;
    const int hi = 31 - __clz(mask);
    return ((f.flags >> hi) & 1u) ? hi : hi + 1;  // an in-cluster contig start opens a run itself, an out gene the one behind it
}

// summary of a thread's genes for both incoming bits
__device__ __forceinline__ Summary thread_summary(const Masks &m, int64_t g0, bool reset_per_contig) {
    Summary f;
    if (is_fast(m, reset_per_contig)) {
        const FastBits x = fast_bits(m);
        const uint32_t prev0 = x.flags << 1;
        f.ends[0] = __popc(prev0 & x.brk);
        f.ends[1] = __popc((prev0 | 1u) & x.brk);
        const int pos = latest_start(x, x.brk);
        const int32_t b1 = pos < 0 ? 0 : (int32_t)(g0 + pos) + 1, ba = pos < 0 ? 0 : __popc(m.ann & below(pos));
        f.brk1[0] = f.brk1[1] = b1;
        f.brk_ann[0] = f.brk_ann[1] = ba;
        const int out = (int)(x.flags >> 31);
        f.state = out | (out << 1);
        f.ann = __popc(m.ann);
        f.cst = __popc(m.cm);
        return f;
    }
    Prefix q0{0, 0, 0, 0, 0, 0}, q1{1, 0, 0, 0, 0, 0};
#pragma unroll 1
    for (int i = 0; i < kItems; ++i) {
        if ((m.valid >> i) & 1u) {
            const Gene g = gene_at(m, i);
            step(q0, (int32_t)(g0 + i), g, reset_per_contig);
            step(q1, (int32_t)(g0 + i), g, reset_per_contig);
        }
    }
    f.state = q0.state | (q1.state << 1);
    f.ends[0] = q0.ends; f.ends[1] = q1.ends;
    f.brk1[0] = q0.brk1; f.brk1[1] = q1.brk1;
    f.brk_ann[0] = q0.brk_ann; f.brk_ann[1] = q1.brk_ann;
    f.ann = q0.ann;
    f.cst = q0.cst;
    return f;
}

__device__ __forceinline__ void write_run(const ScanScratch &w, int32_t r, int32_t start, int32_t end, int32_t contig, int32_t a0, int32_t a1) {
    w.run_start[r] = start;
    w.run_end[r] = end;
    w.run_contig[r] = contig;
    w.run_a0[r] = a0;
    w.run_a1[r] = a1;
}

// A thread's genes replayed from its exclusive prefix p: run records and per-contig ranks (the annotated positions are
// written by the whole warp, see scan_kernel).  Returns the prefix behind the thread's genes.
__device__ __forceinline__ Prefix replay(const ScanScratch &w, Prefix p, const Masks &m, int64_t g0, bool reset_per_contig) {
    if (is_fast(m, reset_per_contig)) {
        const FastBits x = fast_bits(m);
        const uint32_t ended = ((x.flags << 1) | (uint32_t)p.state) & x.brk;
        for (uint32_t rest = m.cm; rest; rest &= rest - 1) {  // every contig start: annotated rank, first run record
            const int i = __ffs(rest) - 1;
            const int32_t c = p.cst + __popc(m.cm & below(i));
            w.contig_ann[c] = p.ann + __popc(m.ann & below(i));
            w.contig_first_run[c] = p.ends + __popc(ended & below(i + 1));  // a run ending right here is the previous contig's
        }
        for (uint32_t rest = ended; rest; rest &= rest - 1) {  // every run that ends in front of gene i
            const int i = __ffs(rest) - 1;
            const int pos = latest_start(x, x.brk & below(i));
            const int32_t start = pos < 0 ? p.brk1 - 1 : (int32_t)(g0 + pos);
            const int32_t a0 = pos < 0 ? p.brk_ann : p.ann + __popc(m.ann & below(pos));
            write_run(w, p.ends + __popc(ended & below(i)), start, (int32_t)(g0 + i), p.cst + __popc(m.cm & below(i)) - 1, a0,
                      p.ann + __popc(m.ann & below(i)));
        }
        const int pos = latest_start(x, x.brk);
        Prefix q;
        q.state = (int32_t)(x.flags >> 31);
        q.ends = p.ends + __popc(ended);
        q.brk1 = pos < 0 ? p.brk1 : (int32_t)(g0 + pos) + 1;
        q.brk_ann = pos < 0 ? p.brk_ann : p.ann + __popc(m.ann & below(pos));
        q.ann = p.ann + __popc(m.ann);
        q.cst = p.cst + __popc(m.cm);
        return q;
    }
#pragma unroll 1
    for (int i = 0; i < kItems; ++i) {
        if ((m.valid >> i) & 1u) {
            const int64_t g = g0 + i;
            const Prefix prev = p;
            const Gene gi = gene_at(m, i);
            if (gi.cm) w.contig_ann[prev.cst] = prev.ann;
            const bool ended = step(p, (int32_t)g, gi, reset_per_contig);
            // a run that ended right here belongs to the contig before: this contig's records start behind it
            if (gi.cm) w.contig_first_run[prev.cst] = p.ends;
            if (ended) write_run(w, prev.ends, prev.brk1 - 1, (int32_t)g, prev.cst - 1, prev.brk_ann, prev.ann);
        }
    }
    return p;
}

__global__ void __launch_bounds__(kThreads)
scan_kernel(const SegmentsArgs a, const ScanScratch w, const int sub_per_cta, const int rounds) {
    __shared__ __align__(16) uint8_t sCm[kSub * kTile];
    __shared__ Summary sWarp[kSub][kWarps];  // ordered totals: sub-tile k, warp w
    __shared__ Summary sPre[kSub][kWarps];   // everything of this CTA in front of (sub-tile k, warp w)
    __shared__ Summary sTotal;
    __shared__ Summary sFold[kWarps];
    __shared__ Prefix sCtaPrefix;
    __shared__ int64_t sFirstContig;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t G = a.G;
    const bool reset = a.reset_per_contig != 0;
    const int grid = (int)gridDim.x, b = (int)blockIdx.x;
    const int64_t chunk = (int64_t)sub_per_cta * kTile;  // genes of one CTA per round
#ifdef GCRF_SEG_PROFILE
    unsigned long long T[8];
#define SEG_T(i) do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(T[i])); } while (0)
#else
#define SEG_T(i) do { } while (0)
#endif
    for (int r = 0; r < rounds; ++r) {
        const int64_t c0 = ((int64_t)r * grid + b) * chunk, c1 = min(G, c0 + chunk);
        SEG_T(0);
        __syncthreads();  // the previous round's shared state is no longer read
        for (int i = tid; i < kSub * kTile / 16; i += kThreads) reinterpret_cast<uint4 *>(sCm)[i] = make_uint4(0u, 0u, 0u, 0u);
        // ---- contig starts inside the chunk, from contig_ptr: the first contig whose start is >= c0 is found by one warp
        //      (32 probes per round: three dependent round trips for 10,000 contigs instead of fourteen)
        if (warp == 0 && c0 < G) {
            int64_t lo = 0, hi = a.C;  // the smallest c with contig_ptr[c] >= c0 lies in [lo, hi] (contig_ptr[C] = G >= c0)
            while (lo < hi) {
                const int64_t st = (hi - lo + 31) / 32;
                const int64_t probe = lo + (int64_t)lane * st;  // sorted probes: "start < c0" holds for a prefix of them
                const bool is_below = probe < hi && (int64_t)__ldg(a.contig_ptr + probe) < c0;
                const int cnt = __popc(__ballot_sync(kFull, is_below));
                if (cnt == 0) {
                    hi = lo;
                } else {
                    const int64_t next = lo + (int64_t)cnt * st;  // first probe that was not below (if it exists)
                    hi = next < hi ? next : hi;
                    lo = lo + (int64_t)(cnt - 1) * st + 1;
                }
            }
            if (lane == 0) sFirstContig = lo;
        }
        __syncthreads();
        if (c0 < G) {
            for (int64_t c = sFirstContig + tid; c < a.C; c += kThreads) {
                const int64_t s = (int64_t)__ldg(a.contig_ptr + c);
                if (s >= c1) break;
                sCm[s - c0] = 1;
            }
        }
        __syncthreads();
        SEG_T(1);
        // ---- phase 1: the chunk's genes -> bit masks in registers; per sub-tile the thread's exclusive prefix inside its
        //      warp (kept for the replay) and the warp's total
        Masks m[kSub];
        Summary exc[kSub];
#pragma unroll
        for (int k = 0; k < kSub; ++k) {
            m[k] = Masks{0u, 0u, 0u, 0u, 0u};
            exc[k] = identity();
            const int64_t tile0 = c0 + (int64_t)k * kTile;
            if (k < sub_per_cta && tile0 < c1) {  // CTA-uniform
                m[k] = load_genes(a, tile0, c1, sCm + k * kTile);
                Summary inc = thread_summary(m[k], tile0 + (int64_t)tid * kItems, reset);
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const Summary earlier = shfl_up1(inc, d);
                    if (lane >= d) inc = compose(earlier, inc);
                }
                exc[k] = shfl_up1(inc, 1);
                if (lane == 0) exc[k] = identity();
                if (lane == 31) sWarp[k][warp] = inc;
            } else if (lane == 31) {
                sWarp[k][warp] = identity();
            }
        }
        __syncthreads();
        SEG_T(2);
        // ---- publish the CTA's summary; when the whole round has published, fold the summaries of the CTAs in front
        if (warp == 0) {
            Summary inc = identity();
            if (lane < kSub * kWarps) inc = sWarp[lane / kWarps][lane % kWarps];  // in order
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const Summary earlier = shfl_up1(inc, d);
                if (lane >= d) inc = compose(earlier, inc);
            }
            Summary pre = shfl_up1(inc, 1);
            if (lane == 0) pre = identity();
            if (lane < kSub * kWarps) sPre[lane / kWarps][lane % kWarps] = pre;
            if (lane == 31) {
                sTotal = inc;
                Descriptor *d = w.desc + (int64_t)r * grid + b;
                d->summary = inc;
                __threadfence();
                atomicAdd(&w.round[r].published, 1);
                // ONE thread per CTA waits, politely, until the whole round has published (everybody gets there at
                // about the same time); thousands of threads spinning on their own descriptor would eat the L2
                // bandwidth the late CTAs still need for their loads
                while (load_acquire(&w.round[r].published) < grid) __nanosleep(64);
            }
        }
        __syncthreads();
        SEG_T(3);
        {
            // thread t folds descriptors [t q, (t+1) q) of this round, q = ceil(b / threads): contiguous, so order is kept
            const int q = (b + kThreads - 1) / kThreads;
            Summary f = identity();
            for (int j = tid * q; j < min(b, (tid + 1) * q); ++j) f = compose(f, load_cg(&w.desc[(int64_t)r * grid + j].summary));
            f = warp_fold(f);
            if (lane == 0) sFold[warp] = f;
        }
        __syncthreads();
        if (tid == 0) {
            Summary f = identity();
#pragma unroll
            for (int k = 0; k < kWarps; ++k) f = compose(f, sFold[k]);
            Prefix base{0, 0, 0, 0, 0, 0};
            if (r > 0) {
                const RoundPrefix *rp = w.round + (r - 1);
                while (load_acquire(&rp->status) == kNotReady) __nanosleep(64);
                base = load_cg(&rp->inclusive);
            }
            const Prefix mine = apply(base, f);
            sCtaPrefix = mine;
            if (b == grid - 1) {  // the last CTA closes the round
                RoundPrefix *out = w.round + r;
                out->inclusive = apply(mine, sTotal);
                store_release(&out->status, kReady);
            }
        }
        __syncthreads();
        SEG_T(4);
        // ---- phase 2: replay from the masks with the exclusive prefix
#pragma unroll
        for (int k = 0; k < kSub; ++k) {
            const int64_t g0 = c0 + (int64_t)k * kTile + (int64_t)tid * kItems;
            if (k < sub_per_cta && c0 + (int64_t)k * kTile < c1) {  // CTA-uniform
                const Prefix start = apply(apply(sCtaPrefix, sPre[k][warp]), exc[k]);
                // annotated positions -> compaction array, transposed: in iteration j the warp takes thread j's genes, one
                // per lane, so that neighbouring lanes write neighbouring slots (a thread writing its own 32 slots
                // would issue 32 scattered 4-byte stores per instruction)
#pragma unroll 8
                for (int j = 0; j < 32; ++j) {
                    const uint32_t ann_j = __shfl_sync(kFull, m[k].ann, j);
                    const int32_t base_j = __shfl_sync(kFull, start.ann, j);
                    if ((ann_j >> lane) & 1u)
                        w.ann_pos[base_j + __popc(ann_j & below(lane))] = (int32_t)(g0 - (int64_t)lane * kItems + (int64_t)j * kItems + lane);
                }
                if (g0 >= c1) continue;
                const Prefix p = replay(w, start, m[k], g0, reset);
                if (g0 <= G - 1 && G - 1 < g0 + kItems) {  // the walker runs off the end: a run still open ends here
                    int32_t runs = p.ends;
                    if (p.state) write_run(w, runs++, p.brk1 - 1, (int32_t)G, p.cst - 1, p.brk_ann, p.ann);
                    *w.n_runs = runs;
                    w.contig_ann[a.C] = p.ann;
                }
            }
        }
        SEG_T(5);
#ifdef GCRF_SEG_PROFILE
        if (tid == 0 && (b == 0 || b == grid / 2 || b == grid - 1))
            printf("cta %d: setup %llu load+scan %llu publish+wait %llu fold %llu replay %llu ns (start %llu)\n", b, T[1] - T[0], T[2] - T[1],
                   T[3] - T[2], T[4] - T[3], T[5] - T[4], T[0] % 1000000ull);
#endif
    }
}

// ---- kernel 2: trim + validate every run (refine.py:139-180), ordered compaction, statistics ----------------------
struct Segment {
    int32_t contig, begin, end, ordinal;
    bool valid;
};

__device__ __forceinline__ int32_t overlap(int32_t a0, int32_t a1, int32_t b0, int32_t b1) {
    return max(0, min(a1, b1) - max(a0, b0));
}

__device__ Segment evaluate_run(const SegmentsArgs &a, const ScanScratch &w, int32_t r) {
    Segment s;
    const int32_t c = w.run_contig[r];
    int32_t b = w.run_start[r], t = w.run_end[r];
    const int32_t k0 = w.run_a0[r], k1 = w.run_a1[r];  // annotated-gene ranks of the run: [k0, k1)
    if (a.trim) {  // _trim_cluster: genes without domains are dropped from both ends
        if (k1 > k0) {
            b = w.ann_pos[k0];
            t = w.ann_pos[k1 - 1] + 1;
        } else {
            t = b;  // nothing annotated: the cluster is emptied
        }
    }
    // _validate_cluster, criterion "gecco": annotated genes of the cluster >= n_cds, and genes of the cluster that
    // are not among the contig's first / last `edge_distance` ANNOTATED genes >= n_cds
    const int32_t n_annot = k1 - k0;
    int32_t n_edge = 0;
    if (a.edge_distance > 0) {
        const int32_t A0 = w.contig_ann[c], nA = w.contig_ann[c + 1] - A0;
        const int32_t r0 = k0 - A0, r1 = k1 - A0;  // ranks (within the contig) of the cluster's annotated genes
        const int32_t low1 = min(a.edge_distance, nA), high0 = max(0, nA - a.edge_distance);
        n_edge = overlap(r0, r1, 0, low1) + overlap(r0, r1, high0, nA) - overlap(r0, r1, high0, low1);
    }
    s.contig = c;
    s.begin = b;
    s.end = t;
    s.ordinal = r - w.contig_first_run[c] + 1;  // enumerate(clusters) per contig, before validation (:199-200)
    s.valid = n_annot >= a.n_cds && (t - b) - n_edge >= a.n_cds;
    return s;
}

struct alignas(8) CountDescriptor {
    int32_t status, count;
};

__global__ void __launch_bounds__(kThreads)
emit_kernel(const SegmentsArgs a, const ScanScratch w, CountDescriptor *desc, CountDescriptor *round_desc) {
    constexpr int kPer = kEmitPer, kRuns = kThreads * kPer;  // runs of one CTA per round; run = base + k * threads + tid
    __shared__ int32_t sWarp[kPer][kWarps];
    __shared__ int32_t sFold[kWarps];
    __shared__ int32_t sBase, sValid;
    __shared__ int32_t sList[kRuns][3];  // slot, begin, end of the CTA's valid clusters
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grid = (int)gridDim.x, b = (int)blockIdx.x;
    const int64_t R = *w.n_runs;
    // runs per thread and round: as few as keep the whole grid busy (few runs: many small CTAs finish sooner than a
    // handful of full ones), at most kPer
    const int per = (int)max((int64_t)1, min((int64_t)kPer, (R + (int64_t)grid * kThreads - 1) / ((int64_t)grid * kThreads)));
    const int cta_runs = per * kThreads;
    const int rounds = (int)((R + (int64_t)grid * cta_runs - 1) / ((int64_t)grid * cta_runs));
    if (R == 0 && b == 0 && tid == 0) *a.count = 0;
    for (int r = 0; r < rounds; ++r) {
        const int64_t c0 = ((int64_t)r * grid + b) * cta_runs;
        // CTAs of this round that have runs at all: only they publish, wait for each other and close the round
        const int64_t left = R - (int64_t)r * grid * cta_runs;
        const int active = (int)min((int64_t)grid, (left + cta_runs - 1) / cta_runs);
        if (b >= active) return;  // (then also in every later round)
        __syncthreads();
        if (tid == 0) sValid = 0;
        Segment seg[kPer];
        int32_t inc[kPer];
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int64_t run = c0 + (int64_t)k * kThreads + tid;
            seg[k].valid = false;
            if (k < per && run < R) seg[k] = evaluate_run(a, w, (int32_t)run);
            inc[k] = seg[k].valid ? 1 : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int32_t y = __shfl_up_sync(kFull, inc[k], d);
                if (lane >= d) inc[k] += y;
            }
            if (lane == 31) sWarp[k][warp] = inc[k];
        }
        __syncthreads();
        int32_t total = 0;
#pragma unroll
        for (int k = 0; k < kPer; ++k)
#pragma unroll
            for (int v = 0; v < kWarps; ++v) total += sWarp[k][v];
        if (tid == 0) {
            desc[(int64_t)r * grid + b].count = total;
            __threadfence();
            atomicAdd(&round_desc[r].status, 4);  // bits 2..: CTAs of the round whose count is out; bit 0: round total written
            while ((load_acquire(&round_desc[r].status) >> 2) < active) __nanosleep(64);
        }
        __syncthreads();
        {
            const int q = (b + kThreads - 1) / kThreads;
            int32_t f = 0;
            for (int j = tid * q; j < min(b, (tid + 1) * q); ++j) f += __ldcg(&desc[(int64_t)r * grid + j].count);
            f = __reduce_add_sync(kFull, f);
            if (lane == 0) sFold[warp] = f;
        }
        __syncthreads();
        if (tid == 0) {
            int32_t f = 0;
#pragma unroll
            for (int k = 0; k < kWarps; ++k) f += sFold[k];
            if (r > 0) {
                const CountDescriptor *rp = round_desc + (r - 1);
                while ((load_acquire(&rp->status) & 1) == 0) __nanosleep(64);
                f += __ldcg(&rp->count);
            }
            sBase = f;
            if (b == active - 1) {
                CountDescriptor *out = round_desc + r;
                out->count = f + total;
                __threadfence();
                atomicAdd(&out->status, 1);
                if (r == rounds - 1) *a.count = (int64_t)(f + total);
            }
        }
        __syncthreads();
        int32_t before = sBase;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            int32_t upto = before;
#pragma unroll
            for (int v = 0; v < kWarps; ++v) {
                if (v < warp) upto += sWarp[k][v];
                before += sWarp[k][v];
            }
            if (seg[k].valid) {
                const int32_t slot = upto + inc[k] - 1;
                if (slot < a.capacity) {
                    a.seg_contig[slot] = seg[k].contig;
                    a.seg_begin[slot] = seg[k].begin;
                    a.seg_end[slot] = seg[k].end;
                    a.seg_ordinal[slot] = seg[k].ordinal;
                    const int n = atomicAdd(&sValid, 1);
                    sList[n][0] = slot;
                    sList[n][1] = seg[k].begin;
                    sList[n][2] = seg[k].end;
                }
            }
        }
        __syncthreads();
        // per-cluster mean / max of the gene probabilities (gecco/model.py:443-454: genes without one are left out);
        // eight lanes per cluster (clusters are a handful to a few dozen genes), NaN when no gene has a probability
        for (int n = tid >> 3; n < sValid; n += kThreads >> 3) {
            const int sub = tid & 7;
            const int32_t bgn = sList[n][1], end = sList[n][2];
            double sum = 0.0, mx = -1.0;
            int32_t cnt = 0;
            for (int32_t g = bgn + sub; g < end; g += 8) {
                const double p = load_prob(a.prob, a.prob_f32, g);
                if (p == p) {
                    sum += p;
                    mx = fmax(mx, p);
                    ++cnt;
                }
            }
            const unsigned group = 0xffu << (lane & 24);
#pragma unroll
            for (int d = 4; d >= 1; d >>= 1) {
                sum += __shfl_xor_sync(group, sum, d);
                mx = fmax(mx, __shfl_xor_sync(group, mx, d));
                cnt += __shfl_xor_sync(group, cnt, d);
            }
            if (sub == 0) {
                const double nan = __longlong_as_double(0x7ff8000000000000ll);
                a.seg_avg_p[sList[n][0]] = cnt ? sum / cnt : nan;
                a.seg_max_p[sList[n][0]] = cnt ? mx : nan;
            }
        }
    }
}

inline size_t round16(size_t x) { return (x + 15) & ~(size_t)15; }

struct Layout {
    size_t n_runs, desc, round, count_desc, count_round, run, ann_pos, contig_ann, contig_first_run, total;
    int64_t max_runs;
    int grid, sub_per_cta, rounds, emit_grid, emit_rounds;
    Layout(int64_t G, int64_t C, int scan_ctas, int emit_ctas) {
        const int64_t tiles = (G + kTile - 1) / kTile;
        grid = (int)(tiles < scan_ctas ? (tiles > 0 ? tiles : 1) : scan_ctas);
        int64_t per = (tiles + grid - 1) / grid;
        sub_per_cta = (int)(per < kSub ? (per > 0 ? per : 1) : kSub);
        rounds = (int)((tiles + (int64_t)grid * sub_per_cta - 1) / ((int64_t)grid * sub_per_cta));
        if (rounds < 1) rounds = 1;
        max_runs = G / 2 + 2;  // two runs need a gene between them
        const int64_t run_tiles = (max_runs + kThreads * kEmitPer - 1) / (kThreads * kEmitPer);
        emit_grid = (int)(run_tiles < emit_ctas ? run_tiles : emit_ctas);
        emit_rounds = (int)((max_runs + (int64_t)emit_grid * kThreads * kEmitPer - 1) / ((int64_t)emit_grid * kThreads * kEmitPer));
        size_t o = 0;
        n_runs = o; o += 64;  // zeroed together with the descriptors (the statuses must read "not ready")
        desc = o; o += round16((size_t)rounds * grid * sizeof(Descriptor));
        round = o; o += round16((size_t)rounds * sizeof(RoundPrefix));
        count_desc = o; o += round16((size_t)emit_rounds * emit_grid * sizeof(CountDescriptor));
        count_round = o; o += round16((size_t)emit_rounds * sizeof(CountDescriptor));
        run = o; o += round16((size_t)max_runs * 4) * 5;
        ann_pos = o; o += round16((size_t)(G + 1) * 4);
        contig_ann = o; o += round16((size_t)(C + 2) * 4);
        contig_first_run = o; o += round16((size_t)(C + 2) * 4);
        total = o;
    }
};

// Resident CTAs of both kernels: every CTA of a launch has to be on the machine at once (they wait for each other).
void resident_ctas(int num_sms, int *scan_ctas, int *emit_ctas) {
    static thread_local int cached_dev = -1, per_scan = 0, per_emit = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_scan, scan_kernel, kThreads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_emit, emit_kernel, kThreads, 0);
        cached_dev = dev;
    }
    *scan_ctas = num_sms * (per_scan >= 2 ? 2 : 1);
    *emit_ctas = num_sms * (per_emit >= 4 ? 4 : per_emit >= 2 ? 2 : 1);
}

}  // namespace

size_t segments_scratch_bytes(int64_t G, int64_t C, int num_sms) {
    int s = 0, e = 0;
    resident_ctas(num_sms, &s, &e);
    return Layout(G, C, s, e).total;
}

cudaError_t launch_segments(SegmentsArgs args, void *scratch, int num_sms, cudaStream_t stream, int64_t *launches) {
    const int64_t G = args.G;
    cudaError_t err = cudaSuccess;
    if (G <= 0) return cudaMemsetAsync(args.count, 0, sizeof(int64_t), stream);  // otherwise emit_kernel writes the count
    int scan_ctas = 0, emit_ctas = 0;
    resident_ctas(num_sms, &scan_ctas, &emit_ctas);
    const Layout lay(G, args.C, scan_ctas, emit_ctas);
    char *base = static_cast<char *>(scratch);
    // the run count and the descriptors of both kernels: one memset (the statuses must read "not ready")
    if ((err = cudaMemsetAsync(base, 0, lay.run, stream)) != cudaSuccess) return err;
    ScanScratch w;
    w.n_runs = reinterpret_cast<int32_t *>(base + lay.n_runs);
    w.desc = reinterpret_cast<Descriptor *>(base + lay.desc);
    w.round = reinterpret_cast<RoundPrefix *>(base + lay.round);
    const size_t run_bytes = round16((size_t)lay.max_runs * 4);
    w.run_start = reinterpret_cast<int32_t *>(base + lay.run);
    w.run_end = reinterpret_cast<int32_t *>(base + lay.run + run_bytes);
    w.run_contig = reinterpret_cast<int32_t *>(base + lay.run + 2 * run_bytes);
    w.run_a0 = reinterpret_cast<int32_t *>(base + lay.run + 3 * run_bytes);
    w.run_a1 = reinterpret_cast<int32_t *>(base + lay.run + 4 * run_bytes);
    w.ann_pos = reinterpret_cast<int32_t *>(base + lay.ann_pos);
    w.contig_ann = reinterpret_cast<int32_t *>(base + lay.contig_ann);
    w.contig_first_run = reinterpret_cast<int32_t *>(base + lay.contig_first_run);
    // cooperative launches: the runtime guarantees (or refuses) that every CTA of the grid is resident at once
    int sub_per_cta = lay.sub_per_cta, rounds = lay.rounds;
    CountDescriptor *cdesc = reinterpret_cast<CountDescriptor *>(base + lay.count_desc);
    CountDescriptor *cround = reinterpret_cast<CountDescriptor *>(base + lay.count_round);
    static const bool plain = [] { const char *e = getenv("GCRF_SEG_COOPERATIVE"); return !(e && e[0] == '1'); }();
    if (plain) {
        scan_kernel<<<(unsigned)lay.grid, kThreads, 0, stream>>>(args, w, sub_per_cta, rounds);
        emit_kernel<<<(unsigned)lay.emit_grid, kThreads, 0, stream>>>(args, w, cdesc, cround);
    } else {
        void *scan_args[] = {&args, &w, &sub_per_cta, &rounds};
        err = cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(scan_kernel), dim3((unsigned)lay.grid), dim3(kThreads), scan_args, 0, stream);
        if (err != cudaSuccess) return err;
        void *emit_args[] = {&args, &w, &cdesc, &cround};
        err = cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(emit_kernel), dim3((unsigned)lay.emit_grid), dim3(kThreads), emit_args, 0, stream);
        if (err != cudaSuccess) return err;
    }
    if (launches) *launches += 2;
    return cudaGetLastError();
}

}  // namespace gcrf
