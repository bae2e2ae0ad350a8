// gcrf_kernels.cuh — device-side parameter blocks and launcher prototypes shared by the kernels
// (gcrf_windowed.cu, gcrf_chain.cu) and the C ABI (gcrf_abi.cu).  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

struct gcrf_wire;  // include/gecco_crf_b200.h

namespace gcrf {

// Two-label model folded for the device ("two-state simplifications"):
// with o = the other label and p = pos_label, only
//     delta_a = W[a][p] - W[a][o]                     (per attribute, table[a]; table[A] = 0)
//     m01 = exp(T[o][p]-T[o][o]), m10 = exp(T[p][o]-T[o][o]), m11 = exp(T[p][p]-T[o][o])
// matter for the marginal of p: messages are carried as odds ratios x[p]/x[o].
struct ModelDev {
    const float *table;  // [A+1] on device, table[A] = 0 is the slot of unknown attributes
    int32_t A;
    float m01, m10, m11;
    float clamp;  // |delta_gene| is clamped to this before exp() so that odds stay finite in FP32
    // fixed-point twin of `table` for the streaming kernel: table_fx[a] = round(delta_a * 2^fx_bits),
    // table_fx[A] = 0.  Row sums are formed in wrapping int32 arithmetic, which is exact as long as a row
    // has fewer than fx_nsafe ids (rows with more take a float path).
    const int32_t *table_fx;
    int32_t fx_bits, fx_nsafe;
};

struct CsrDev {
    const int32_t *contig_ptr;  // [C+1]
    const int32_t *gene_ptr32;  // [G+1] or nullptr
    const int64_t *gene_ptr64;  // [G+1] or nullptr
    const int32_t *attr_idx;    // [nnz]
    int64_t C, G, nnz;
    // Sub-batch launches (the overlapped host path): contig_ptr, gene_ptr and out point at the slice's first
    // contig / gene, but the VALUES of contig_ptr still count genes from the start of the whole batch:
    // gene_base is subtracted from them.  gene_ptr values keep indexing the whole attr_idx array.
    int64_t gene_base;
    // ... and slice_ids, when positive, the attribute ids of the slice itself (nnz stays the length of the whole
    // attr_idx array, the bound of what may be read): the density the streaming kernel sizes its tiles by
    int64_t slice_ids;
};

struct WindowedArgs {
    ModelDev model;
    CsrDev csr;
    void *out;        // double[G] or float[G]
    int32_t out_f32;  // 0: double, 1: float
    int32_t window, step, pad;
    unsigned long long *prof;  // optional [16] cycle accumulators per kernel phase (nullptr = off); tuning aid
    int32_t debug_skip;        // tuning aid (GCRF_DEBUG_SKIP): 1 = skip the walk, 2 = skip the DP, 4 = skip pool/output
    // gcrf_marginals_windowed_peers (streaming kernel only): besides `out` (may be nullptr) every result is stored to
    // peer_out[k][gene], k < n_peer_out — the output arrays of the OTHER GPUs of a contig-sharded batch, mapped into this
    // process over NVLink (pointers already offset to this shard's first gene).  peer_multicast: peer_out[0] is an NVLS
    // multicast address — one multimem.st, replicated to every GPU by the switch.
    static constexpr int kMaxPeers = 8;
    void *peer_out[kMaxPeers];
    int32_t n_peer_out, peer_multicast;
};

// Geometry of the fused windowed kernel, fixed on the host so that tests can query it.
struct WindowedPlan {
    int threads;          // CTA size
    int tile_out;         // genes written per tile
    int slots;            // streaming kernel: window slots per tile (256; 128 / 64 for dense batches)
    int chunk;            // attribute ids staged per gather round (elements)
    size_t smem_bytes;    // dynamic shared memory per CTA
    int64_t num_tiles;
    int grid;             // persistent CTAs
    int tiles_per_cta;
    int ctas_per_sm;
};

// Fills `plan` (queries occupancy on the current device); returns cudaSuccess or the failure.
cudaError_t plan_windowed(const WindowedArgs &args, int num_sms, WindowedPlan *plan);

// Enqueue the fused gather + windowed forward-backward + max-pool kernel.  *launches += kernels launched.
cudaError_t launch_windowed(const WindowedArgs &args, const WindowedPlan &plan, cudaStream_t stream,
                            int64_t *launches);
// Largest window the generic kernel's shared-memory layout holds for a model of A attributes.
int windowed_max_window(int A);

// GCRF_FLAG_F64 — the windowed marginals in the reference's own arithmetic (gcrf_exact.cu): CRFsuite's scaled
// forward-backward in f64, operation by operation (restated on the CPU by the test oracle).
struct ExactArgs {
    CsrDev csr;
    int32_t A;
    const double *state_w;   // [A][2] state weights as given to gcrf_model_create (label-major pairs)
    double M[4];             // exp(trans_w), from -> to
    int32_t pos_label;
    int32_t window, step, pad;
    void *out;               // double[G] or float[G]
    int32_t out_f32;
    double *unary;           // [2*G] scratch: exp of the two state scores of every gene
    double *pool;            // [G] scratch for the max-pool when out is float, else nullptr (pool straight into out)
};
size_t exact_work_bytes(int64_t G, int32_t window, int num_sms);  // global work area of runtime window sizes
cudaError_t launch_exact(const ExactArgs &args, double *work, int num_sms, cudaStream_t stream, int64_t *launches);

// Fast fused kernel for the compile-time window sizes (gcrf_stream.cu); same arguments and results.
bool stream_supported(const WindowedArgs &args);
cudaError_t plan_stream(const WindowedArgs &args, int num_sms, WindowedPlan *plan);
cudaError_t launch_stream(const WindowedArgs &args, const WindowedPlan &plan, cudaStream_t stream, int64_t *launches);

struct ChainArgs {
    ModelDev model;
    CsrDev csr;
    void *out;
    int32_t out_f32;
    const double *table64;  // [A+1] f64 twin of ModelDev::table (table64[A] = 0)
    double m01, m10, m11;   // f64 twins of the folded transition odds
    double *scratch;        // [2*G] device scratch: unary odds, then forward odds
};

// Whole-contig (un-windowed) marginals: unary gather kernel + block-per-contig 2x2 scan kernel.
cudaError_t launch_chain(const ChainArgs &args, int num_sms, cudaStream_t stream, int64_t *launches);

// accession -> attribute id with the reference's set semantics (repeats inside a gene and unknown accessions -> -1)
cudaError_t launch_features(const int32_t *accession, const int32_t *gene_ptr32, const int64_t *gene_ptr64, int64_t G,
                            int64_t nnz, const int32_t *lut, int32_t lut_size, int32_t num_attrs, int32_t *attr_idx_out,
                            int num_sms, cudaStream_t stream, int64_t *launches);

// uint16 attribute ids -> int32 (0xFFFF -> -1); out holds at least round_up(n, 8) entries
cudaError_t launch_widen_u16(const uint16_t *in, int32_t *out, int64_t n, int num_sms, cudaStream_t stream, int64_t *launches);

// Compact wire format (gcrf_wire.cu): rebuild gene_ptr[G+1] / attr_idx[nnz] of one slice from its length arrays and its
// stretch of the delta-coded id stream.  `sums` = the encoder's chunk sums of the slice (device copy), id_base = attribute
// ids in front of the slice; gene_ptr (the slice's first entry) receives absolute offsets into attr_idx; rice_k = the
// batch's code parameter.  One launch.
int64_t wire_chunks(int64_t G);
cudaError_t launch_wire_decode(const void *len_ids, const void *len_bytes, int32_t len_width, const uint8_t *stream, int64_t G,
                               int64_t id_base, int32_t rice_k, const int64_t *sums, int32_t *gene_ptr, int32_t *attr_idx,
                               cudaStream_t stream_, int64_t *launches);
int32_t wire_rice_k(const gcrf_wire *w);
int wire_slices(const gcrf_wire *w);
void wire_slice(const gcrf_wire *w, int k, int64_t *contig, int64_t *gene, int64_t *id, int64_t *byte);  // k in [0, slices]
const char *wire_block(const gcrf_wire *w);
size_t wire_total(const gcrf_wire *w);
size_t wire_head_bytes(const gcrf_wire *w);
size_t wire_off_sums(const gcrf_wire *w);
void wire_section(const gcrf_wire *w, int k, size_t *off, size_t *size, size_t *rel_len_bytes, size_t *rel_stream, int64_t *first_chunk);
int32_t wire_len_width(const gcrf_wire *w);

// Threshold + segment extraction (gcrf_segments.cu; gecco/refine.py:51-200, criterion "gecco").
struct SegmentsArgs {
    const int32_t *contig_ptr;  // [C+1]
    const void *prob;           // double[G] or float[G]; NaN = the gene has no probability
    int32_t prob_f32;
    const uint8_t *annotated;   // [G] gene has >= 1 domain (gene.protein.domains is non-empty)
    int64_t C, G;
    double threshold;
    int32_t n_cds, edge_distance, trim;
    int32_t reset_per_contig;   // in-cluster state starts at False in every contig (one iter_clusters call per contig)
    // outputs, `capacity` entries each, in the reference's order
    int32_t *seg_contig, *seg_begin, *seg_end, *seg_ordinal;
    double *seg_avg_p, *seg_max_p;
    int64_t capacity;
    int64_t *count;             // device: number of valid clusters found (may exceed capacity)
};
size_t segments_scratch_bytes(int64_t G, int64_t C, int num_sms);
cudaError_t launch_segments(SegmentsArgs args, void *scratch, int num_sms, cudaStream_t stream, int64_t *launches);

}  // namespace gcrf
