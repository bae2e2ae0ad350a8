/*
 * gecco_crf_b200.h — C ABI of libgecco_crf_b200.so: linear-chain CRF marginal inference for
 * GECCO's ClusterCRF on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for ONE path of the reference (zellerlab/GECCO v0.11.0): the body
 * of gecco.crf.ClusterCRF.predict_probabilities (gecco/crf/__init__.py:148-273).  The reference
 * crosses into native code at exactly one place on that path,
 *     self.model.predict_marginals_single(feats[win])          gecco/crf/__init__.py:253
 * (sklearn-crfsuite -> python-crfsuite Tagger.set()/Tagger.marginal() -> CRFsuite C), once per
 * W-gene window.  The functions below replace that binding in bulk: the caller packs all contigs
 * into CSR arrays once and gets every gene's max-pooled marginal back.
 *
 * Conventions
 *   - plain C types only; the caller owns every buffer, the library owns only the handle;
 *   - every function returns GCRF_OK (0) or a negative gcrf_status; gcrf_last_error() returns a
 *     thread-local message for the last failure on the calling thread;
 *   - a handle is bound to one CUDA device and is not thread-safe; distinct handles are;
 *   - there is NO CPU fallback: without a usable CUDA device gcrf_model_create fails with
 *     GCRF_ENODEVICE.
 *
 * Data layout (all row pointers are int32 unless GCRF_FLAG_PTR64 is given)
 *   contig_ptr[C+1]  genes of contig c are [contig_ptr[c], contig_ptr[c+1]); strictly increasing
 *                    (a contig has >= 1 gene: itertools.groupby never yields an empty group,
 *                    gecco/crf/__init__.py:204-206); contig_ptr[0] = 0, contig_ptr[C] = G
 *   gene_ptr[G+1]    attribute ids of gene g are attr_idx[gene_ptr[g] .. gene_ptr[g+1]);
 *                    a gene without domains is an empty row (the `{}` item of
 *                    gecco/crf/features.py:31-35)
 *   attr_idx[nnz]    attribute ids in [0, A); any id outside that range (use -1) is an attribute
 *                    the model does not know and contributes nothing (python-crfsuite drops
 *                    unknown attribute strings).  Ids must be unique within a gene: the reference
 *                    builds a dict keyed by domain name (features.py:32), so duplicates collapse
 *                    at feature extraction — the packer does that.
 *   genes are ordered as the reference orders them: by contig, then by start (:199).
 */
#ifndef GECCO_CRF_B200_H
#define GECCO_CRF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCRF_ABI_VERSION 1

typedef enum gcrf_status {
    GCRF_OK = 0,
    GCRF_EINVAL = -1,       /* bad argument (also: window/step rule of gecco/_meta.py:127-130) */
    GCRF_ENODEVICE = -2,    /* no usable CUDA device / wrong architecture */
    GCRF_ECUDA = -3,        /* a CUDA runtime call failed; see gcrf_last_error() */
    GCRF_ENOMEM = -4,       /* host or device allocation failed */
    GCRF_EUNSUPPORTED = -5  /* model shape not supported by the device path (e.g. L != 2) */
} gcrf_status;

/* flags of the marginal calls */
#define GCRF_FLAG_DEVICE_PTRS 0x1u /* contig_ptr/gene_ptr/attr_idx/out are device pointers on the
                                      handle's device; the call only enqueues work on the handle's
                                      stream (gcrf_model_set_stream) and does not synchronise.
                                      The arrays are trusted (no validation pass): contig_ptr strictly
                                      increasing from 0 to G, gene_ptr non-decreasing from 0 to nnz.
                                      attr_idx must be 16-byte aligned and READABLE up to the next
                                      multiple of 4 elements past nnz (the streaming kernel moves ids
                                      with 16-byte bulk copies; the over-read values are ignored) */
#define GCRF_FLAG_OUT_F32     0x2u /* out is float[G] instead of double[G] */
#define GCRF_FLAG_PTR64       0x4u /* gene_ptr is int64_t[G+1] (nnz >= 2^31); contig_ptr stays int32 */
#define GCRF_FLAG_IDX_U16     0x20u /* attr_idx is uint16_t[nnz] (0xFFFF = unknown attribute; models with fewer than
                                      65535 attributes): half the bytes over PCIe, widened on the device */
#define GCRF_FLAG_ACCESSIONS  0x40u /* gcrf_marginals_*: attr_idx holds integer domain accessions (one row per domain in
                                      domain-start order, e.g. 394 for PF00394), not attribute ids: they are mapped
                                      through the vocabulary (gcrf_model_set_vocabulary) and de-duplicated per gene on
                                      the device first — gcrf_features_from_accessions and the marginals in one call
                                      (gecco/crf/features.py:13-35 + crf/__init__.py:253).  Not with GCRF_FLAG_IDX_U16 */
#define GCRF_FLAG_F64         0x80u /* gcrf_marginals_windowed: the reference's own arithmetic — CRFsuite's scaled
                                      forward-backward in f64 (floatval_t is double), operation by operation and without
                                      fused multiply-adds, exp() correctly rounded: bit-identical to python-crfsuite's
                                      output on the reference's golden fixture, within a few ulps of it elsewhere (the
                                      residue is the host libm's rounding of exp).  Any window size.  Several times
                                      slower than the default FP32 odds-ratio kernels, whose results stay within 1e-5 */
#define GCRF_FLAG_MULTICAST   0x100u /* gcrf_marginals_windowed_peers: peer_out[0] is an NVLS multicast address */
#define GCRF_FLAG_PROB_F32    0x8u /* gcrf_segments: prob is float[G] instead of double[G] */
#define GCRF_FLAG_RESET_PER_CONTIG 0x10u /* gcrf_segments: the in-cluster state starts at "out" in every
                                      contig, i.e. one ClusterRefiner.iter_clusters call per contig as
                                      `gecco run` does (gecco/cli/commands/_common.py:616-618); default: one
                                      call over all contigs, where the state of GeneGrouper leaks across
                                      contig boundaries (gecco/refine.py:190) */

typedef struct gcrf_model gcrf_model;

/* Library / ABI version (GCRF_ABI_VERSION of the build). */
int gcrf_version(void);

/* Message of the last error raised on the calling thread ("" if none). Never NULL. */
const char *gcrf_last_error(void);

/* Number of CUDA devices the library can see (0 if none / driver missing). */
int gcrf_device_count(void);

/*
 * Create a model handle on CUDA device `device`.
 *
 * Replaces: unpickling the tagger and pycrfsuite.Tagger.open() behind
 * ClusterCRF.trained() (gecco/crf/__init__.py:61-99).
 *
 *   state_w   [A][L] row-major state-feature weights, 0 where the model has no feature
 *   trans_w   [L][L] transition weights, from -> to
 *   pos_label id of the label whose marginal is reported (the id of '1', looked up by NAME by the
 *             caller — gecco/crf/__init__.py:253 asks for p['1'])
 * The device path supports L == 2 (GECCO's labels are '0'/'1'); other L -> GCRF_EUNSUPPORTED.
 */
int gcrf_model_create(const double *state_w, int32_t A, int32_t L, const double *trans_w,
                      int32_t pos_label, int32_t device, gcrf_model **out);

void gcrf_model_destroy(gcrf_model *model);

/* Use `cuda_stream` (a cudaStream_t; NULL = the handle's own stream) for all later work. */
int gcrf_model_set_stream(gcrf_model *model, void *cuda_stream);

/* Block until everything enqueued on the handle's stream has finished. */
int gcrf_model_synchronize(gcrf_model *model);

/*
 * Per-gene cluster probability for a batch of contigs — the whole hot loop of
 * ClusterCRF.predict_probabilities (gecco/crf/__init__.py:209-258) in one call:
 *
 *   for every contig with n genes
 *     n <  window, pad != 0 : one window over delta/2 empty items + the genes + (delta+1)/2 empty
 *                             items, delta = window - n              (:216-227, read-back :258)
 *     n <  window, pad == 0 : the contig is skipped; its genes get NaN  (:228-234, :246-248 leave
 *                             the genes without a probability)
 *     n >= window           : windows [i, i+window) for i = 0, step, 2*step, ... while
 *                             i + window <= n        (gecco/_meta.py:124-132)
 *     out[g] = max over the windows covering g of P(y_g = pos_label | window), 0.0 if no window
 *              covers g (possible only when step > 1)                 (:251-254)
 *
 * where P(.) is the first-order CRF marginal of CRFsuite (Tagger.marginal): unary scores are the
 * sums of the state weights of the gene's attributes, see SURVEY.md Appendix B.
 *
 * Host-pointer mode (default): the call copies the inputs to the device, runs, copies `out` back
 * and returns when `out` is complete.  Device-pointer mode: see GCRF_FLAG_DEVICE_PTRS.
 * Result tolerance vs the f64 reference arithmetic: |dp| <= 1e-5 (FP32 device arithmetic;
 * measured <= 2e-6, tests/test_gpu_parity.py); with GCRF_FLAG_F64 <= 1e-12.
 * Window sizes (the reference takes any window_size >= 1, gecco/crf/__init__.py:134-137): the FP32 kernels hold a
 * window's state in shared memory — 5, 10 and 20 run the streaming kernel, other sizes up to gcrf_max_window(model, 0)
 * (128, fewer for very large vocabularies) the generic one, larger ones fail with GCRF_EUNSUPPORTED; with
 * GCRF_FLAG_F64 any window size works.
 */
int gcrf_marginals_windowed(gcrf_model *model, const int32_t *contig_ptr, const void *gene_ptr,
                            const void *attr_idx, int64_t C, int64_t G, int64_t nnz,
                            int32_t window, int32_t step, int32_t pad, void *out, uint32_t flags);

/*
 * The multi-GPU form of gcrf_marginals_windowed for ONE batch sharded by contig over the GPUs of a node (contigs are
 * independent: every GPU runs the kernels on its own contigs, nothing is exchanged inside the math).  What the GPUs do
 * exchange is the result — every gene's marginal has to end up in one array — and this entry point fuses that gather into
 * the kernel: besides `out` (this GPU's own array, may be NULL) every result is stored straight into peer_out[k]
 * [out_offset + g], k < n_peer_out <= 8: the output arrays of the other GPUs (and, if wanted, of this one) mapped into
 * this process over NVLink / NVSwitch (cudaIpc*, cuMem* fabric handles, or torch symmetric memory —
 * gecco_b200/sharding.py).  With GCRF_FLAG_MULTICAST peer_out[0] is an NVLS multicast address: one store, replicated to
 * every GPU by the switch.  The transfer overlaps the computation tile by tile; no collective follows, only a barrier
 * among the GPUs (theirs to provide) before anybody reads.  out_offset = index of this shard's first gene in the
 * gathered array.  Device pointers only (GCRF_FLAG_DEVICE_PTRS is required); the streaming kernel only (window 5 or 20,
 * FP32 arithmetic), GCRF_EUNSUPPORTED otherwise — fall back to gcrf_marginals_windowed + a collective.
 * Replaces: nothing in the reference, which is single-process (gecco/crf/__init__.py:244-258).
 */
int gcrf_marginals_windowed_peers(gcrf_model *model, const int32_t *contig_ptr, const void *gene_ptr,
                                  const void *attr_idx, int64_t C, int64_t G, int64_t nnz, int32_t window,
                                  int32_t step, int32_t pad, void *out, void *const *peer_out, int32_t n_peer_out,
                                  int64_t out_offset, uint32_t flags);

/*
 * Primitive equal to predict_marginals_single() on whole rows: every contig is ONE chain of
 * n items (no windows, no padding); out[g] = P(y_g = pos_label | whole contig).
 * Replaces: sklearn_crfsuite.CRF.predict_marginals_single (call site gecco/crf/__init__.py:253)
 * for callers that want the un-windowed marginal, and is the deep-chain path (BASELINE config 5).
 */
int gcrf_marginals_chain(gcrf_model *model, const int32_t *contig_ptr, const void *gene_ptr,
                         const void *attr_idx, int64_t C, int64_t G, int64_t nnz, void *out,
                         uint32_t flags);

/*
 * On-device feature extraction (gecco/crf/features.py:13-35 + the attribute dictionary of the tagger).
 *
 * gcrf_model_set_vocabulary: accession_of_attr[a] is the integer accession of attribute a (for Pfam, the
 * number in "PF00109" -> 109); accessions must be >= 0 and distinct.  Replaces: the CQDB string -> id
 * dictionary python-crfsuite consults for every attribute of every item (Tagger.set()).
 *
 * gcrf_features_from_accessions: accession[nnz] holds, row by row (gene_ptr), the accessions of the
 * domain rows of every gene in domain-start order; attr_idx_out[p] receives the attribute id, or -1 when
 * the accession is not in the model or repeats an earlier row of the same gene — a gene's features are
 * a dict keyed by domain name (features.py:32), so repeats collapse.  gene_ptr is unchanged: the -1
 * entries are simply ignored by gcrf_marginals_*.  Honours GCRF_FLAG_DEVICE_PTRS and GCRF_FLAG_PTR64.
 */
int gcrf_model_set_vocabulary(gcrf_model *model, const int32_t *accession_of_attr, int32_t A);
int gcrf_features_from_accessions(gcrf_model *model, const int32_t *accession, const void *gene_ptr,
                                  int64_t G, int64_t nnz, int32_t *attr_idx_out, uint32_t flags);

/*
 * Threshold + contiguous-segment extraction on the per-gene probabilities — the array form of
 * gecco.refine.ClusterRefiner(criterion="gecco").iter_clusters (gecco/refine.py:120-200), the immediate
 * consumer of gcrf_marginals_windowed's output in `gecco run` (gecco/cli/commands/_common.py:594-618).
 *
 *   prob[G]       per-gene probability in the same gene order as gcrf_marginals_*; NaN = the gene has no
 *                 probability (Gene.average_probability is None): it inherits the in-cluster state of the
 *                 previous gene — also across a contig boundary, like the reference's single GeneGrouper
 *                 (refine.py:51-64, :190)
 *   annotated[G]  1 if the gene has at least one domain (gene.protein.domains is non-empty)
 *   threshold     a gene is in a cluster when prob > threshold                          (refine.py:62)
 *   trim          drop un-annotated genes from both ends of every run                   (refine.py:167-180)
 *   n_cds, edge_distance   validation, criterion "gecco"                                (refine.py:139-156)
 *
 * Outputs, in the reference's order (contigs as given, runs left to right), `capacity` entries each:
 *   seg_contig / seg_begin / seg_end   cluster = genes [seg_begin, seg_end) (global gene indices) of that contig
 *   seg_ordinal   1-based index of the RAW run inside its contig: the reference names clusters
 *                 "{contig}_cluster_{ordinal}" before validation filters them           (refine.py:199-200)
 *   seg_avg_p / seg_max_p   mean / max of the probabilities of the cluster's genes that have one
 *                 (Cluster.average_probability / maximum_probability, gecco/model.py:443-454), else NaN
 * *n_segments (always a HOST pointer) receives the number of valid clusters; if it exceeds `capacity` only
 * the first `capacity` were written — call again with larger arrays.  The call synchronises the handle's
 * stream in both pointer modes (it has to read the count).  Honours GCRF_FLAG_DEVICE_PTRS, GCRF_FLAG_PROB_F32,
 * GCRF_FLAG_RESET_PER_CONTIG.
 */
int gcrf_segments(gcrf_model *model, const int32_t *contig_ptr, const void *prob, const uint8_t *annotated,
                  int64_t C, int64_t G, double threshold, int32_t n_cds, int32_t edge_distance, int32_t trim,
                  int32_t *seg_contig, int32_t *seg_begin, int32_t *seg_end, int32_t *seg_ordinal,
                  double *seg_avg_p, double *seg_max_p, int64_t capacity, int64_t *n_segments, uint32_t flags);

/*
 * A compact wire format for host-buffer batches.  A host-pointer call is one PCIe copy (the kernels are ~2 % of it), so
 * the bytes that cross are its cost: 4 per attribute id, 4 per gene (row pointer), 8 per gene (float64 marginal) in the
 * CSR layout above.  gcrf_wire_encode packs the same batch into ONE page-locked block — contig_ptr unchanged, per gene the
 * number of ids and of stream bytes (uint8 each, uint16 when a gene has more than 255), and per gene its ids SORTED
 * ascending, unknown ids mapped to A, as Rice-coded deltas (parameter chosen per batch): ~1.06 bytes per id for 25 ids
 * out of 2,659 — and gcrf_marginals_windowed_wire moves that block in contig-aligned slices (copy in, decode + kernels
 * and copy back of successive slices overlap on three streams), rebuilds gene_ptr / attr_idx on the device (one kernel
 * per slice) and runs the regular kernels.  Results equal gcrf_marginals_windowed's on the unsorted batch bit for bit in the default
 * FP32 arithmetic (row sums are exact integer sums; only rows long enough to take the float path, >= ~100 ids for the
 * shipped model, can differ in the last bit).  With GCRF_FLAG_F64 the row sums run in the SORTED order, so the result is
 * within a few ulps of the first-occurrence order of the reference, not bit-identical to it.  Combine with
 * GCRF_FLAG_OUT_F32 to halve the bytes coming back (the FP32 results, un-widened).
 * Replaces nothing in the reference (which never leaves the host); it is the packers' output format for bulk calls.
 * The encoder is host code (no device needed); nnz must stay below 2^31, a gene below 65,536 ids and stream bytes,
 * the vocabulary below 2^24 attributes.
 */
typedef struct gcrf_wire gcrf_wire;
const char *gcrf_wire_last_error(void);
int gcrf_wire_encode(const int32_t *contig_ptr, const void *gene_ptr, const int32_t *attr_idx, int64_t C, int64_t G,
                     int64_t nnz, int32_t num_attrs, uint32_t flags /* GCRF_FLAG_PTR64 */, gcrf_wire **out);
void gcrf_wire_destroy(gcrf_wire *wire);
int64_t gcrf_wire_bytes(const gcrf_wire *wire);   /* size of the block = host-to-device bytes of a call */
int64_t gcrf_wire_contigs(const gcrf_wire *wire);
int64_t gcrf_wire_genes(const gcrf_wire *wire);
int64_t gcrf_wire_ids(const gcrf_wire *wire);
/* decode on the host (tests, debugging): gene_ptr[G+1], attr_idx[nnz] = the sorted batch the device will see */
int gcrf_wire_decode_host(const gcrf_wire *wire, int32_t *gene_ptr, int32_t *attr_idx);
int gcrf_marginals_windowed_wire(gcrf_model *model, const gcrf_wire *wire, int32_t window, int32_t step, int32_t pad,
                                 void *out, uint32_t flags /* GCRF_FLAG_OUT_F32, GCRF_FLAG_F64 */);

/*
 * Pinned host memory helpers so that host-pointer calls can run their copies at PCIe speed
 * (pageable buffers are accepted everywhere, they are just slower).
 */
int gcrf_host_alloc(void **ptr, uint64_t bytes);
int gcrf_host_free(void *ptr);

/* Largest `window` gcrf_marginals_windowed accepts for this model (f64 != 0: with GCRF_FLAG_F64 — INT32_MAX). */
int32_t gcrf_max_window(const gcrf_model *model, int32_t f64);

/* Number of kernel launches the handle has issued so far (for bench.py's gpu_launches). */
int64_t gcrf_model_launch_count(const gcrf_model *model);

/*
 * Kernel timing.  With timing enabled (off by default: the two event records sit between back-to-back
 * launches) every marginal call brackets its compute kernels with CUDA events on the handle's stream;
 * gcrf_model_last_kernel_ms returns the device time in milliseconds of the LAST such call (kernels only,
 * no copies), synchronising the stream, or < 0 when timing was off or on error.
 */
int gcrf_model_set_timing(gcrf_model *model, int32_t enable);
double gcrf_model_last_kernel_ms(gcrf_model *model);

/*
 * Tables (host code; no device needed) — the input and output side of `gecco predict` without per-row
 * Python objects.  Replaces, in the reference's own order of operations (gecco/cli/commands/predict.py:62-100):
 *   GeneTable.load / FeatureTable.load           gecco/_base.py:119-131, gecco/model.py:621-637, 773-789
 *   annotate_genes                               gecco/cli/commands/_common.py:211-262
 *   the coordinate sorts                         gecco/cli/commands/predict.py:81-83
 *   filter_domains                               gecco/cli/commands/_common.py:419-448
 *   extract_features_protein / _domain           gecco/crf/features.py:13-48
 *   GeneTable.from_genes(...).dump               gecco/model.py:791-813, gecco/_base.py:133-151
 *   FeatureTable.from_genes(...).dump            gecco/model.py:644-670
 * The reference reads the tables with polars (Rust) and rebuilds Gene/Protein/Domain objects row by row.
 *
 * gcrf_table_load reads one genes table and n_features feature tables (tab-separated, header line, "\n" or
 * "\r\n"; decompress first if they are gzipped — gcrf_table_parse takes memory buffers).  Genes end up ordered by
 * (sequence_id, start, end), every gene's domain rows by (domain_start, domain_end), both stable; rows with
 * i_evalue >= e_filter or pvalue >= p_filter are dropped (pass NaN for "no filter"; `gecco predict` defaults to
 * p_filter = 1e-9 and no e_filter).  Errors mirror the reference's ValueErrors (duplicate gene names, a feature row
 * that disagrees with its gene) and are reported through gcrf_table_last_error().
 */
typedef struct gcrf_table gcrf_table;

const char *gcrf_table_last_error(void);
int gcrf_table_load(const char *genes_tsv, const char *const *features_tsv, int32_t n_features, double e_filter,
                    double p_filter, gcrf_table **out);
int gcrf_table_parse(const char *genes, uint64_t genes_len, const char *const *features, const uint64_t *features_len,
                     int32_t n_features, double e_filter, double p_filter, gcrf_table **out);
void gcrf_table_destroy(gcrf_table *table);

int64_t gcrf_table_contigs(const gcrf_table *table); /* C */
int64_t gcrf_table_genes(const gcrf_table *table);   /* G */
int64_t gcrf_table_domains(const gcrf_table *table); /* domain rows left after the filters */
const char *gcrf_table_contig_id(const gcrf_table *table, int64_t contig);
const char *gcrf_table_gene_id(gcrf_table *table, int64_t gene);
const int32_t *gcrf_table_contig_ptr(const gcrf_table *table); /* [C+1] into genes */
const uint8_t *gcrf_table_annotated(const gcrf_table *table);  /* [G] gene kept >= 1 domain: gcrf_segments' input */
int gcrf_table_gene_coordinates(const gcrf_table *table, int64_t *start /* [G] or NULL */, int64_t *end /* [G] or NULL */);

/*
 * The CSR batch of gcrf_marginals_windowed.  attr_names[a] is the model's name of attribute a.  feature_type 0
 * ("protein"): one row per gene holding the set of its known domain names, first occurrence first; 1 ("domain"):
 * one row per domain, one empty row per domain-less gene.  The returned arrays belong to the table and stay
 * valid until the next gcrf_table_pack / gcrf_table_destroy; contig_ptr indexes rows.
 */
int gcrf_table_pack(gcrf_table *table, const char *const *attr_names, int32_t A, int32_t feature_type,
                    const int32_t **contig_ptr, const int32_t **row_ptr, const int32_t **attr_idx, int64_t *rows,
                    int64_t *nnz);
/*
 * The same batch with the feature extraction left to the device (GCRF_FLAG_ACCESSIONS): `accession` holds one entry
 * per kept domain row — the number behind a "PF" prefix (PF00394 -> 394), -1 for any other name — in the table's
 * domain order, nothing looked up or de-duplicated on the host (gecco/crf/features.py:13-35 then runs in
 * gcrf::features_kernel).  Same feature types, ownership and row layout as gcrf_table_pack; for models whose
 * attributes are all "PF" + `digits` ASCII digits (gcrf_model_set_vocabulary): the reference compares NAMES, so a
 * domain called PF394 is not PF00394 — only names with exactly `digits` digits count (0 = any number of digits).
 */
int gcrf_table_pack_accessions(gcrf_table *table, int32_t feature_type, int32_t digits, const int32_t **contig_ptr,
                               const int32_t **row_ptr, const int32_t **accession, int64_t *rows, int64_t *nnz);
const int32_t *gcrf_table_row_gene(const gcrf_table *table); /* [rows] gene of every packed row */

/*
 * Results.  row_prob[rows] is what gcrf_marginals_windowed returned for the packed batch (NaN = no probability;
 * NULL = none at all).  Per gene: average_p / max_p as Gene.average_probability / maximum_probability compute them
 * (gecco/model.py:274-290).  The writers produce the reference's genes / features tables: schema column order, NaN
 * as an empty field, an all-NaN probability column left out, floats in shortest round-trip form laid out like
 * Python's repr() — byte-identical to the reference's committed result tables when given the same numbers.
 */
int gcrf_table_gene_probabilities(const gcrf_table *table, const double *row_prob, double *average_p, double *max_p);
int gcrf_table_write_genes(const gcrf_table *table, const double *row_prob, const char *path);
int gcrf_table_write_features(const gcrf_table *table, const double *row_prob, const char *path);
/*
 * The clusters table for segments found by gcrf_segments on this table's genes (seg_* as that call returns them;
 * n_segments entries): ClusterTable.from_clusters(...).dump for clusters without a predicted type
 * (gecco/model.py:735-771): cluster_id = "<sequence_id>_cluster_<ordinal>" (gecco/refine.py:200), start / end = the
 * extreme gene coordinates, average_p / max_p over the member genes (gecco/model.py:442-455), proteins and domains
 * sorted and ";"-joined, type "Unknown".  The type classifier (gecco/types/) is outside this library.
 */
int gcrf_table_write_clusters(const gcrf_table *table, const double *row_prob, const int32_t *seg_contig,
                              const int32_t *seg_begin, const int32_t *seg_end, const int32_t *seg_ordinal,
                              int64_t n_segments, const char *path);

#ifdef __cplusplus
}
#endif
#endif /* GECCO_CRF_B200_H */
