// gcrf_abi.cu — the extern "C" surface declared in include/gecco_crf_b200.h.
// Owns the model handle (device weight table, stream, grow-only staging buffers) and turns the
// host- or device-pointer CSR batch into kernel launches.  No CPU fallback exists on purpose.
#include "../../include/gecco_crf_b200.h"
#include "gcrf_kernels.cuh"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges show up in Nsight Systems / ncu --nvtx, cost nothing otherwise

// internal: a nested call (wire batch -> regular entry point) keeps the timed region its caller opened
#define GCRF_FLAG_KEEP_TIMING 0x40000000u
// internal: the call carries peer output arrays (gcrf_marginals_windowed_peers)
#define GCRF_FLAG_HAS_PEERS 0x20000000u
// internal: no local output array (results go to the peer arrays only)
#define GCRF_FLAG_NO_LOCAL_OUT 0x10000000u
// internal: the arrays are a contig-aligned slice of a larger batch whose first gene is gcrf_model::slice_gene_base
// (contig_ptr VALUES keep counting from the start of the whole batch, see CsrDev::gene_base)
#define GCRF_FLAG_SLICE 0x08000000u

namespace {

thread_local char g_error[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

int fail_cuda(cudaError_t err, const char *what) {
    const int code = (err == cudaErrorMemoryAllocation) ? GCRF_ENOMEM : GCRF_ECUDA;
    return fail(code, "%s: %s (%s)", what, cudaGetErrorString(err), cudaGetErrorName(err));
}

#define GCRF_CUDA(call)                                        \
    do {                                                       \
        cudaError_t err__ = (call);                            \
        if (err__ != cudaSuccess) return fail_cuda(err__, #call); \
    } while (0)

// NVTX range over a scope: stage (validation + H2D), launch (kernels), finish (D2H + sync) of every ABI call
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

struct DeviceBuffer {
    void *ptr = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;  // slack so slightly larger batches do not reallocate
        cudaError_t err = cudaMalloc(&ptr, want);
        if (err != cudaSuccess) {
            want = bytes;
            err = cudaMalloc(&ptr, want);
        }
        if (err == cudaSuccess) cap = want;
        return err;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

}  // namespace

struct gcrf_model {
    int device = 0;
    int num_sms = 0;
    int32_t A = 0;
    gcrf::ModelDev dev{};
    float *d_table = nullptr;
    int32_t *d_table_fx = nullptr;
    int32_t *d_lut = nullptr;  // accession -> attribute id (gcrf_model_set_vocabulary)
    int32_t lut_size = 0;
    double *d_table64 = nullptr;
    double m01 = 0, m10 = 0, m11 = 0;
    double *d_state_w = nullptr;  // [A][2] raw state weights (GCRF_FLAG_F64)
    double exp_trans[4] = {0, 0, 0, 0};
    int32_t pos_label = 1;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    // overlapped host path of gcrf_marginals_windowed: results of slice k travel back on copy_stream while the
    // inputs of slice k+1 arrive on `stream` (PCIe is full duplex; the two directions have their own copy engines)
    static constexpr int kMaxSlices = 8;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_slice[kMaxSlices] = {};
    // gcrf_marginals_windowed_wire with several slices: the copies in run on their own stream as well, so that the
    // kernels of slice k (stream) overlap the copy in of slice k+1 (in_stream) and the copy back of slice k-1
    cudaStream_t in_stream = nullptr;
    cudaEvent_t ev_in[kMaxSlices] = {};
    bool timed = false;
    bool timing = false;  // record events around the kernels (gcrf_model_set_timing)
    bool ev_open = false; // ev_start already recorded by this call (in front of a widening / feature-extraction kernel)
    int64_t launches = 0;
    DeviceBuffer b_contig, b_gene, b_attr, b_out, b_scratch;
    DeviceBuffer b_ann, b_seg;  // gcrf_segments: annotation marks, outputs + count
    DeviceBuffer b_unary, b_pool, b_work;  // GCRF_FLAG_F64: exp of the state scores, max-pool (float output), work area
    void *peer_out[gcrf::WindowedArgs::kMaxPeers] = {};  // gcrf_marginals_windowed_peers: valid during that call only
    int32_t n_peer_out = 0, peer_multicast = 0;
    int64_t slice_gene_base = 0, slice_ids = 0;  // GCRF_FLAG_SLICE
    DeviceBuffer b_wire;        // gcrf_marginals_windowed_wire: the block as it came over PCIe
    DeviceBuffer b_idx16;       // GCRF_FLAG_IDX_U16, host buffers: the compact ids as they came over PCIe
    DeviceBuffer b_acc;         // GCRF_FLAG_ACCESSIONS, host buffers: the accessions as they came over PCIe
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// gene_ptr non-decreasing: branch-free OR-reduction of the sign of every difference (a pass at memory speed:
// ~1 ms for the 8 MB of config 2, against a PCIe-bound call of several ms)
template <typename T>
bool non_decreasing(const T *p, int64_t n) {
    T bad = 0;
    for (int64_t i = 0; i < n; ++i) bad |= (T)(p[i + 1] - p[i]);
    return bad >= 0;
}

// Start of the timed region of a call (gcrf_model_set_timing): in front of the FIRST kernel it launches, which may be
// the uint16 widening or the feature extraction inside stage_batch.
cudaError_t timing_begin(gcrf_model *m) {
    if (!m->timing || m->ev_open) return cudaSuccess;
    m->ev_open = true;
    return cudaEventRecord(m->ev_start, m->stream);
}
cudaError_t timing_end(gcrf_model *m) {
    m->timed = m->timing;
    if (!m->timing) return cudaSuccess;
    m->ev_open = false;
    return cudaEventRecord(m->ev_stop, m->stream);
}

// Host-side sanity checks of a CSR batch given as host pointers: O(C + G), so that no malformed pointer array can
// send the kernels out of bounds.
int check_host_csr(const int32_t *contig_ptr, const void *gene_ptr, bool ptr64, int64_t C, int64_t G, int64_t nnz) {
    if (contig_ptr[0] != 0 || contig_ptr[C] != G) return fail(GCRF_EINVAL, "contig_ptr must start at 0 and end at G");
    for (int64_t c = 0; c < C; ++c)
        if (contig_ptr[c + 1] <= contig_ptr[c]) return fail(GCRF_EINVAL, "contig_ptr must be strictly increasing (contig %lld is empty)", (long long)c);
    const int64_t first = ptr64 ? static_cast<const int64_t *>(gene_ptr)[0] : static_cast<const int32_t *>(gene_ptr)[0];
    const int64_t last = ptr64 ? static_cast<const int64_t *>(gene_ptr)[G] : static_cast<const int32_t *>(gene_ptr)[G];
    if (first != 0 || last != nnz) return fail(GCRF_EINVAL, "gene_ptr must start at 0 and end at nnz");
    const bool mono = ptr64 ? non_decreasing(static_cast<const int64_t *>(gene_ptr), G) : non_decreasing(static_cast<const int32_t *>(gene_ptr), G);
    if (!mono) {
        for (int64_t g = 0; g < G; ++g) {
            const int64_t a = ptr64 ? static_cast<const int64_t *>(gene_ptr)[g] : static_cast<const int32_t *>(gene_ptr)[g];
            const int64_t b = ptr64 ? static_cast<const int64_t *>(gene_ptr)[g + 1] : static_cast<const int32_t *>(gene_ptr)[g + 1];
            if (b < a) return fail(GCRF_EINVAL, "gene_ptr must be non-decreasing (gene %lld: %lld > %lld)", (long long)g, (long long)a, (long long)b);
        }
    }
    return GCRF_OK;
}

struct Batch {
    gcrf::CsrDev csr{};
    void *d_out = nullptr;
    size_t out_bytes = 0;
    bool device_ptrs = false;
};

// Validates the common arguments and, in host-pointer mode, stages the inputs on the device.
int stage_batch(gcrf_model *m, const int32_t *contig_ptr, const void *gene_ptr, const void *attr_idx_any,
                int64_t C, int64_t G, int64_t nnz, void *out, uint32_t flags, Batch *b, bool defer_copies = false) {
    NvtxRange range("gcrf:stage");
    const int32_t *attr_idx = static_cast<const int32_t *>(attr_idx_any);
    const bool idx16 = (flags & GCRF_FLAG_IDX_U16) != 0;
    const bool accessions = (flags & GCRF_FLAG_ACCESSIONS) != 0;
    if (idx16 && m && m->A >= 0xFFFF) return fail(GCRF_EINVAL, "GCRF_FLAG_IDX_U16 needs a model with fewer than 65535 attributes");
    if (!m) return fail(GCRF_EINVAL, "model handle is NULL");
    if (accessions && idx16) return fail(GCRF_EINVAL, "GCRF_FLAG_ACCESSIONS and GCRF_FLAG_IDX_U16 exclude each other");
    if (accessions && !m->d_lut && m->A > 0) return fail(GCRF_EINVAL, "gcrf_model_set_vocabulary has not been called");
    if (accessions && defer_copies) return fail(GCRF_EINVAL, "GCRF_FLAG_ACCESSIONS cannot be combined with sliced copies");
    if (C < 0 || G < 0 || nnz < 0) return fail(GCRF_EINVAL, "negative size");
    if (G > 0x7fffffff - 1024) return fail(GCRF_EINVAL, "G must fit in int32 (shard the batch)");
    if ((C == 0) != (G == 0)) return fail(GCRF_EINVAL, "C and G must both be zero or both be positive");
    if (C > G) return fail(GCRF_EINVAL, "more contigs than genes");
    if (G > 0 && (!contig_ptr || !gene_ptr || !out)) return fail(GCRF_EINVAL, "NULL array");
    if (nnz > 0 && !attr_idx) return fail(GCRF_EINVAL, "attr_idx is NULL");
    const bool ptr64 = (flags & GCRF_FLAG_PTR64) != 0;
    if (!ptr64 && nnz > 0x7fffffff) return fail(GCRF_EINVAL, "nnz >= 2^31 needs GCRF_FLAG_PTR64");
    b->device_ptrs = (flags & GCRF_FLAG_DEVICE_PTRS) != 0;
    b->out_bytes = (size_t)G * ((flags & GCRF_FLAG_OUT_F32) ? sizeof(float) : sizeof(double));
    b->csr.C = C;
    b->csr.G = G;
    b->csr.nnz = nnz;
    if (G == 0) return GCRF_OK;

    if (b->device_ptrs) {
        if ((reinterpret_cast<uintptr_t>(attr_idx) & 15u) != 0)
            return fail(GCRF_EINVAL, "device attr_idx must be 16-byte aligned");
        if (idx16) {  // widen into the library's own buffer
            GCRF_CUDA(m->b_attr.reserve((size_t)(nnz > 0 ? nnz : 1) * 4 + 64));
            GCRF_CUDA(timing_begin(m));
            cudaError_t werr = gcrf::launch_widen_u16(static_cast<const uint16_t *>(attr_idx_any), static_cast<int32_t *>(m->b_attr.ptr),
                                                      nnz, m->num_sms, m->stream, &m->launches);
            if (werr != cudaSuccess) return fail_cuda(werr, "launch_widen_u16");
            attr_idx = static_cast<const int32_t *>(m->b_attr.ptr);
        }
        if (accessions && nnz > 0) {  // accession -> attribute id, repeats inside a gene dropped, into the library's buffer
            GCRF_CUDA(m->b_attr.reserve((size_t)nnz * 4 + 64));
            GCRF_CUDA(timing_begin(m));
            cudaError_t ferr = gcrf::launch_features(attr_idx, ptr64 ? nullptr : static_cast<const int32_t *>(gene_ptr),
                                                     ptr64 ? static_cast<const int64_t *>(gene_ptr) : nullptr, G, nnz, m->d_lut,
                                                     m->lut_size, m->A, static_cast<int32_t *>(m->b_attr.ptr), m->num_sms,
                                                     m->stream, &m->launches);
            if (ferr != cudaSuccess) return fail_cuda(ferr, "launch_features");
            attr_idx = static_cast<const int32_t *>(m->b_attr.ptr);
        }
        b->csr.contig_ptr = contig_ptr;
        b->csr.gene_ptr32 = ptr64 ? nullptr : static_cast<const int32_t *>(gene_ptr);
        b->csr.gene_ptr64 = ptr64 ? static_cast<const int64_t *>(gene_ptr) : nullptr;
        b->csr.attr_idx = attr_idx;
        b->d_out = out;
        return GCRF_OK;
    }

    const int rc = check_host_csr(contig_ptr, gene_ptr, ptr64, C, G, nnz);
    if (rc != GCRF_OK) return rc;
    const size_t gene_bytes = (size_t)(G + 1) * (ptr64 ? 8 : 4);
    GCRF_CUDA(m->b_contig.reserve((size_t)(C + 1) * 4));
    GCRF_CUDA(m->b_gene.reserve(gene_bytes));
    GCRF_CUDA(m->b_attr.reserve((size_t)(nnz > 0 ? nnz : 1) * 4 + 64));
    GCRF_CUDA(m->b_out.reserve(b->out_bytes));
    if (!defer_copies) {
        GCRF_CUDA(cudaMemcpyAsync(m->b_contig.ptr, contig_ptr, (size_t)(C + 1) * 4, cudaMemcpyHostToDevice, m->stream));
        GCRF_CUDA(cudaMemcpyAsync(m->b_gene.ptr, gene_ptr, gene_bytes, cudaMemcpyHostToDevice, m->stream));
        if (nnz > 0 && idx16) {
            // half the PCIe bytes: the ids cross as uint16 and are widened on the device
            GCRF_CUDA(m->b_idx16.reserve((size_t)nnz * 2 + 16));
            GCRF_CUDA(cudaMemcpyAsync(m->b_idx16.ptr, attr_idx_any, (size_t)nnz * 2, cudaMemcpyHostToDevice, m->stream));
            GCRF_CUDA(timing_begin(m));
            cudaError_t werr = gcrf::launch_widen_u16(static_cast<const uint16_t *>(m->b_idx16.ptr), static_cast<int32_t *>(m->b_attr.ptr),
                                                      nnz, m->num_sms, m->stream, &m->launches);
            if (werr != cudaSuccess) return fail_cuda(werr, "launch_widen_u16");
        } else if (nnz > 0 && accessions) {
            GCRF_CUDA(m->b_acc.reserve((size_t)nnz * 4 + 16));
            GCRF_CUDA(cudaMemcpyAsync(m->b_acc.ptr, attr_idx, (size_t)nnz * 4, cudaMemcpyHostToDevice, m->stream));
            GCRF_CUDA(timing_begin(m));
            cudaError_t ferr = gcrf::launch_features(static_cast<const int32_t *>(m->b_acc.ptr),
                                                     ptr64 ? nullptr : static_cast<const int32_t *>(m->b_gene.ptr),
                                                     ptr64 ? static_cast<const int64_t *>(m->b_gene.ptr) : nullptr, G, nnz, m->d_lut,
                                                     m->lut_size, m->A, static_cast<int32_t *>(m->b_attr.ptr), m->num_sms,
                                                     m->stream, &m->launches);
            if (ferr != cudaSuccess) return fail_cuda(ferr, "launch_features");
        } else if (nnz > 0) {
            GCRF_CUDA(cudaMemcpyAsync(m->b_attr.ptr, attr_idx, (size_t)nnz * 4, cudaMemcpyHostToDevice, m->stream));
        }
    }
    b->csr.contig_ptr = static_cast<const int32_t *>(m->b_contig.ptr);
    b->csr.gene_ptr32 = ptr64 ? nullptr : static_cast<const int32_t *>(m->b_gene.ptr);
    b->csr.gene_ptr64 = ptr64 ? static_cast<const int64_t *>(m->b_gene.ptr) : nullptr;
    b->csr.attr_idx = static_cast<const int32_t *>(m->b_attr.ptr);
    b->d_out = m->b_out.ptr;
    return GCRF_OK;
}

int finish_batch(gcrf_model *m, const Batch &b, void *out) {
    if (b.device_ptrs || b.csr.G == 0) return GCRF_OK;
    NvtxRange range("gcrf:finish");
    GCRF_CUDA(cudaMemcpyAsync(out, b.d_out, b.out_bytes, cudaMemcpyDeviceToHost, m->stream));
    GCRF_CUDA(cudaStreamSynchronize(m->stream));
    return GCRF_OK;
}

}  // namespace

extern "C" {

int gcrf_version(void) { return GCRF_ABI_VERSION; }

const char *gcrf_last_error(void) { return g_error; }

int gcrf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int gcrf_model_create(const double *state_w, int32_t A, int32_t L, const double *trans_w, int32_t pos_label,
                      int32_t device, gcrf_model **out) {
    if (!out) return fail(GCRF_EINVAL, "out is NULL");
    *out = nullptr;
    if (!state_w || !trans_w || A < 0 || L <= 0) return fail(GCRF_EINVAL, "bad model arrays");
    if (L != 2) return fail(GCRF_EUNSUPPORTED, "the device path handles 2-label models (GECCO labels '0'/'1'), got L=%d", L);
    if (pos_label < 0 || pos_label >= L) return fail(GCRF_EINVAL, "pos_label out of range");
    int ndev = 0;
    cudaError_t err = cudaGetDeviceCount(&ndev);
    if (err != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(GCRF_ENODEVICE, "no CUDA device available (%s); this library has no CPU path",
                    err == cudaSuccess ? "device count is 0" : cudaGetErrorString(err));
    }
    if (device < 0 || device >= ndev) return fail(GCRF_ENODEVICE, "device %d out of range (have %d)", device, ndev);
    cudaDeviceProp prop;
    GCRF_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(GCRF_ENODEVICE, "device %d is sm_%d%d; this build targets sm_100a (B200)", device, prop.major, prop.minor);

    // fold the 2-label model (see gcrf_kernels.cuh): p = pos_label, o = the other label
    const int p = pos_label, o = 1 - pos_label;
    const double too = trans_w[o * 2 + o];
    const double m01 = std::exp(trans_w[o * 2 + p] - too);
    const double m10 = std::exp(trans_w[p * 2 + o] - too);
    const double m11 = std::exp(trans_w[p * 2 + p] - too);
    if (!std::isfinite(m01) || !std::isfinite(m10) || !std::isfinite(m11) || m01 <= 0 || m10 <= 0 || m11 <= 0)
        return fail(GCRF_EUNSUPPORTED, "transition weights out of FP32 range");
    // Odds stay finite in FP32 when f_max * max(m11,1) * e^{2 clamp} < ~e^85, f_max = sup of the
    // message map (m01 + r m11)/(1 + r m10) and of its backward twin.
    const double fmax = std::fmax(std::fmax(m01, m11 / m10), std::fmax(m10, m11 / m01));
    double clamp = 0.5 * (85.0 - std::log(std::fmax(fmax, 1.0)) - std::log(std::fmax(m11, 1.0)));
    if (clamp > 30.0) clamp = 30.0;
    if (clamp < 16.0) return fail(GCRF_EUNSUPPORTED, "transition weights too extreme for the FP32 device path");

    std::vector<float> table((size_t)A + 1);
    std::vector<double> table64((size_t)A + 1);
    for (int32_t a = 0; a < A; ++a) {
        const double d = state_w[(size_t)a * 2 + p] - state_w[(size_t)a * 2 + o];
        if (!std::isfinite(d)) return fail(GCRF_EINVAL, "non-finite state weight for attribute %d", a);
        table[a] = (float)d;
        table64[a] = d;
    }
    table[A] = 0.0f;
    table64[A] = 0.0;
    // fixed-point table: as many fraction bits as keep rows of up to 64 ids exact in wrapping int32 sums
    double maxabs = 0.0;
    for (int32_t a = 0; a < A; ++a) maxabs = std::fmax(maxabs, std::fabs(table64[a]));
    int fx_bits = 24;
    while (fx_bits > 8 && maxabs * 64.0 * std::ldexp(1.0, fx_bits) >= 2147483647.0) --fx_bits;
    const double fx_one = std::ldexp(1.0, fx_bits);
    std::vector<int32_t> table_fx((size_t)A + 4, 0);  // padded: the kernel bulk-copies it in 16-byte units
    for (int32_t a = 0; a < A; ++a) table_fx[a] = (int32_t)std::llround(table64[a] * fx_one);
    table_fx[A] = 0;
    const double fx_nsafe_d = maxabs > 0 ? std::floor(2147483647.0 / (maxabs * fx_one + 1.0)) : 2147483647.0;
    const int32_t fx_nsafe = (int32_t)std::fmin(fx_nsafe_d, 2147483647.0);

    gcrf_model *m = new (std::nothrow) gcrf_model();
    if (!m) return fail(GCRF_ENOMEM, "out of host memory");
    m->device = device;
    m->num_sms = prop.multiProcessorCount;
    m->A = A;
    DeviceGuard guard(device);
    auto cleanup = [&](int code) {
        gcrf_model_destroy(m);
        return code;
    };
    if ((err = cudaMalloc(reinterpret_cast<void **>(&m->d_table), table.size() * sizeof(float))) != cudaSuccess)
        return cleanup(fail_cuda(err, "cudaMalloc(table)"));
    if ((err = cudaMemcpy(m->d_table, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess)
        return cleanup(fail_cuda(err, "cudaMemcpy(table)"));
    if ((err = cudaMalloc(reinterpret_cast<void **>(&m->d_table_fx), table_fx.size() * sizeof(int32_t))) != cudaSuccess)
        return cleanup(fail_cuda(err, "cudaMalloc(table_fx)"));
    if ((err = cudaMemcpy(m->d_table_fx, table_fx.data(), table_fx.size() * sizeof(int32_t), cudaMemcpyHostToDevice)) != cudaSuccess)
        return cleanup(fail_cuda(err, "cudaMemcpy(table_fx)"));
    if ((err = cudaMalloc(reinterpret_cast<void **>(&m->d_table64), table64.size() * sizeof(double))) != cudaSuccess)
        return cleanup(fail_cuda(err, "cudaMalloc(table64)"));
    if ((err = cudaMemcpy(m->d_table64, table64.data(), table64.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess)
        return cleanup(fail_cuda(err, "cudaMemcpy(table64)"));
    m->m01 = m01;
    m->m10 = m10;
    m->m11 = m11;
    if (A > 0) {
        if ((err = cudaMalloc(reinterpret_cast<void **>(&m->d_state_w), (size_t)A * 2 * sizeof(double))) != cudaSuccess)
            return cleanup(fail_cuda(err, "cudaMalloc(state_w)"));
        if ((err = cudaMemcpy(m->d_state_w, state_w, (size_t)A * 2 * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess)
            return cleanup(fail_cuda(err, "cudaMemcpy(state_w)"));
    }
    for (int k = 0; k < 4; ++k) m->exp_trans[k] = std::exp(trans_w[k]);  // CRFsuite exponentiates the raw transition scores
    m->pos_label = pos_label;
    if ((err = cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return cleanup(fail_cuda(err, "cudaStreamCreate"));
    m->stream = m->own_stream;
    if ((err = cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return cleanup(fail_cuda(err, "cudaStreamCreate"));
    if ((err = cudaStreamCreateWithFlags(&m->in_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return cleanup(fail_cuda(err, "cudaStreamCreate"));
    for (int k = 0; k < gcrf_model::kMaxSlices; ++k) {
        if ((err = cudaEventCreateWithFlags(&m->ev_slice[k], cudaEventDisableTiming)) != cudaSuccess)
            return cleanup(fail_cuda(err, "cudaEventCreate"));
        if ((err = cudaEventCreateWithFlags(&m->ev_in[k], cudaEventDisableTiming)) != cudaSuccess)
            return cleanup(fail_cuda(err, "cudaEventCreate"));
    }
    if ((err = cudaEventCreate(&m->ev_start)) != cudaSuccess) return cleanup(fail_cuda(err, "cudaEventCreate"));
    if ((err = cudaEventCreate(&m->ev_stop)) != cudaSuccess) return cleanup(fail_cuda(err, "cudaEventCreate"));
    m->dev.table = m->d_table;
    m->dev.A = A;
    m->dev.m01 = (float)m01;
    m->dev.m10 = (float)m10;
    m->dev.m11 = (float)m11;
    m->dev.clamp = (float)clamp;
    m->dev.table_fx = m->d_table_fx;
    m->dev.fx_bits = fx_bits;
    m->dev.fx_nsafe = fx_nsafe;
    *out = m;
    return GCRF_OK;
}

void gcrf_model_destroy(gcrf_model *m) {
    if (!m) return;
    DeviceGuard guard(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    m->b_contig.release();
    m->b_gene.release();
    m->b_attr.release();
    m->b_out.release();
    m->b_scratch.release();
    m->b_ann.release();
    m->b_seg.release();
    m->b_idx16.release();
    m->b_acc.release();
    m->b_wire.release();
    m->b_unary.release();
    m->b_pool.release();
    m->b_work.release();
    if (m->d_state_w) cudaFree(m->d_state_w);
    if (m->d_table) cudaFree(m->d_table);
    if (m->d_table64) cudaFree(m->d_table64);
    if (m->d_table_fx) cudaFree(m->d_table_fx);
    if (m->d_lut) cudaFree(m->d_lut);
    if (m->ev_start) cudaEventDestroy(m->ev_start);
    if (m->ev_stop) cudaEventDestroy(m->ev_stop);
    for (int k = 0; k < gcrf_model::kMaxSlices; ++k) {
        if (m->ev_slice[k]) cudaEventDestroy(m->ev_slice[k]);
        if (m->ev_in[k]) cudaEventDestroy(m->ev_in[k]);
    }
    if (m->in_stream) {
        cudaStreamSynchronize(m->in_stream);
        cudaStreamDestroy(m->in_stream);
    }
    if (m->copy_stream) {
        cudaStreamSynchronize(m->copy_stream);
        cudaStreamDestroy(m->copy_stream);
    }
    if (m->own_stream) cudaStreamDestroy(m->own_stream);
    delete m;
}

int gcrf_model_set_stream(gcrf_model *m, void *cuda_stream) {
    if (!m) return fail(GCRF_EINVAL, "model handle is NULL");
    m->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : m->own_stream;
    return GCRF_OK;
}

int gcrf_model_synchronize(gcrf_model *m) {
    if (!m) return fail(GCRF_EINVAL, "model handle is NULL");
    DeviceGuard guard(m->device);
    GCRF_CUDA(cudaStreamSynchronize(m->stream));
    return GCRF_OK;
}

}  // extern "C"

namespace {

// Plans and enqueues the windowed kernel for one (sub-)batch on the handle's stream.
int launch_windowed_path(gcrf_model *m, gcrf::WindowedArgs &args, bool prof) {
    NvtxRange range("gcrf:launch windowed");
    gcrf::WindowedPlan plan{};
    // GCRF_FORCE_GENERIC=1 routes W=20 through the generic kernel too (A/B testing of the two device paths)
    const char *force = getenv("GCRF_FORCE_GENERIC");
    const bool want_generic = force && force[0] == '1';
    const bool fast = gcrf::stream_supported(args) && !want_generic;
    if (!fast && args.n_peer_out > 0)
        return fail(GCRF_EUNSUPPORTED, "peer output arrays need the streaming kernel's peer-store variant (window 5 or 20, FP32 arithmetic)");
    cudaError_t err = fast ? gcrf::plan_stream(args, m->num_sms, &plan) : gcrf::plan_windowed(args, m->num_sms, &plan);
    if (err == cudaErrorInvalidValue) {
        cudaGetLastError();
        return fail(GCRF_EUNSUPPORTED, "window size %d with %d attributes does not fit the FP32 kernels' shared memory (largest window: %d); "
                    "GCRF_FLAG_F64 accepts any window", args.window, m->A, gcrf::windowed_max_window(m->A));
    }
    if (err != cudaSuccess) return fail_cuda(err, "plan_windowed");
    GCRF_CUDA(timing_begin(m));
    err = fast ? gcrf::launch_stream(args, plan, m->stream, &m->launches)
               : gcrf::launch_windowed(args, plan, m->stream, &m->launches);
    if (err != cudaSuccess) return fail_cuda(err, "launch_windowed");
    GCRF_CUDA(timing_end(m));
    if (prof) {
        unsigned long long h[16];
        GCRF_CUDA(cudaMemcpyAsync(h, args.prof, sizeof(h), cudaMemcpyDeviceToHost, m->stream));
        GCRF_CUDA(cudaStreamSynchronize(m->stream));
        static const char *names[10] = {"setup", "wait_ids", "walk", "unary", "contig", "short", "dp", "pool", "loop", "-"};
        unsigned long long tot = 0;
        for (int k = 0; k < 9; ++k) tot += h[k];
        fprintf(stderr, "[gcrf phases] ctas=%llu grid=%d tiles/cta=%d:", h[15], plan.grid, plan.tiles_per_cta);
        for (int k = 0; k < 9; ++k) fprintf(stderr, " %s=%.1f%%", names[k], tot ? 100.0 * h[k] / tot : 0.0);
        fprintf(stderr, " | cycles/cta=%.0f\n", h[15] ? (double)tot / h[15] : 0.0);
    }
    return GCRF_OK;
}

// Host-buffer batches large enough to be PCIe-bound are cut into contig-aligned slices and pipelined over three
// streams: while slice k runs through the kernel (stream), slice k+1's CSR arrays are on their way in (in_stream) and
// slice k-1's marginals travel back to the host (copy_stream).  Contigs are independent, so every slice is a plain
// kernel launch over its own contigs (CsrDev::gene_base).  The slices shrink 8:4:2:1 — what the pipeline adds to the
// copy in, the long pole, is the last slice's kernel and copy back.
int windowed_sliced(gcrf_model *m, gcrf::WindowedArgs args, const int32_t *contig_ptr, const void *gene_ptr,
                    const int32_t *attr_idx, void *out, int slices) {
    const gcrf::CsrDev whole = args.csr;
    const bool ptr64 = whole.gene_ptr64 != nullptr;
    const size_t psz = ptr64 ? 8 : 4, osz = args.out_f32 ? 4 : 8;
    auto row = [&](int64_t g) -> int64_t {
        return ptr64 ? static_cast<const int64_t *>(gene_ptr)[g] : static_cast<const int32_t *>(gene_ptr)[g];
    };
    // slice boundaries: contigs, balanced by bytes moved (4 per id, 4 + 8 per gene)
    auto cost_at = [&](int64_t c) -> double { return 4.0 * (double)row(contig_ptr[c]) + 12.0 * (double)contig_ptr[c]; };
    const double total = cost_at(whole.C);
    int64_t cut[gcrf_model::kMaxSlices + 1];
    cut[0] = 0;
    for (int k = 1; k < slices; ++k) {
        const double want = total * (1.0 - (double)((1 << (slices - k)) - 1) / (double)((1 << slices) - 1));
        int64_t lo = cut[k - 1], hi = whole.C;
        while (lo < hi) {
            const int64_t mid = (lo + hi) / 2;
            if (cost_at(mid) < want) lo = mid + 1; else hi = mid;
        }
        cut[k] = lo;
    }
    cut[slices] = whole.C;
    char *d_contig = static_cast<char *>(m->b_contig.ptr), *d_gene = static_cast<char *>(m->b_gene.ptr);
    char *d_attr = static_cast<char *>(m->b_attr.ptr), *d_out = static_cast<char *>(m->b_out.ptr);
    int used = 0;
    cudaStream_t in = m->in_stream;
    GCRF_CUDA(cudaEventRecord(m->ev_slice[0], m->stream));  // work queued on the caller's stream stays ahead of the copies
    GCRF_CUDA(cudaStreamWaitEvent(in, m->ev_slice[0], 0));
    // the small arrays go first, whole: one copy each instead of one per slice
    GCRF_CUDA(cudaMemcpyAsync(d_contig, contig_ptr, (size_t)(whole.C + 1) * 4, cudaMemcpyHostToDevice, in));
    for (int k = 0; k < slices; ++k) {
        const int64_t c0 = cut[k], c1 = cut[k + 1];
        if (c1 <= c0) continue;
        const int64_t g0 = contig_ptr[c0], g1 = contig_ptr[c1];
        const int64_t p0 = row(g0), p1 = row(g1);
        // every array lands where the whole batch would have put it, so all indices stay valid as they are
        GCRF_CUDA(cudaMemcpyAsync(d_gene + (size_t)g0 * psz, static_cast<const char *>(gene_ptr) + (size_t)g0 * psz,
                                  (size_t)(g1 - g0 + 1) * psz, cudaMemcpyHostToDevice, in));
        if (p1 > p0)
            GCRF_CUDA(cudaMemcpyAsync(d_attr + (size_t)p0 * 4, attr_idx + p0, (size_t)(p1 - p0) * 4, cudaMemcpyHostToDevice, in));
        GCRF_CUDA(cudaEventRecord(m->ev_in[used], in));
        GCRF_CUDA(cudaStreamWaitEvent(m->stream, m->ev_in[used], 0));
        args.csr = whole;
        args.csr.contig_ptr = reinterpret_cast<const int32_t *>(d_contig) + c0;
        args.csr.gene_ptr32 = ptr64 ? nullptr : reinterpret_cast<const int32_t *>(d_gene) + g0;
        args.csr.gene_ptr64 = ptr64 ? reinterpret_cast<const int64_t *>(d_gene) + g0 : nullptr;
        args.csr.C = c1 - c0;
        args.csr.G = g1 - g0;
        args.csr.slice_ids = p1 - p0;  // the slice's own density picks the tile size
        args.csr.gene_base = g0;
        args.out = d_out + (size_t)g0 * osz;
        const int rc = launch_windowed_path(m, args, false);
        if (rc != GCRF_OK) return rc;
        GCRF_CUDA(cudaEventRecord(m->ev_slice[used], m->stream));
        GCRF_CUDA(cudaStreamWaitEvent(m->copy_stream, m->ev_slice[used], 0));
        GCRF_CUDA(cudaMemcpyAsync(static_cast<char *>(out) + (size_t)g0 * osz, d_out + (size_t)g0 * osz, (size_t)(g1 - g0) * osz,
                                  cudaMemcpyDeviceToHost, m->copy_stream));
        ++used;
    }
    GCRF_CUDA(cudaStreamSynchronize(m->in_stream));
    GCRF_CUDA(cudaStreamSynchronize(m->copy_stream));
    GCRF_CUDA(cudaStreamSynchronize(m->stream));
    return GCRF_OK;
}

}  // namespace

extern "C" {

int gcrf_marginals_windowed(gcrf_model *m, const int32_t *contig_ptr, const void *gene_ptr, const void *attr_idx,
                            int64_t C, int64_t G, int64_t nnz, int32_t window, int32_t step, int32_t pad, void *out,
                            uint32_t flags) {
    if (!m) return fail(GCRF_EINVAL, "model handle is NULL");
    if (!(flags & GCRF_FLAG_KEEP_TIMING)) m->ev_open = false;
    const bool has_peers = (flags & GCRF_FLAG_HAS_PEERS) != 0, no_local_out = (flags & GCRF_FLAG_NO_LOCAL_OUT) != 0;
    const int64_t gene_base = (flags & GCRF_FLAG_SLICE) ? m->slice_gene_base : 0;
    const int64_t slice_ids = (flags & GCRF_FLAG_SLICE) ? m->slice_ids : 0;
    flags &= ~(uint32_t)(GCRF_FLAG_KEEP_TIMING | GCRF_FLAG_HAS_PEERS | GCRF_FLAG_NO_LOCAL_OUT | GCRF_FLAG_SLICE);
    // gecco/_meta.py:127-130
    if (window <= 0) return fail(GCRF_EINVAL, "Window size must be strictly positive");
    if (step <= 0 || step > window) return fail(GCRF_EINVAL, "Window step must be strictly positive and under `window_size`");
    DeviceGuard guard(m->device);
    // GCRF_PHASE_PROFILE=1: per-phase cycle counters of the streaming kernel, printed to stderr (tuning aid)
    const char *prof_env = getenv("GCRF_PHASE_PROFILE");
    const bool prof = prof_env && prof_env[0] == '1';
    // host buffers, PCIe-bound size: overlap the two copy directions over contig-aligned slices
    int slices = 1;
    if (!(flags & (GCRF_FLAG_DEVICE_PTRS | GCRF_FLAG_IDX_U16 | GCRF_FLAG_ACCESSIONS | GCRF_FLAG_F64)) && !m->timing && !prof && C >= 2) {
        const double bytes = 4.0 * (double)nnz + 12.0 * (double)G;
        const char *env = getenv("GCRF_HOST_SLICES");  // tuning / A-B: 1 turns the overlap off
        // (round 1 measured "no gain" from slicing — with the copies in queued behind the kernels on one stream;
        // numbers of the three-stream pipeline: profiles/r2_e2e_wire_slices.txt)
        slices = env ? atoi(env) : bytes >= 24e6 ? 4 : bytes >= 8e6 ? 2 : 1;
        if (slices > gcrf_model::kMaxSlices) slices = gcrf_model::kMaxSlices;
        if (slices > C) slices = (int)C;
        if (slices < 1) slices = 1;
    }
    Batch b;
    int rc = stage_batch(m, contig_ptr, gene_ptr, attr_idx, C, G, nnz, out, flags, &b, slices > 1);
    if (rc != GCRF_OK || G == 0) return rc;

    if ((flags & GCRF_FLAG_F64) && has_peers)
        return fail(GCRF_EUNSUPPORTED, "peer output arrays need the streaming kernel's peer-store variant (window 5 or 20, FP32 arithmetic)");
    if (flags & GCRF_FLAG_F64) {
        // the reference's own arithmetic (gcrf_exact.cu)
        gcrf::ExactArgs ex{};
        ex.csr = b.csr;
        ex.csr.gene_base = gene_base;
        ex.A = m->A;
        ex.state_w = m->d_state_w;
        for (int k = 0; k < 4; ++k) ex.M[k] = m->exp_trans[k];
        ex.pos_label = m->pos_label;
        ex.window = window;
        ex.step = step;
        ex.pad = pad ? 1 : 0;
        ex.out = b.d_out;
        ex.out_f32 = (flags & GCRF_FLAG_OUT_F32) ? 1 : 0;
        GCRF_CUDA(m->b_unary.reserve((size_t)G * 2 * sizeof(double)));
        ex.unary = static_cast<double *>(m->b_unary.ptr);
        if (ex.out_f32) {
            GCRF_CUDA(m->b_pool.reserve((size_t)G * sizeof(double)));
            ex.pool = static_cast<double *>(m->b_pool.ptr);
        }
        const size_t work_bytes = gcrf::exact_work_bytes(G, window, m->num_sms);
        if (work_bytes) GCRF_CUDA(m->b_work.reserve(work_bytes));
        NvtxRange range("gcrf:launch exact f64");
        GCRF_CUDA(timing_begin(m));
        cudaError_t err = gcrf::launch_exact(ex, static_cast<double *>(m->b_work.ptr), m->num_sms, m->stream, &m->launches);
        if (err != cudaSuccess) return fail_cuda(err, "launch_exact");
        GCRF_CUDA(timing_end(m));
        return finish_batch(m, b, out);
    }

    gcrf::WindowedArgs args{};
    args.model = m->dev;
    args.csr = b.csr;
    args.csr.gene_base = gene_base;
    args.csr.slice_ids = slice_ids;
    args.out = no_local_out ? nullptr : b.d_out;
    args.out_f32 = (flags & GCRF_FLAG_OUT_F32) ? 1 : 0;
    args.window = window;
    args.step = step;
    args.pad = pad ? 1 : 0;
    args.prof = nullptr;
    if (has_peers) {
        for (int k = 0; k < m->n_peer_out; ++k) args.peer_out[k] = m->peer_out[k];
        args.n_peer_out = m->n_peer_out;
        args.peer_multicast = m->peer_multicast;
    }
    {
        const char *skip = getenv("GCRF_DEBUG_SKIP");  // results are WRONG when set: timing experiments only
        args.debug_skip = skip ? atoi(skip) : 0;
    }
    if (prof) {
        GCRF_CUDA(m->b_scratch.reserve(16 * sizeof(unsigned long long)));
        GCRF_CUDA(cudaMemsetAsync(m->b_scratch.ptr, 0, 16 * sizeof(unsigned long long), m->stream));
        args.prof = static_cast<unsigned long long *>(m->b_scratch.ptr);
    }
    if (slices > 1) return windowed_sliced(m, args, contig_ptr, gene_ptr, static_cast<const int32_t *>(attr_idx), out, slices);
    rc = launch_windowed_path(m, args, prof);
    if (rc != GCRF_OK) return rc;
    return finish_batch(m, b, out);
}

}  // extern "C"

extern "C" {

int gcrf_marginals_windowed_peers(gcrf_model *m, const int32_t *contig_ptr, const void *gene_ptr, const void *attr_idx, int64_t C,
                                  int64_t G, int64_t nnz, int32_t window, int32_t step, int32_t pad, void *out,
                                  void *const *peer_out, int32_t n_peer_out, int64_t out_offset, uint32_t flags) {
    if (!m) return fail(GCRF_EINVAL, "model handle is NULL");
    if (n_peer_out < 0 || n_peer_out > gcrf::WindowedArgs::kMaxPeers)
        return fail(GCRF_EINVAL, "between 0 and %d peer output arrays", gcrf::WindowedArgs::kMaxPeers);
    if (n_peer_out > 0 && !peer_out) return fail(GCRF_EINVAL, "peer_out is NULL");
    if (out_offset < 0) return fail(GCRF_EINVAL, "negative out_offset");
    if (!(flags & GCRF_FLAG_DEVICE_PTRS)) return fail(GCRF_EINVAL, "peer output arrays are device memory: GCRF_FLAG_DEVICE_PTRS is required");
    const bool multicast = (flags & GCRF_FLAG_MULTICAST) != 0;
    if (multicast && n_peer_out != 1) return fail(GCRF_EINVAL, "GCRF_FLAG_MULTICAST takes exactly one (multicast) address");
    const size_t item = (flags & GCRF_FLAG_OUT_F32) ? 4 : 8;
    for (int k = 0; k < n_peer_out; ++k) {
        if (!peer_out[k]) return fail(GCRF_EINVAL, "peer_out[%d] is NULL", k);
        m->peer_out[k] = static_cast<char *>(peer_out[k]) + (size_t)out_offset * item;
    }
    m->n_peer_out = n_peer_out;
    m->peer_multicast = multicast ? 1 : 0;
    // `out` may be NULL here (results only go to the peer arrays); the regular entry point insists on an array
    static char dummy;
    const int rc = gcrf_marginals_windowed(m, contig_ptr, gene_ptr, attr_idx, C, G, nnz, window, step, pad, out ? out : &dummy,
                                           (flags & ~(uint32_t)GCRF_FLAG_MULTICAST) | GCRF_FLAG_HAS_PEERS | (out ? 0u : GCRF_FLAG_NO_LOCAL_OUT));
    m->n_peer_out = 0;
    return rc;
}

}  // extern "C"

extern "C" {

namespace {
// GCRF_WIRE_TRACE=1: a timeline of one sliced call on stderr (events on the three streams), for tools/wire_slices_probe.py
struct WireTrace {
    bool on = false;
    cudaEvent_t t0 = nullptr;
    struct Mark { cudaEvent_t ev; const char *what; int slice; };
    std::vector<Mark> marks;
    WireTrace() {
        static const bool enabled = getenv("GCRF_WIRE_TRACE") != nullptr;
        on = enabled;
    }
    void start(cudaStream_t s) {
        if (!on) return;
        cudaEventCreate(&t0);
        cudaEventRecord(t0, s);
    }

    void mark(cudaStream_t s, const char *what, int slice) {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        marks.push_back({e, what, slice});
    }
    ~WireTrace() {
        if (!on || !t0) return;
        for (auto &k : marks) {
            float ms = 0;
            cudaEventElapsedTime(&ms, t0, k.ev);
            fprintf(stderr, "[gcrf wire] slice %d %-10s done at %.3f ms\n", k.slice, k.what, ms);
            cudaEventDestroy(k.ev);
        }
        cudaEventDestroy(t0);
    }
};
}  // namespace

int gcrf_marginals_windowed_wire(gcrf_model *m, const gcrf_wire *w, int32_t window, int32_t step, int32_t pad, void *out,
                                 uint32_t flags) {
    if (!m) return fail(GCRF_EINVAL, "model handle is NULL");
    if (!w) return fail(GCRF_EINVAL, "wire batch is NULL");
    if (flags & ~(uint32_t)(GCRF_FLAG_OUT_F32 | GCRF_FLAG_F64)) return fail(GCRF_EINVAL, "only GCRF_FLAG_OUT_F32 / GCRF_FLAG_F64 apply to a wire batch");
    const int64_t C = gcrf_wire_contigs(w), G = gcrf_wire_genes(w), nnz = gcrf_wire_ids(w);
    if (G == 0) return GCRF_OK;
    if (!out) return fail(GCRF_EINVAL, "out is NULL");
    m->ev_open = false;
    DeviceGuard guard(m->device);
    const size_t osz = (flags & GCRF_FLAG_OUT_F32) ? 4 : 8;
    const int lw = gcrf::wire_len_width(w);
    const char *h = gcrf::wire_block(w);
    GCRF_CUDA(m->b_wire.reserve(gcrf::wire_total(w)));
    GCRF_CUDA(m->b_gene.reserve((size_t)(G + 1) * 4));
    GCRF_CUDA(m->b_attr.reserve((size_t)(nnz > 0 ? nnz : 1) * 4 + 64));
    GCRF_CUDA(m->b_out.reserve((size_t)G * osz));
    char *d = static_cast<char *>(m->b_wire.ptr);  // the device copy has the layout of the host block
    // The block was cut into contig-aligned slices when it was encoded, each one section of the block.  Three streams:
    // slice k+1 comes in (in_stream) while slice k is decoded and run through the kernels (stream) and slice k-1
    // travels back (copy_stream); every array lands where the whole batch would have put it.  The copy in is the long
    // pole (PCIe); the rest hides behind it except for the last — smallest — slice's kernels and copy back.
    // With the kernels being timed (gcrf_model_set_timing) the three stages run one after the other on `stream`, so
    // that the bracketed region holds kernels only.
    const int slices = gcrf::wire_slices(w);
    const bool pipelined = slices > 1 && !m->timing;
    cudaStream_t in = pipelined ? m->in_stream : m->stream, back = pipelined ? m->copy_stream : m->stream;
    WireTrace trace;
    trace.start(m->stream);
    if (pipelined) {  // work queued on the caller's stream before this call stays ahead of it
        GCRF_CUDA(cudaEventRecord(m->ev_slice[0], m->stream));
        GCRF_CUDA(cudaStreamWaitEvent(in, m->ev_slice[0], 0));
    }
    GCRF_CUDA(cudaMemcpyAsync(d, h, gcrf::wire_head_bytes(w), cudaMemcpyHostToDevice, in));  // contig_ptr + chunk sums

    auto copy_in = [&](int k) -> int {
        NvtxRange range("gcrf:stage wire");
        size_t off, size, rel_bytes, rel_stream;
        int64_t chunk0;
        gcrf::wire_section(w, k, &off, &size, &rel_bytes, &rel_stream, &chunk0);
        GCRF_CUDA(cudaMemcpyAsync(d + off, h + off, size, cudaMemcpyHostToDevice, in));
        trace.mark(in, "copy in", k);
        if (pipelined) GCRF_CUDA(cudaEventRecord(m->ev_in[k], in));
        return GCRF_OK;
    };
    bool timed_before = false;
    auto kernels = [&](int k) -> int {
        int64_t c0, c1, g0, g1, p0, p1, b0, b1;
        gcrf::wire_slice(w, k, &c0, &g0, &p0, &b0);
        gcrf::wire_slice(w, k + 1, &c1, &g1, &p1, &b1);
        size_t off, size, rel_bytes, rel_stream;
        int64_t chunk0;
        gcrf::wire_section(w, k, &off, &size, &rel_bytes, &rel_stream, &chunk0);
        if (pipelined) GCRF_CUDA(cudaStreamWaitEvent(m->stream, m->ev_in[k], 0));
        if (timed_before) m->ev_open = true;  // one bracket around the kernels of all slices: the first start, the last stop
        GCRF_CUDA(timing_begin(m));
        timed_before = m->timing;
        trace.mark(m->stream, "ready", k);
        cudaError_t derr = gcrf::launch_wire_decode(d + off, d + off + rel_bytes, lw, reinterpret_cast<const uint8_t *>(d + off + rel_stream),
                                                    g1 - g0, p0, gcrf::wire_rice_k(w),
                                                    reinterpret_cast<const int64_t *>(d + gcrf::wire_off_sums(w)) + 2 * chunk0,
                                                    static_cast<int32_t *>(m->b_gene.ptr) + g0, static_cast<int32_t *>(m->b_attr.ptr),
                                                    m->stream, &m->launches);
        if (derr != cudaSuccess) return fail_cuda(derr, "launch_wire_decode");
        trace.mark(m->stream, "decoded", k);
        // the decoded slice is a device-pointer batch of the regular entry point; its result lands in the library's buffer
        m->slice_gene_base = g0;
        m->slice_ids = p1 - p0;  // the slice's own density picks the tile size
        const int rc = gcrf_marginals_windowed(m, reinterpret_cast<const int32_t *>(d) + c0, static_cast<int32_t *>(m->b_gene.ptr) + g0,
                                               m->b_attr.ptr, c1 - c0, g1 - g0, nnz,
                                               window, step, pad, static_cast<char *>(m->b_out.ptr) + (size_t)g0 * osz,
                                               flags | GCRF_FLAG_DEVICE_PTRS | GCRF_FLAG_KEEP_TIMING | GCRF_FLAG_SLICE);
        if (rc != GCRF_OK) return rc;
        trace.mark(m->stream, "kernels", k);
        if (pipelined) GCRF_CUDA(cudaEventRecord(m->ev_slice[k], m->stream));
        return GCRF_OK;
    };
    auto copy_back = [&](int k) -> int {
        NvtxRange range("gcrf:finish");
        int64_t c0, c1, g0, g1, p0, p1, b0, b1;
        gcrf::wire_slice(w, k, &c0, &g0, &p0, &b0);
        gcrf::wire_slice(w, k + 1, &c1, &g1, &p1, &b1);
        if (pipelined) GCRF_CUDA(cudaStreamWaitEvent(back, m->ev_slice[k], 0));
        GCRF_CUDA(cudaMemcpyAsync(static_cast<char *>(out) + (size_t)g0 * osz, static_cast<char *>(m->b_out.ptr) + (size_t)g0 * osz,
                                  (size_t)(g1 - g0) * osz, cudaMemcpyDeviceToHost, back));
        trace.mark(back, "copy back", k);
        return GCRF_OK;
    };
    auto empty = [&](int k) {
        int64_t c0, c1, g0, g1, p0, p1, b0, b1;
        gcrf::wire_slice(w, k, &c0, &g0, &p0, &b0);
        gcrf::wire_slice(w, k + 1, &c1, &g1, &p1, &b1);
        return g1 <= g0;
    };
    int rc = GCRF_OK;
    if (pipelined) {
        for (int k = 0; k < slices; ++k) {
            if (empty(k)) continue;
            if ((rc = copy_in(k)) != GCRF_OK || (rc = kernels(k)) != GCRF_OK || (rc = copy_back(k)) != GCRF_OK) return rc;
        }
    } else {
        for (int k = 0; k < slices; ++k)
            if (!empty(k) && (rc = copy_in(k)) != GCRF_OK) return rc;
        for (int k = 0; k < slices; ++k)
            if (!empty(k) && (rc = kernels(k)) != GCRF_OK) return rc;
        for (int k = 0; k < slices; ++k)
            if (!empty(k) && (rc = copy_back(k)) != GCRF_OK) return rc;
    }
    if (pipelined) {
        GCRF_CUDA(cudaStreamSynchronize(m->in_stream));
        GCRF_CUDA(cudaStreamSynchronize(m->copy_stream));
    }
    GCRF_CUDA(cudaStreamSynchronize(m->stream));
    return GCRF_OK;
}

}  // extern "C"

extern "C" {

int gcrf_marginals_chain(gcrf_model *m, const int32_t *contig_ptr, const void *gene_ptr, const void *attr_idx,
                         int64_t C, int64_t G, int64_t nnz, void *out, uint32_t flags) {
    if (!m) return fail(GCRF_EINVAL, "model handle is NULL");
    m->ev_open = false;
    DeviceGuard guard(m->device);
    Batch b;
    int rc = stage_batch(m, contig_ptr, gene_ptr, attr_idx, C, G, nnz, out, flags, &b);
    if (rc != GCRF_OK || G == 0) return rc;
    GCRF_CUDA(m->b_scratch.reserve((size_t)G * 2 * sizeof(double)));
    gcrf::ChainArgs args{};
    args.model = m->dev;
    args.csr = b.csr;
    args.out = b.d_out;
    args.out_f32 = (flags & GCRF_FLAG_OUT_F32) ? 1 : 0;
    args.table64 = m->d_table64;
    args.m01 = m->m01;
    args.m10 = m->m10;
    args.m11 = m->m11;
    args.scratch = static_cast<double *>(m->b_scratch.ptr);
    NvtxRange range("gcrf:launch chain");
    GCRF_CUDA(timing_begin(m));
    cudaError_t err = gcrf::launch_chain(args, m->num_sms, m->stream, &m->launches);
    if (err != cudaSuccess) return fail_cuda(err, "launch_chain");
    GCRF_CUDA(timing_end(m));
    return finish_batch(m, b, out);
}

int gcrf_model_set_vocabulary(gcrf_model *m, const int32_t *accession_of_attr, int32_t A) {
    if (!m) return fail(GCRF_EINVAL, "model handle is NULL");
    if (A != m->A) return fail(GCRF_EINVAL, "vocabulary has %d entries, the model has %d attributes", A, m->A);
    if (A > 0 && !accession_of_attr) return fail(GCRF_EINVAL, "accession_of_attr is NULL");
    int32_t max_acc = -1;
    for (int32_t a = 0; a < A; ++a) {
        if (accession_of_attr[a] < 0) return fail(GCRF_EINVAL, "negative accession for attribute %d", a);
        if (accession_of_attr[a] > max_acc) max_acc = accession_of_attr[a];
    }
    if (max_acc > (1 << 26)) return fail(GCRF_EUNSUPPORTED, "accessions above 2^26 need a hashed vocabulary");
    std::vector<int32_t> lut((size_t)max_acc + 1, -1);
    for (int32_t a = 0; a < A; ++a) {
        if (lut[accession_of_attr[a]] != -1) return fail(GCRF_EINVAL, "accession %d appears twice", accession_of_attr[a]);
        lut[accession_of_attr[a]] = a;
    }
    DeviceGuard guard(m->device);
    if (m->d_lut) cudaFree(m->d_lut);
    m->d_lut = nullptr;
    m->lut_size = 0;
    if (lut.empty()) return GCRF_OK;
    GCRF_CUDA(cudaMalloc(reinterpret_cast<void **>(&m->d_lut), lut.size() * sizeof(int32_t)));
    GCRF_CUDA(cudaMemcpy(m->d_lut, lut.data(), lut.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    m->lut_size = (int32_t)lut.size();
    return GCRF_OK;
}

int gcrf_features_from_accessions(gcrf_model *m, const int32_t *accession, const void *gene_ptr, int64_t G, int64_t nnz,
                                  int32_t *attr_idx_out, uint32_t flags) {
    if (!m) return fail(GCRF_EINVAL, "model handle is NULL");
    if (!m->d_lut && m->A > 0) return fail(GCRF_EINVAL, "gcrf_model_set_vocabulary has not been called");
    if (G < 0 || nnz < 0) return fail(GCRF_EINVAL, "negative size");
    if (G == 0 || nnz == 0) return GCRF_OK;
    if (!accession || !gene_ptr || !attr_idx_out) return fail(GCRF_EINVAL, "NULL array");
    const bool ptr64 = (flags & GCRF_FLAG_PTR64) != 0;
    DeviceGuard guard(m->device);
    const int32_t *d_acc = accession;
    const void *d_ptr = gene_ptr;
    int32_t *d_out = attr_idx_out;
    const bool host = (flags & GCRF_FLAG_DEVICE_PTRS) == 0;
    if (host) {
        const int64_t last = ptr64 ? static_cast<const int64_t *>(gene_ptr)[G] : static_cast<const int32_t *>(gene_ptr)[G];
        if (last != nnz) return fail(GCRF_EINVAL, "gene_ptr must end at nnz");
        const size_t gene_bytes = (size_t)(G + 1) * (ptr64 ? 8 : 4);
        GCRF_CUDA(m->b_gene.reserve(gene_bytes));
        GCRF_CUDA(m->b_attr.reserve((size_t)nnz * 4 + 16));
        GCRF_CUDA(m->b_scratch.reserve((size_t)nnz * 4));
        GCRF_CUDA(cudaMemcpyAsync(m->b_gene.ptr, gene_ptr, gene_bytes, cudaMemcpyHostToDevice, m->stream));
        GCRF_CUDA(cudaMemcpyAsync(m->b_scratch.ptr, accession, (size_t)nnz * 4, cudaMemcpyHostToDevice, m->stream));
        d_acc = static_cast<const int32_t *>(m->b_scratch.ptr);
        d_ptr = m->b_gene.ptr;
        d_out = static_cast<int32_t *>(m->b_attr.ptr);
    }
    cudaError_t err = gcrf::launch_features(d_acc, ptr64 ? nullptr : static_cast<const int32_t *>(d_ptr),
                                            ptr64 ? static_cast<const int64_t *>(d_ptr) : nullptr, G, nnz, m->d_lut,
                                            m->lut_size, m->A, d_out, m->num_sms, m->stream, &m->launches);
    if (err != cudaSuccess) return fail_cuda(err, "launch_features");
    if (host) {
        GCRF_CUDA(cudaMemcpyAsync(attr_idx_out, d_out, (size_t)nnz * 4, cudaMemcpyDeviceToHost, m->stream));
        GCRF_CUDA(cudaStreamSynchronize(m->stream));
    }
    return GCRF_OK;
}

int gcrf_segments(gcrf_model *m, const int32_t *contig_ptr, const void *prob, const uint8_t *annotated, int64_t C, int64_t G,
                  double threshold, int32_t n_cds, int32_t edge_distance, int32_t trim, int32_t *seg_contig,
                  int32_t *seg_begin, int32_t *seg_end, int32_t *seg_ordinal, double *seg_avg_p, double *seg_max_p,
                  int64_t capacity, int64_t *n_segments, uint32_t flags) {
    if (!m) return fail(GCRF_EINVAL, "model handle is NULL");
    if (!n_segments) return fail(GCRF_EINVAL, "n_segments is NULL");
    *n_segments = 0;
    if (C < 0 || G < 0 || capacity < 0) return fail(GCRF_EINVAL, "negative size");
    if (G > 0x7fffffff - 1024) return fail(GCRF_EINVAL, "G must fit in int32 (shard the batch)");
    if ((C == 0) != (G == 0) || C > G) return fail(GCRF_EINVAL, "C and G are inconsistent");
    if (edge_distance < 0) return fail(GCRF_EINVAL, "edge_distance must be >= 0");
    if (threshold != threshold) return fail(GCRF_EINVAL, "threshold is NaN");
    if (G == 0) return GCRF_OK;
    if (!contig_ptr || !prob || !annotated) return fail(GCRF_EINVAL, "NULL array");
    if (capacity > 0 && (!seg_contig || !seg_begin || !seg_end || !seg_ordinal || !seg_avg_p || !seg_max_p))
        return fail(GCRF_EINVAL, "NULL output array");
    const bool device_ptrs = (flags & GCRF_FLAG_DEVICE_PTRS) != 0;
    const bool f32 = (flags & GCRF_FLAG_PROB_F32) != 0;
    DeviceGuard guard(m->device);

    gcrf::SegmentsArgs a{};
    a.C = C;
    a.G = G;
    a.prob_f32 = f32 ? 1 : 0;
    a.threshold = threshold;
    a.n_cds = n_cds;
    a.edge_distance = edge_distance;
    a.trim = trim ? 1 : 0;
    a.reset_per_contig = (flags & GCRF_FLAG_RESET_PER_CONTIG) ? 1 : 0;
    a.capacity = capacity;
    const size_t cap = (size_t)capacity;
    // outputs (host mode) and the count live in one block: [count | contig | begin | end | ordinal | avg | max]
    const size_t out_bytes = device_ptrs ? 16 : 16 + cap * (4 * sizeof(int32_t) + 2 * sizeof(double)) + 64;
    GCRF_CUDA(m->b_seg.reserve(out_bytes));
    char *ob = static_cast<char *>(m->b_seg.ptr);
    a.count = reinterpret_cast<int64_t *>(ob);
    if (device_ptrs) {
        a.contig_ptr = contig_ptr;
        a.prob = prob;
        a.annotated = annotated;
        a.seg_contig = seg_contig;
        a.seg_begin = seg_begin;
        a.seg_end = seg_end;
        a.seg_ordinal = seg_ordinal;
        a.seg_avg_p = seg_avg_p;
        a.seg_max_p = seg_max_p;
    } else {
        if (contig_ptr[0] != 0 || contig_ptr[C] != G) return fail(GCRF_EINVAL, "contig_ptr must start at 0 and end at G");
        for (int64_t c = 0; c < C; ++c)
            if (contig_ptr[c + 1] <= contig_ptr[c]) return fail(GCRF_EINVAL, "contig_ptr must be strictly increasing (contig %lld is empty)", (long long)c);
        const size_t pbytes = (size_t)G * (f32 ? 4 : 8);
        GCRF_CUDA(m->b_contig.reserve((size_t)(C + 1) * 4));
        GCRF_CUDA(m->b_out.reserve(pbytes));
        GCRF_CUDA(m->b_ann.reserve((size_t)G));
        GCRF_CUDA(cudaMemcpyAsync(m->b_contig.ptr, contig_ptr, (size_t)(C + 1) * 4, cudaMemcpyHostToDevice, m->stream));
        GCRF_CUDA(cudaMemcpyAsync(m->b_out.ptr, prob, pbytes, cudaMemcpyHostToDevice, m->stream));
        GCRF_CUDA(cudaMemcpyAsync(m->b_ann.ptr, annotated, (size_t)G, cudaMemcpyHostToDevice, m->stream));
        a.contig_ptr = static_cast<const int32_t *>(m->b_contig.ptr);
        a.prob = m->b_out.ptr;
        a.annotated = static_cast<const uint8_t *>(m->b_ann.ptr);
        char *q = ob + 16;
        a.seg_avg_p = reinterpret_cast<double *>(q); q += cap * sizeof(double);
        a.seg_max_p = reinterpret_cast<double *>(q); q += cap * sizeof(double);
        a.seg_contig = reinterpret_cast<int32_t *>(q); q += cap * sizeof(int32_t);
        a.seg_begin = reinterpret_cast<int32_t *>(q); q += cap * sizeof(int32_t);
        a.seg_end = reinterpret_cast<int32_t *>(q); q += cap * sizeof(int32_t);
        a.seg_ordinal = reinterpret_cast<int32_t *>(q);
    }
    GCRF_CUDA(m->b_scratch.reserve(gcrf::segments_scratch_bytes(G, C, m->num_sms)));
    NvtxRange range("gcrf:launch segments");
    GCRF_CUDA(timing_begin(m));
    cudaError_t err = gcrf::launch_segments(a, m->b_scratch.ptr, m->num_sms, m->stream, &m->launches);
    if (err != cudaSuccess) return fail_cuda(err, "launch_segments");
    GCRF_CUDA(timing_end(m));
    int64_t count = 0;
    GCRF_CUDA(cudaMemcpyAsync(&count, a.count, sizeof(count), cudaMemcpyDeviceToHost, m->stream));
    GCRF_CUDA(cudaStreamSynchronize(m->stream));
    *n_segments = count;
    if (!device_ptrs) {
        const size_t n = (size_t)(count < capacity ? count : capacity);
        if (n > 0) {
            GCRF_CUDA(cudaMemcpyAsync(seg_avg_p, a.seg_avg_p, n * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
            GCRF_CUDA(cudaMemcpyAsync(seg_max_p, a.seg_max_p, n * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
            GCRF_CUDA(cudaMemcpyAsync(seg_contig, a.seg_contig, n * sizeof(int32_t), cudaMemcpyDeviceToHost, m->stream));
            GCRF_CUDA(cudaMemcpyAsync(seg_begin, a.seg_begin, n * sizeof(int32_t), cudaMemcpyDeviceToHost, m->stream));
            GCRF_CUDA(cudaMemcpyAsync(seg_end, a.seg_end, n * sizeof(int32_t), cudaMemcpyDeviceToHost, m->stream));
            GCRF_CUDA(cudaMemcpyAsync(seg_ordinal, a.seg_ordinal, n * sizeof(int32_t), cudaMemcpyDeviceToHost, m->stream));
            GCRF_CUDA(cudaStreamSynchronize(m->stream));
        }
    }
    return GCRF_OK;
}

int gcrf_host_alloc(void **ptr, uint64_t bytes) {
    if (!ptr) return fail(GCRF_EINVAL, "ptr is NULL");
    *ptr = nullptr;
    GCRF_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return GCRF_OK;
}

int gcrf_host_free(void *ptr) {
    if (!ptr) return GCRF_OK;
    GCRF_CUDA(cudaFreeHost(ptr));
    return GCRF_OK;
}

int64_t gcrf_model_launch_count(const gcrf_model *m) { return m ? m->launches : 0; }

int32_t gcrf_max_window(const gcrf_model *m, int32_t f64) {
    if (!m) return 0;
    return f64 ? 0x7fffffff : gcrf::windowed_max_window(m->A);
}

int gcrf_model_set_timing(gcrf_model *m, int32_t enable) {
    if (!m) return fail(GCRF_EINVAL, "model handle is NULL");
    m->timing = enable != 0;
    if (!m->timing) m->timed = false;
    return GCRF_OK;
}

double gcrf_model_last_kernel_ms(gcrf_model *m) {
    if (!m || !m->timed) return -1.0;
    DeviceGuard guard(m->device);
    if (cudaEventSynchronize(m->ev_stop) != cudaSuccess) return -1.0;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, m->ev_start, m->ev_stop) != cudaSuccess) return -1.0;
    return (double)ms;
}

}  // extern "C"
