// gcrf_features.cu — feature extraction on device: integer domain accessions -> attribute ids of the CRF model,
// with the set semantics of gecco/crf/features.py:13-35 (a gene's features are a dict keyed by domain name, so a
// domain that occurs twice in a gene counts once) and the tagger's handling of unknown attributes (dropped).
// Row pointers are left untouched: dropped entries become -1, which the marginal kernels ignore.
#include "gcrf_kernels.cuh"

namespace gcrf {

namespace {

constexpr int kThreads = 256;

template <typename PtrT>
__global__ void __launch_bounds__(kThreads)
features_kernel(const int32_t *__restrict__ accession, const PtrT *__restrict__ gene_ptr, int64_t G,
                const int32_t *__restrict__ lut, int32_t lut_size, int32_t *__restrict__ out) {
    // one warp per gene: lane i resolves row i (+32, +64, ...) and looks for an earlier equal accession
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t g = warp0; g < G; g += nwarps) {
        const int64_t rb = (int64_t)__ldg(gene_ptr + g), re = (int64_t)__ldg(gene_ptr + g + 1);
        for (int64_t p = rb + lane; p < re; p += 32) {
            const int32_t acc = __ldg(accession + p);
            int32_t id = (acc >= 0 && acc < lut_size) ? __ldg(lut + acc) : -1;
            if (id >= 0) {
                for (int64_t q = rb; q < p; ++q) {
                    if (__ldg(accession + q) == acc) {
                        id = -1;  // features.py:32 — the first occurrence keeps the key
                        break;
                    }
                }
            }
            out[p] = id;
        }
    }
}

}  // namespace

cudaError_t launch_features(const int32_t *accession, const int32_t *gene_ptr32, const int64_t *gene_ptr64, int64_t G,
                            const int32_t *lut, int32_t lut_size, int32_t *attr_idx_out, int num_sms,
                            cudaStream_t stream, int64_t *launches) {
    if (G <= 0) return cudaSuccess;
    int64_t blocks = (G * 32 + kThreads - 1) / kThreads;
    if (blocks > (int64_t)num_sms * 16) blocks = (int64_t)num_sms * 16;
    if (gene_ptr64)
        features_kernel<int64_t><<<(int)blocks, kThreads, 0, stream>>>(accession, gene_ptr64, G, lut, lut_size, attr_idx_out);
    else
        features_kernel<int32_t><<<(int)blocks, kThreads, 0, stream>>>(accession, gene_ptr32, G, lut, lut_size, attr_idx_out);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace gcrf
