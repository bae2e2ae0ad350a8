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
