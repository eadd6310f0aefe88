// frisk_b200: strand-symmetric k-mer composition vectors of sequence regions, in batch.
//
// Second caller of the counting machinery: for every anomalous window / merged region the
// reference calls computeKmers(pcaMode=True, sym=True) (F:1576-1578: every word and its reverse
// complement counted, orders pcaMin..pcaMax, default 1..6), drops one of each {k-mer, reverse
// complement} pair (scrubMirrors F:797-811: the first in table order is kept) and turns each order into
// proportions (flattenKmerMap(prop=True) F:813-831).  Here: one CTA per region, tables in shared memory,
// the same "longest valid word + marginalise" counting as the background kernel, one row of doubles out.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/frisk_b200.h"
#include "frisk_internal.h"

namespace {

constexpr int kFT = 256;
constexpr uint32_t kFullMask = 0xffffffffu;

__host__ __device__ inline uint32_t fp4(int k) { return 1u << (2 * k); }
__host__ __device__ inline uint32_t foff(int x) { return (fp4(x) - 4u) / 3u; }

__device__ __forceinline__ uint32_t frevcomp(uint32_t idx, int x) {
    uint32_t r = __brev(idx) >> (32 - 2 * x);
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    return r ^ (0x55555555u & (fp4(x) - 1u));
}

__global__ void __launch_bounds__(kFT)
region_features_kernel(const uint32_t* __restrict__ codes, const uint32_t* __restrict__ inv,
                       const unsigned long long* __restrict__ reg_off, const uint32_t* __restrict__ reg_len, uint32_t n_reg,
                       int kmin, int kmax, const int32_t* __restrict__ slot, uint32_t n_feat, double* __restrict__ out) {
    extern __shared__ __align__(16) uint32_t tab[];          // orders 1..kmax, forward counts
    __shared__ unsigned long long red[kFT / 32];
    __shared__ unsigned long long total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_tab = foff(kmax + 1);
    for (uint32_t r = blockIdx.x; r < n_reg; r += gridDim.x) {
        const uint64_t o = reg_off[r];
        const uint32_t len = reg_len[r];
        for (uint32_t i = tid; i < n_tab; i += kFT) tab[i] = 0;
        __syncthreads();
        // every position adds 1 to the order of its longest valid word (upper-cased: soft-masked bases count, F:334-335)
        for (uint32_t p = tid; p < len; p += kFT) {
            const uint64_t a = o + p, wi = a >> 4, mi = a >> 5;
            const uint32_t c32 = __funnelshift_l(__ldg(codes + wi + 1), __ldg(codes + wi), (uint32_t)(a & 15) * 2u);
            const uint32_t m = __funnelshift_l(__ldg(inv + mi + 1), __ldg(inv + mi), (uint32_t)(a & 31));
            const uint32_t rest = len - p;
            const int v = min(min(__clz(m), kmax), (int)(rest < 8u ? rest : 8u));
            if (v > 0) atomicAdd(&tab[foff(v) + (c32 >> (32 - 2 * v))], 1u);
        }
        __syncthreads();
        for (int x = kmax - 1; x >= 1; --x) {                  // F_x = short words of order x + marginal of F_{x+1}
            for (uint32_t t = tid; t < fp4(x); t += kFT) {
                const uint32_t* ch = tab + foff(x + 1) + 4 * t;
                tab[foff(x) + t] += ch[0] + ch[1] + ch[2] + ch[3];
            }
            __syncthreads();
        }
        double* row = out + (size_t)r * n_feat;
        for (int k = kmin; k <= kmax; ++k) {
            // sum over the kept k-mers of (count + count of the reverse complement)
            unsigned long long s = 0;
            for (uint32_t t = tid; t < fp4(k); t += kFT)
                if (slot[foff(k) + t] >= 0) s += (unsigned long long)tab[foff(k) + t] + tab[foff(k) + frevcomp(t, k)];
            for (int ofs = 16; ofs; ofs >>= 1) s += __shfl_xor_sync(kFullMask, s, ofs);
            if (lane == 0) red[warp] = s;
            __syncthreads();
            if (tid == 0) {
                unsigned long long a = 0;
                for (int w = 0; w < kFT / 32; ++w) a += red[w];
                total = a;
            }
            __syncthreads();
            const double tk = (double)total;
            for (uint32_t t = tid; t < fp4(k); t += kFT) {
                const int32_t sl = slot[foff(k) + t];
                if (sl >= 0) {
                    const double c = (double)((unsigned long long)tab[foff(k) + t] + tab[foff(k) + frevcomp(t, k)]);
                    row[sl] = total ? c / tk : CUDART_NAN;     // the reference divides by zero here (F:824)
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace

extern "C" {

int frisk_b200_feature_slots(int kmin, int kmax, int32_t* slot, uint64_t* n_features) {
    if (kmin < 1 || kmin > kmax || kmax > 7 || !n_features) return kmax > 7 ? FRISK_E_UNSUPPORTED : FRISK_E_INVALID;
    uint64_t n = 0;
    for (int k = 1; k <= kmax; ++k) {
        for (uint32_t t = 0; t < fp4(k); ++t) {
            uint32_t rc = 0, x = t;
            for (int i = 0; i < k; ++i) { rc = (rc << 2) | ((x & 3u) ^ 1u); x >>= 2; }
            const bool keep = k >= kmin && t <= rc;            // scrubMirrors keeps the first of each pair in table order
            if (slot) slot[foff(k) + t] = keep ? (int32_t)n : -1;
            if (keep) ++n;
        }
    }
    *n_features = n;
    return FRISK_OK;
}

int frisk_b200_region_features(const uint32_t* d_codes, const uint32_t* d_inv, const uint64_t* d_reg_off,
                               const uint32_t* d_reg_len, uint64_t n_regions, int kmin, int kmax, const int32_t* d_slot,
                               uint64_t n_features, double* d_out, void* stream) {
    if (n_regions == 0) return FRISK_OK;
    if (!d_codes || !d_inv || !d_reg_off || !d_reg_len || !d_slot || !d_out || kmin < 1 || kmin > kmax) return FRISK_E_INVALID;
    if (kmax > 7 || n_regions > 0xffffffffull) return FRISK_E_UNSUPPORTED;
    int dev = 0, sms = 0;
    FRISK_CK(cudaGetDevice(&dev));
    FRISK_CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (sms <= 0) return FRISK_E_NO_DEVICE;
    const size_t smem = (size_t)foff(kmax + 1) * 4;
    FRISK_CK(cudaFuncSetAttribute(region_features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint64_t grid = (uint64_t)sms * 4;
    if (grid > n_regions) grid = n_regions;
    region_features_kernel<<<(unsigned)grid, kFT, smem, (cudaStream_t)stream>>>(
        d_codes, d_inv, reinterpret_cast<const unsigned long long*>(d_reg_off), d_reg_len, (uint32_t)n_regions, kmin, kmax,
        d_slot, (uint32_t)n_features, d_out);
    FRISK_CK(cudaGetLastError());
    return FRISK_OK;
}

}  // extern "C"
