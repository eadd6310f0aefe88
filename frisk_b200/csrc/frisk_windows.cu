// frisk_b200_windows (frisk_host.cpp: crawlGenome's enumeration, F:194-251) on the device, from a record table that lives in
// device memory.  frisk_b200_run_fasta uses it so that nothing between the last byte of the FASTA text and the window kernel
// waits for the host: record lengths and offsets are written by the ingest's pack pass, the window list, the number of
// windows and the genome space (totalLen - nnTotal, F:379) by the two kernels here, and the window kernel reads its window
// count from device memory.  Same windows, same order as the host function (tests/test_ingest_gpu.py compares them).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/frisk_b200.h"
#include "frisk_internal.h"

namespace {

constexpr int kWT = 1024;

// windows of one scaffold: F:211/F:222 size rule (in double, like the reference), then one window per j = 0, step, ... with
// j + step <= size (the xrange of F:228) -- regular ones while j + w <= size, the jump-back ones after that
__device__ __forceinline__ unsigned long long windows_of(unsigned long long size, int w, int step, int scaffolds_all, double min_size) {
    if ((double)size <= min_size) return (scaffolds_all && size > 0) ? 1ull : 0ull;
    return size >= (unsigned long long)step ? size / (unsigned long long)step : 0ull;
}

__global__ void __launch_bounds__(kWT)
windows_count_kernel(const unsigned long long* __restrict__ rec_len, const unsigned long long* __restrict__ n_rec_dev,
                     const unsigned long long* __restrict__ bad, const unsigned long long* __restrict__ non_upper, uint64_t rec_cap,
                     int w, int step, int scaffolds_all, unsigned long long* __restrict__ first, unsigned long long* __restrict__ n_win,
                     long long* __restrict__ space) {
    __shared__ unsigned long long sm_w[32], sm_b[32];
    __shared__ unsigned long long s_tot_w, s_tot_b;
    unsigned long long n = *n_rec_dev;
    if (n > rec_cap || (bad && *bad)) n = 0;                 // (the speculative capacities were exceeded: the caller starts over)
    const double min_size = (double)w + (((double)w * 0.75) - (double)step);
    const unsigned long long per = (n + kWT - 1) / kWT;
    const unsigned long long lo = min((unsigned long long)threadIdx.x * per, n), hi = min(lo + per, n);
    unsigned long long cw = 0, cb = 0;
    for (unsigned long long s = lo; s < hi; ++s) {
        const unsigned long long size = rec_len[s];
        cb += size;
        cw += windows_of(size, w, step, scaffolds_all, min_size);
    }
    // exclusive scan of the window counts over the threads (and the two totals)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long iw = cw, ib = cb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long a = __shfl_up_sync(0xffffffffu, iw, o), b = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { iw += a; ib += b; }
    }
    if (lane == 31) { sm_w[warp] = iw; sm_b[warp] = ib; }
    __syncthreads();
    if (warp == 0) {
        unsigned long long a = sm_w[lane], b = sm_b[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long x = __shfl_up_sync(0xffffffffu, a, o), y = __shfl_up_sync(0xffffffffu, b, o);
            if (lane >= o) { a += x; b += y; }
        }
        sm_w[lane] = a;                                       // inclusive over the warps
        if (lane == 31) { s_tot_w = a; s_tot_b = b; }
    }
    __syncthreads();
    unsigned long long run = (warp ? sm_w[warp - 1] : 0ull) + iw - cw;
    if (first) {
        for (unsigned long long s = lo; s < hi; ++s) {
            first[s] = run;
            run += windows_of(rec_len[s], w, step, scaffolds_all, min_size);
        }
        if (threadIdx.x == 0) first[n] = s_tot_w;
    }
    if (threadIdx.x == 0) {
        if (n_win) *n_win = s_tot_w;
        if (space) *space = (long long)s_tot_b - (long long)*non_upper;
    }
}

// one warp per scaffold, lanes over its windows
__global__ void __launch_bounds__(256)
windows_fill_kernel(const unsigned long long* __restrict__ rec_len, const unsigned long long* __restrict__ scaf_off,
                    const unsigned long long* __restrict__ n_rec_dev, const unsigned long long* __restrict__ bad, uint64_t rec_cap,
                    int w, int step, const unsigned long long* __restrict__ first, uint64_t cap,
                    unsigned long long* __restrict__ win_off, uint32_t* __restrict__ win_len) {
    unsigned long long n = *n_rec_dev;
    if (n > rec_cap || (bad && *bad)) return;
    const double min_size = (double)w + (((double)w * 0.75) - (double)step);
    const int lane = threadIdx.x & 31;
    const unsigned long long warps = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    for (unsigned long long s = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < n; s += warps) {
        const unsigned long long f = first[s], cnt = first[s + 1] - f;
        if (!cnt) continue;
        const unsigned long long size = rec_len[s], base = scaf_off[s];
        if ((double)size <= min_size) {                       // --scaffoldsAll rescue (F:211-221): the whole scaffold
            if (lane == 0 && f < cap) { win_off[f] = base; win_len[f] = (uint32_t)size; }
            continue;
        }
        const unsigned long long uw = (unsigned long long)w, us = (unsigned long long)step;
        const unsigned long long n_reg = size >= uw ? min(size - uw, size - us) / us + 1ull : 0ull;
        for (unsigned long long k = lane; k < cnt; k += 32) {
            const unsigned long long idx = f + k;
            if (idx >= cap) break;
            unsigned long long o, l;
            if (k < n_reg) { o = k * us; l = uw; }
            else if (size >= uw) { o = size - uw; l = uw; }   // F:230-232, F:243: the last w bases again
            else {
                // size < w (possible when 0.75 w < step): seq[size - w : size] has a negative start, which Python counts
                // from the end -- the slice is the last min(w - size, size) bases
                const unsigned long long d = uw - size;
                if (d <= size) { o = size - d; l = d; } else { o = 0; l = size; }
            }
            win_off[idx] = base + o;
            win_len[idx] = (uint32_t)l;
        }
    }
}

}  // namespace

int frisk_internal::windows_device(const unsigned long long* d_len, const unsigned long long* d_scaf_off,
                                   const unsigned long long* d_n_rec, const unsigned long long* d_bad,
                                   const unsigned long long* d_non_upper, uint64_t rec_cap, int w, int step, int scaffolds_all,
                                   uint64_t cap, unsigned long long* d_first, unsigned long long* d_win_off, uint32_t* d_win_len,
                                   unsigned long long* d_n_win, long long* d_space, cudaStream_t st) {
    if (!d_len || !d_n_rec || w < 1 || step < 1 || (uint64_t)w > FRISK_B200_MAX_WINDOW) return FRISK_E_INVALID;
    if (d_space && !d_non_upper) return FRISK_E_INVALID;
    const bool fill = d_win_off && d_win_len && d_first && d_scaf_off;
    windows_count_kernel<<<1, kWT, 0, st>>>(d_len, d_n_rec, d_bad, d_non_upper, rec_cap, w, step, scaffolds_all, fill ? d_first : nullptr,
                                            d_n_win, d_space);
    if (fill) {
        const int sms = frisk_internal::sm_count_cached();
        windows_fill_kernel<<<(unsigned)(sms > 0 ? sms * 8 : 1024), 256, 0, st>>>(d_len, d_scaf_off, d_n_rec, d_bad, rec_cap, w, step, d_first, cap,
                                                                             d_win_off, d_win_len);
    }
    FRISK_CK(cudaGetLastError());
    return FRISK_OK;
}

extern "C" {

// (tests) the window list of frisk_b200_windows from a record table in device memory
int frisk_b200_windows_device(const uint64_t* d_scaf_len, const uint64_t* d_scaf_off, uint64_t n_scaf, int w, int step, int scaffolds_all,
                              uint64_t cap, uint64_t* d_win_off, uint32_t* d_win_len, uint64_t* d_n_windows, int64_t* d_genome_space,
                              void* stream) {
    if (!d_scaf_len || !d_scaf_off || !d_win_off || !d_win_len || !d_n_windows) return FRISK_E_INVALID;
    if (frisk_b200_device_count() <= 0) return FRISK_E_NO_DEVICE;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = frisk_internal::pool_ready();
    if (rc) return rc;
    unsigned long long* scratch = nullptr;                   // [0] n_scaf, [1] zero (nnTotal), [2..] first window of every scaffold
    FRISK_CK(cudaMallocAsync((void**)&scratch, (n_scaf + 4) * 8, st));
    const unsigned long long head[2] = {n_scaf, 0ull};
    FRISK_CK(cudaMemcpyAsync(scratch, head, 16, cudaMemcpyHostToDevice, st));
    rc = frisk_internal::windows_device((const unsigned long long*)d_scaf_len, (const unsigned long long*)d_scaf_off, scratch, nullptr,
                                        scratch + 1, n_scaf, w, step, scaffolds_all, cap, scratch + 2, (unsigned long long*)d_win_off,
                                        d_win_len, (unsigned long long*)d_n_windows, (long long*)d_genome_space, st);
    const cudaError_t e = cudaFreeAsync(scratch, st);
    if (!rc && e != cudaSuccess) return frisk_internal::cuda_fail(e, "cudaFreeAsync");
    return rc;
}

}  // extern "C"
