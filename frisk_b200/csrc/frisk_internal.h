// Shared by the translation units of libfrisk_b200.so; not part of the C ABI.
#ifndef FRISK_INTERNAL_H
#define FRISK_INTERNAL_H

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <functional>

struct frisk_b200_fasta;

namespace frisk_internal {

// records the CUDA error text for frisk_b200_last_cuda_error() and returns FRISK_E_CUDA
int cuda_fail(cudaError_t e, const char* what);

// cached per-device workspace (grown on demand, freed by frisk_b200_release_workspace)
int ws_get(int slot, size_t bytes, void** out);

// once per device: let the default memory pool keep up to 1 GiB of freed cudaMallocAsync scratch
int pool_ready();
extern int g_ingest_exact, g_ingest_chunk_tiles;   // frisk_ingest.cu; set through frisk_b200_set_option

// Streamed FASTA open that also builds the planes (frisk_ingest.cu; used by frisk_b200_run_fasta).  While the text is
// still arriving, on_range (optional) is handed the plane words [d_word_range[0], d_word_range[1]) -- a DEVICE-side range --
// that became final: the caller counts them.  If the streamed attempt has to be abandoned (see frisk_b200_fasta_open),
// abandon() undoes what on_range accumulated, the text is opened and packed in one piece, and counted stays false.
struct IngestSink {
    std::function<int(const uint32_t* codes, const uint32_t* inv, const uint32_t* low, const unsigned long long* d_word_range,
                      uint64_t words_hint, cudaStream_t st)> on_range;
    std::function<int(cudaStream_t st)> abandon;
    // called once the last chunk's passes (and its on_range) are queued: the record table and the planes are complete in
    // stream order, the host has not seen anything yet -- whatever only needs DEVICE-side results can be queued here
    std::function<int(frisk_b200_fasta* h, cudaStream_t st)> on_complete;
    cudaEvent_t uploaded_mark = nullptr;     // recorded behind the last text chunk's copy
    bool counted = false;                    // out: on_range saw every word of [0, padded_len / 32 - 1)
};
// the device side of a handle's record table, for on_complete: lengths, offsets, capacity, and the ingest's counters
// (records at [kIngestRecords], non-upper-case bases at [kIngestNonUpper], "capacities exceeded" at [kIngestOverflow])
enum { kIngestRecords = 0, kIngestNonUpper = 1, kIngestOverflow = 5 };
bool fasta_open_was_streamed(const frisk_b200_fasta* h);     // false: the handle comes from the exact re-open
int fasta_device_table(const frisk_b200_fasta* h, const unsigned long long** d_len, const unsigned long long** d_scaf_off,
                       const unsigned long long** d_counters, uint64_t* rec_cap, const uint32_t** d_codes, const uint32_t** d_inv,
                       const uint32_t** d_low);
int fasta_open_planes(const char* text, uint64_t n, cudaStream_t st, IngestSink* sink, frisk_b200_fasta** out);
// background count of a word range that lives in device memory (frisk_kernels.cu); kmax <= FRISK_B200_FAST_K
int background_device_range(const uint32_t* codes, const uint32_t* inv, const uint32_t* low, const unsigned long long* d_word_range,
                            uint64_t words_hint, int kmax, int mask_host, uint64_t* fwd, cudaStream_t st);

// frisk_b200_windows on the device (frisk_windows.cu): window list, window count and genome space from a record table in
// device memory; d_n_rec / d_bad / d_non_upper point into the ingest's counters.  d_win_off == nullptr: count / space only.
int windows_device(const unsigned long long* d_len, const unsigned long long* d_scaf_off, const unsigned long long* d_n_rec,
                   const unsigned long long* d_bad, const unsigned long long* d_non_upper, uint64_t rec_cap, int w, int step,
                   int scaffolds_all, uint64_t cap, unsigned long long* d_first, unsigned long long* d_win_off, uint32_t* d_win_len,
                   unsigned long long* d_n_win, long long* d_space, cudaStream_t st);

// FASTA header rule of the reference (F:156): name = line.strip().strip('>').split()[0].
// The line starts at `line_start`; returns false for an empty name (the reference raises IndexError).
bool parse_header_name(const unsigned char* t, uint64_t n, uint64_t line_start, uint64_t* name_off, uint32_t* name_len);

// general path (frisk_general.cu): run-time K up to FRISK_B200_MAX_K, windows of any length
int general_background(const uint32_t* codes, const uint32_t* inv, const uint32_t* low, uint64_t w_lo, uint64_t w_hi, int K,
                       int mask_host, uint64_t* fwd, cudaStream_t st);
int general_finalize(const uint64_t* fwd, int K, int symmetric, uint64_t* tables, uint64_t* valid, cudaStream_t st);
int general_genome_ivom(const uint64_t* tables, int kmin, int K, int64_t space, double* ig, cudaStream_t st);
int general_score(const uint32_t* codes, const uint32_t* inv, const uint32_t* low, const uint64_t* win_off,
                  const uint32_t* win_len, uint64_t n_win, uint32_t max_len, const double* ig, int kmin, int K, int want_rip,
                  double* rows, uint32_t* status, uint16_t* dump, cudaStream_t st, const uint32_t* redo_src = nullptr,
                  int grid_cap = 0);

// kmax 9..12, windows <= 8,186 bases (frisk_nibble_ext.cu): orders 1..8 as in the kmax-8 nibble kernel, orders 9..K from the
// observation that an 8-mer seen once has only unique extensions (the few repeated ones go through a shared-memory hash
// table).  What it cannot hold is marked kRowRedo and re-done by general_score behind it.
int score_nibble_ext(const uint32_t* codes, const uint32_t* inv, const uint32_t* low, const uint64_t* win_off,
                     const uint32_t* win_len, uint64_t n_win, uint32_t max_len, const double* ig, int kmin, int K, int want_rip,
                     double* rows, uint32_t* status, cudaStream_t st);


// direct score kernel (frisk_direct.cu): kmax 7, 8 and windows <= 8,186 bases.  Windows it cannot finish exactly
// (a K-mer seen 256+ times, more than 64 N-boundary words) are marked kRowRedo and re-done by the bucketed kernel
// (frisk_kernels.cu), launched behind it on the same stream.  The marks live in `status` when that is device memory,
// in stream-ordered device scratch when `status` is a pinned host buffer written over PCIe (frisk_b200_run_host):
// the second launch must not read its marks back across the bus one window at a time.
constexpr uint32_t kRowRedo = 0x80000000u;
int score_direct(const uint32_t* codes, const uint32_t* inv, const uint32_t* low, const uint64_t* win_off,
                 const uint32_t* win_len, uint64_t n_win, uint32_t max_len, const double* ig, int kmin, int K, int want_rip,
                 double* rows, uint32_t* status, uint16_t* dump, cudaStream_t st);
int score_direct_occupancy(int K, uint32_t max_len, int* ctas_per_sm, int* threads_per_cta);
int score_bucket_redo(const uint32_t* codes, const uint32_t* inv, const uint32_t* low, const uint64_t* win_off,
                      const uint32_t* win_len, uint64_t n_win, uint32_t max_len, const double* ig, int kmin, int K, int want_rip,
                      double* rows, uint32_t* status, uint16_t* dump, const uint32_t* redo_src, cudaStream_t st);
int sm_count_cached();

// nibble score kernel (frisk_nibble.cu): kmax 7, 8 and windows <= 8,186 bases; one 4-bit counter per K-mer, one atomic per
// position, the first occurrence of a K-mer scores it.  Windows with a K-mer seen 16+ times (or more than 64 N-boundary
// words) are marked kRowRedo and re-done by the bucketed kernel, exactly like the direct kernel's hand-over.
int score_nibble(const uint32_t* codes, const uint32_t* inv, const uint32_t* low, const uint64_t* win_off,
                 const uint32_t* win_len, uint64_t n_win, uint32_t max_len, const double* ig, int kmin, int K, int want_rip,
                 double* rows, uint32_t* status, uint16_t* dump, cudaStream_t st, const unsigned long long* n_win_dev = nullptr);
int score_nibble_occupancy(int K, uint32_t max_len, int* ctas_per_sm, int* threads_per_cta);
// the k sweep (BASELINE config C3): rows of kmax' = 1..8, kmin 1, from one pass over every window; ig / rows / status are
// HOST arrays of 8 device pointers
int score_sweep(const uint32_t* codes, const uint32_t* inv, const uint32_t* low, const uint64_t* win_off, const uint32_t* win_len,
                uint64_t n_win, uint32_t max_len, const double* const* ig, int want_rip, double* const* rows,
                uint32_t* const* status, cudaStream_t st);

}  // namespace frisk_internal

#define FRISK_CK(call)                                                        \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) return frisk_internal::cuda_fail(e_, #call);   \
    } while (0)

#endif
