// frisk_b200 device-side FASTA ingest: FASTA text in, packed 2-bit planes out, on the GPU.
//
// Replaces the reference's iterFasta (F:139-164: line.strip(), blank lines skipped, '>' lines start
// a record, everything else is sequence) and countN (F:106-118) -- which the reference runs three
// times per file (F:170, F:203, F:297) in pure Python -- and this library's own host packer
// (frisk_b200_fasta_scan + frisk_b200_pack), whose outputs it reproduces bit for bit.  The host
// only copies the raw text to the device and reads back one small record table.
//
// The text is cut into tiles of 4096 bytes (256 threads x 16 bytes).  What a byte means depends on
// the line it is in (header or sequence) and on the record it belongs to, i.e. on everything before
// it, so the work is three tile passes separated by two scans over per-tile summaries:
//   fasta_lines_kernel     per tile: number of header lines starting in it, and whether the last
//                          line starting in it is a header
//   fasta_scan1_kernel     over tiles: records before the tile, header state carried into the tile
//   fasta_tile_kernel<0>   per tile: classify every byte; per record: header position and length
//                          (one atomic per record piece, not per base); per tile: bases before the
//                          first / after the last header; countN statistics
//   fasta_scan2_kernel     over tiles (segmented): bases of the open record before the tile
//   [host: names from the header positions, 128-base aligned layout -> scaf_off]
//   fasta_tile_kernel<1>   per tile: every thread ORs the <= 16 bases of its 16 bytes into the planes
//   fasta_padding_kernel   per record: the padding up to the next 128-base boundary is flagged invalid
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/frisk_b200.h"
#include "frisk_internal.h"

namespace frisk_internal {
int g_ingest_exact = 0;          // tests: always take the one-copy, count-first open
int g_ingest_chunk_tiles = 0;    // tests: tiles per upload chunk at least this (0 = default), so small texts are chunked too
}  // namespace frisk_internal

namespace {

constexpr int kTT = 256;                    // threads per tile
constexpr uint32_t kTile = kTT * 16u;       // text bytes per tile
constexpr int kST = 1024;                   // threads of the (single-CTA) tile scans
constexpr uint32_t kFullMask = 0xffffffffu;
// device counters of one open(): records, countN statistics (2), interior whitespace seen, and the state of a chunked open
enum { kCtrRecords = 0, kCtrNonUpper = 1, kCtrLower = 2, kCtrWhitespace = 3, kCtrAmbig = 4, kCtrOverflow = 5, kCtrKeyCarry = 6,
       kCtrBaseCarry = 7,
       // streamed planes (fasta_layout_kernel): records that have their offset, plane words handed to the count hook,
       // records whose padding is flagged, padded length of the finished layout
       kLayAssigned = 8, kLayCounted = 9, kLayPadded = 10, kLayPaddedLen = 11, kCtrCount = 12,
       kRangeSlots = 4 };                 // per chunk, behind the counters: {word_lo, word_hi, rec_lo, rec_hi}

// class of a byte: 0..3 = A,T,G,C (F:70 order); 4..7 = a,t,g,c; 8 = anything else; 9 = whitespace
// removed by the reference's line.strip() (F:149)
__device__ __forceinline__ uint32_t class_of(uint32_t c) {
    switch (c) {
        case 'A': return 0; case 'T': return 1; case 'G': return 2; case 'C': return 3;
        case 'a': return 4; case 't': return 5; case 'g': return 6; case 'c': return 7;
        case ' ': case '\t': case '\n': case '\r': case '\v': case '\f': return 9;
        default: return 8;
    }
}

__device__ __forceinline__ uint32_t byte_of(const uint4& v, int k) {
    const uint32_t w = k < 4 ? v.x : (k < 8 ? v.y : (k < 12 ? v.z : v.w));
    return (w >> (8 * (k & 3))) & 0xffu;
}

// Is the line starting at byte i a header?  Its first non-blank character decides (the reference
// strips the line before looking at it, F:149-153).
// `avail` (<= n): bytes already on the device (chunked upload).  A decision that would need a byte beyond it is reported
// through *ambig (the caller then falls back to the one-shot path); with avail == n it cannot happen.
__device__ bool line_is_header(const uint8_t* __restrict__ t, uint64_t i, uint64_t n, uint64_t avail, uint32_t* ambig) {
    while (i < n) {
        if (i >= avail) { *ambig = 1u; return false; }
        const uint32_t c = t[i];
        if (c == '\n') return false;                 // blank line
        if (class_of(c) != 9u) return c == '>';
        ++i;
    }
    return false;
}

// bit k of start_mask: a line starts at byte i0 + k; of hdr_mask: ... and it is a header line
__device__ __forceinline__ void find_starts(const uint8_t* __restrict__ t, uint64_t n, uint64_t i0, const uint4& raw,
                                            uint32_t& start_mask, uint32_t& hdr_mask, uint64_t avail, uint32_t* ambig) {
    start_mask = 0; hdr_mask = 0;
    uint32_t prev = i0 ? (uint32_t)t[i0 - 1] : (uint32_t)'\n';
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (prev == '\n') {
            start_mask |= 1u << k;
            if (line_is_header(t, i0 + k, n, avail, ambig)) hdr_mask |= 1u << k;
        }
        prev = byte_of(raw, k);
    }
}

// state of the last line start of a thread / tile: 0 = no line starts here, 1 = sequence line, 2 = header
__device__ __forceinline__ uint32_t last_key(uint32_t start_mask, uint32_t hdr_mask) {
    return start_mask ? 1u + ((hdr_mask >> (31 - __clz(start_mask))) & 1u) : 0u;
}

// ---- block-wide scans (blockDim.x a multiple of 32, <= 1024); sm: >= 64 words of shared memory ----------
template <typename T>
__device__ __forceinline__ T block_excl_sum(T v, T* sm, T* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    T incl = v;
#pragma unroll
    for (int ofs = 1; ofs < 32; ofs <<= 1) {
        const T y = __shfl_up_sync(kFullMask, incl, ofs);
        if (lane >= ofs) incl += y;
    }
    if (lane == 31) sm[warp] = incl;
    __syncthreads();
    T before = 0, tot = 0;
    for (int w = 0; w < nw; ++w) {
        const T x = sm[w];
        if (w < warp) before += x;
        tot += x;
    }
    __syncthreads();
    *total = tot;
    return before + incl - v;
}

// "last writer wins": the last non-zero key among the threads before this one (0 if none);
// *last = the last non-zero key of the whole block
__device__ __forceinline__ uint32_t block_excl_last(uint32_t key, uint32_t* sm, uint32_t* last) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t ball = __ballot_sync(kFullMask, key != 0u);
    const uint32_t prior = ball & ((1u << lane) - 1u);
    uint32_t v = __shfl_sync(kFullMask, key, prior ? 31 - __clz(prior) : 0);
    if (!prior) v = 0;
    const uint32_t wl = __shfl_sync(kFullMask, key, ball ? 31 - __clz(ball) : 0);
    if (lane == 0) sm[warp] = ball ? wl : 0u;
    __syncthreads();
    uint32_t carry = 0, tot = 0;
    for (int w = 0; w < nw; ++w) {
        const uint32_t x = sm[w];
        if (x) { tot = x; if (w < warp) carry = x; }
    }
    __syncthreads();
    *last = tot;
    return v ? v : carry;
}

// Segmented sum.  Element = (f, v): f = "a reset happens inside this element", v = the amount after
// its last reset (the whole amount when !f).  Returns the amount accumulated since the last reset
// before this thread; *reset_before = some earlier thread has a reset.
template <typename T>
__device__ __forceinline__ T block_excl_seg(bool f, T v, T* smv, uint32_t* smf, bool* reset_before) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T iv = v;
    uint32_t fl = f ? 1u : 0u;
#pragma unroll
    for (int ofs = 1; ofs < 32; ofs <<= 1) {
        const T uv = __shfl_up_sync(kFullMask, iv, ofs);
        const uint32_t uf = __shfl_up_sync(kFullMask, fl, ofs);
        if (lane >= ofs) { if (!fl) iv += uv; fl |= uf; }
    }
    T ev = __shfl_up_sync(kFullMask, iv, 1);
    uint32_t ef = __shfl_up_sync(kFullMask, fl, 1);
    if (lane == 0) { ev = 0; ef = 0; }
    if (lane == 31) { smv[warp] = iv; smf[warp] = fl; }
    __syncthreads();
    T c = 0;
    uint32_t cf = 0;
    for (int w = 0; w < warp; ++w) {
        if (smf[w]) { c = smv[w]; cf = 1; } else c += smv[w];
    }
    __syncthreads();
    if (!ef) ev += c;
    *reset_before = (ef | cf) != 0u;
    return ev;
}

// ---- pass 1: line starts ------------------------------------------------------------------------
__global__ void __launch_bounds__(kTT)
fasta_lines_kernel(const uint8_t* __restrict__ t, uint64_t n, uint32_t* __restrict__ tile_nhdr, uint8_t* __restrict__ tile_key,
                   uint64_t tile0, uint64_t avail, uint32_t* __restrict__ ambig) {
    __shared__ uint32_t sm[64];
    const uint64_t tile = tile0 + blockIdx.x;
    const uint64_t i0 = tile * kTile + threadIdx.x * 16u;
    const uint4 raw = *reinterpret_cast<const uint4*>(t + i0);
    uint32_t sm_, hm_;
    find_starts(t, n, i0, raw, sm_, hm_, avail, ambig);
    uint32_t total, last;
    block_excl_sum<uint32_t>(__popc(hm_), sm, &total);
    block_excl_last(last_key(sm_, hm_), sm, &last);
    if (threadIdx.x == 0) { tile_nhdr[tile] = total; tile_key[tile] = (uint8_t)last; }
}

// ---- scan 1 over tiles: records before each tile, header state carried into it -----------------
__global__ void __launch_bounds__(kST)
fasta_scan1_kernel(const uint32_t* __restrict__ tile_nhdr, const uint8_t* __restrict__ tile_key, uint64_t tile_lo, uint64_t tile_hi,
                   uint32_t* __restrict__ tile_rec_base, uint8_t* __restrict__ tile_carry_hdr,
                   unsigned long long* __restrict__ counters, uint32_t* __restrict__ key_carry) {
    // tiles [tile_lo, tile_hi): one chunk of a chunked upload (or everything); counters[0] = records before tile_lo on entry,
    // records before tile_hi on exit; *key_carry = state of the last line start before the range (0 none, 1 sequence, 2 header)
    __shared__ uint32_t sm[64];
    const uint64_t n_tiles = tile_hi - tile_lo;
    const uint64_t per = (n_tiles + kST - 1) / kST;
    const uint64_t lo = tile_lo + min((uint64_t)threadIdx.x * per, n_tiles), hi = min(lo + per, tile_hi);
    const uint32_t rec_in = (uint32_t)counters[0], key_in = *key_carry;
    __syncthreads();
    uint32_t sum = 0, key = 0;
    for (uint64_t i = lo; i < hi; ++i) {
        sum += tile_nhdr[i];
        const uint32_t k = tile_key[i];
        if (k) key = k;
    }
    uint32_t total, last;
    uint32_t run = rec_in + block_excl_sum<uint32_t>(sum, sm, &total);
    uint32_t rk = block_excl_last(key, sm, &last);
    if (!rk) rk = key_in;
    for (uint64_t i = lo; i < hi; ++i) {
        tile_rec_base[i] = run;
        tile_carry_hdr[i] = (uint8_t)(rk == 2u);
        run += tile_nhdr[i];
        const uint32_t k = tile_key[i];
        if (k) rk = k;
    }
    if (threadIdx.x == 0) { counters[0] = (unsigned long long)rec_in + total; if (last) *key_carry = last; }
}

// ---- scan 2 over tiles (segmented by headers): bases of the open record before each tile ------------
__global__ void __launch_bounds__(kST)
fasta_scan2_kernel(const uint32_t* __restrict__ tile_nhdr, const uint32_t* __restrict__ tile_pre,
                   const uint32_t* __restrict__ tile_post, uint64_t tile_lo, uint64_t tile_hi,
                   unsigned long long* __restrict__ tile_base_in, unsigned long long* __restrict__ base_carry) {
    // tiles [tile_lo, tile_hi); *base_carry = bases of the open record before tile_lo on entry, before tile_hi on exit
    __shared__ unsigned long long smv[32];
    __shared__ uint32_t smf[32];
    const uint64_t n_tiles = tile_hi - tile_lo;
    const uint64_t per = (n_tiles + kST - 1) / kST;
    const uint64_t lo = tile_lo + min((uint64_t)threadIdx.x * per, n_tiles), hi = min(lo + per, tile_hi);
    const unsigned long long carry_in = *base_carry;
    __syncthreads();
    bool f = false;
    unsigned long long v = 0;
    for (uint64_t i = lo; i < hi; ++i) {
        if (tile_nhdr[i]) { f = true; v = tile_post[i]; } else v += tile_pre[i];
    }
    bool rb;
    unsigned long long run = block_excl_seg<unsigned long long>(f, v, smv, smf, &rb);
    if (!rb) run += carry_in;                                   // no header yet in this range: the open record continues
    for (uint64_t i = lo; i < hi; ++i) {
        tile_base_in[i] = run;
        if (tile_nhdr[i]) run = tile_post[i]; else run += tile_pre[i];
    }
    if (threadIdx.x == kST - 1) *base_carry = run;              // (an empty last range leaves `run` = the prefix of everything)
}

// ---- passes 2 and 3: classify every byte of a tile -------------------------------------------------
// MODE 0: record table (header position, length), per-tile base counts, countN statistics.
// MODE 1: write every base to the planes at scaf_off[record] + index in record.
template <int MODE>
__global__ void __launch_bounds__(kTT)
fasta_tile_kernel(const uint8_t* __restrict__ t, uint64_t n, const uint32_t* __restrict__ tile_rec_base,
                  const uint8_t* __restrict__ tile_carry_hdr,
                  unsigned long long* __restrict__ rec_hdr_pos, unsigned long long* __restrict__ rec_len,
                  uint32_t* __restrict__ tile_pre, uint32_t* __restrict__ tile_post, unsigned long long* __restrict__ counters,
                  const unsigned long long* __restrict__ tile_base_in, const unsigned long long* __restrict__ scaf_off,
                  uint32_t* __restrict__ codes, uint32_t* __restrict__ inv, uint32_t* __restrict__ low,
                  uint64_t tile0, uint64_t avail, uint64_t rec_cap) {
    // tiles tile0 .. of a chunked upload whose first `avail` bytes are on the device (MODE 0; avail == n otherwise);
    // rec_cap: entries of rec_hdr_pos / rec_len (MODE 0) -- a record beyond it raises counters[kCtrOverflow]
    __shared__ uint32_t sm[64];
    __shared__ uint32_t smf[32];
    __shared__ uint8_t cls_tab[256];
    const int tid = threadIdx.x;
    if (MODE == 1 && counters[kCtrOverflow]) return;                   // streamed planes: no layout to write to (whole grid)
    cls_tab[tid] = (uint8_t)class_of((uint32_t)tid);
    const uint64_t tile = tile0 + blockIdx.x;
    const uint64_t i0 = tile * kTile + (uint64_t)tid * 16u;
    const uint4 raw = *reinterpret_cast<const uint4*>(t + i0);
    uint32_t start_mask, hdr_mask;
    uint32_t* const ambig = reinterpret_cast<uint32_t*>(counters + kCtrAmbig);
    find_starts(t, n, i0, raw, start_mask, hdr_mask, avail, ambig);
    const uint32_t n_hdr_t = __popc(hdr_mask);
    uint32_t tile_hdrs, last;
    const uint32_t hdr_before = block_excl_sum<uint32_t>(n_hdr_t, sm, &tile_hdrs);     // (also orders cls_tab)
    const uint32_t key_before = block_excl_last(last_key(start_mask, hdr_mask), sm, &last);
    bool in_hdr = key_before ? key_before == 2u : tile_carry_hdr[tile] != 0;
    const uint32_t rec_start = tile_rec_base[tile] + hdr_before;       // headers before this thread; open record = rec_start - 1
    const bool fits = MODE != 0 || (uint64_t)rec_start + n_hdr_t <= rec_cap;   // every record this thread writes to exists
    if (MODE == 0 && !fits) counters[kCtrOverflow] = 1ull;

    // walk the 16 bytes: which are bases, how many before the first / after the last header
    uint32_t base_mask = 0, cnt = 0, pre = 0, non_upper = 0, lower = 0;
    uint32_t cls16[2] = {0, 0};                                        // 4 bits per byte
    bool seen_hdr = false;
    {
        uint32_t rec = rec_start;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if ((start_mask >> k) & 1u) {
                const bool h = (hdr_mask >> k) & 1u;
                if (h) {
                    if (!seen_hdr) { pre = cnt; seen_hdr = true; }
                    else if (MODE == 0 && fits && cnt && rec >= 1u) atomicAdd(&rec_len[rec - 1u], (unsigned long long)cnt);
                    cnt = 0;
                    ++rec;
                    if (MODE == 0 && fits) rec_hdr_pos[rec - 1u] = i0 + k;
                }
                in_hdr = h;
            }
            const uint32_t c = cls_tab[byte_of(raw, k)];
            if (MODE == 0 && c == 9u && !in_hdr && rec >= 1u && byte_of(raw, k) != '\n') {
                // Whitespace in a sequence line: harmless at the ends (the reference strips them, F:149 -- every '\r' of a
                // CRLF file lands here and is cleared by its two neighbours), refused INSIDE the line, where the
                // reference keeps it as a character of the sequence (same rule as frisk_b200_fasta_scan).
                bool before = false, after = false;
                uint64_t p = i0 + k;
                int s = 0;
                for (; s < 4096 && p > 0; ++s) {
                    const uint32_t b = t[--p];
                    if (b == '\n') break;
                    if (cls_tab[b] != 9u) { before = true; break; }
                }
                if (s == 4096) before = true;                          // a whitespace run this long is not a line ending
                p = i0 + k;
                for (s = 0; s < 4096 && p + 1 < n; ++s) {
                    if (p + 1 >= avail) { *ambig = 1u; break; }         // the rest of the line is not on the device yet
                    const uint32_t b = t[++p];
                    if (b == '\n') break;
                    if (cls_tab[b] != 9u) { after = true; break; }
                }
                if (s == 4096) after = true;
                if (before && after) counters[kCtrWhitespace] = 1ull;
            }
            if (!in_hdr && c != 9u && rec >= 1u) {
                base_mask |= 1u << k;
                ++cnt;
                non_upper += c >= 4u;
                lower += (c >= 4u) & (c < 8u);
                cls16[k >> 3] |= c << (4 * (k & 7));
            }
        }
    }
    bool reset_before;
    const uint32_t carry = block_excl_seg<uint32_t>(seen_hdr, cnt, sm, smf, &reset_before);

    if (MODE == 0) {
        if (seen_hdr) {
            const uint32_t amount = carry + pre;                        // this tile's share of the record the header closes
            if (fits && rec_start >= 1u && amount) atomicAdd(&rec_len[rec_start - 1u], (unsigned long long)amount);
            if (!reset_before) tile_pre[tile] = amount;
        }
        if (tid == kTT - 1) {
            const uint32_t s = seen_hdr ? cnt : carry + cnt;            // bases after the tile's last header (all, if none)
            const uint32_t rec_end = rec_start + n_hdr_t;
            if (fits && rec_end >= 1u && s) atomicAdd(&rec_len[rec_end - 1u], (unsigned long long)s);
            tile_post[tile] = s;
            if (!reset_before && !seen_hdr) tile_pre[tile] = s;
        }
        uint32_t tot_non, tot_low;
        block_excl_sum<uint32_t>(non_upper, sm, &tot_non);
        block_excl_sum<uint32_t>(lower, sm, &tot_low);
        if (tid == 0) {
            if (tot_non) atomicAdd(&counters[kCtrNonUpper], (unsigned long long)tot_non);
            if (tot_low) atomicAdd(&counters[kCtrLower], (unsigned long long)tot_low);
        }
    } else {
        // A thread's bases land on consecutive packed positions (per record), i.e. in at most two
        // code words and two mask words: assemble them in registers and OR them into the zeroed
        // planes (neighbouring threads share words; all-zero contributions are skipped).
        if (base_mask) {
            uint32_t rec = rec_start;
            unsigned long long pos = (rec >= 1u ? scaf_off[rec - 1u] : 0ull) + (unsigned long long)carry +
                                     (reset_before ? 0ull : tile_base_in[tile]);
            unsigned long long cidx = pos >> 4, midx = pos >> 5;
            uint32_t cacc = 0, iacc = 0, lacc = 0;
            auto flush_codes = [&]() { if (cacc) atomicOr(&codes[cidx], cacc); cacc = 0; };
            auto flush_masks = [&]() {
                if (iacc) atomicOr(&inv[midx], iacc);
                if (lacc && low) atomicOr(&low[midx], lacc);
                iacc = 0; lacc = 0;
            };
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if ((hdr_mask >> k) & 1u) {
                    flush_codes(); flush_masks();
                    ++rec;
                    pos = scaf_off[rec - 1u];
                    cidx = pos >> 4; midx = pos >> 5;
                }
                if ((base_mask >> k) & 1u) {
                    if ((pos >> 4) != cidx) { flush_codes(); cidx = pos >> 4; }
                    if ((pos >> 5) != midx) { flush_masks(); midx = pos >> 5; }
                    const uint32_t c = (cls16[k >> 3] >> (4 * (k & 7))) & 15u;
                    const uint32_t b32 = 0x80000000u >> ((uint32_t)pos & 31u);
                    if (c < 8u) cacc |= (c & 3u) << (30u - 2u * ((uint32_t)pos & 15u)); else iacc |= b32;
                    if ((c >= 4u) & (c < 8u)) lacc |= b32;
                    ++pos;
                }
            }
            flush_codes(); flush_masks();
        }
    }
}

// ---- pass 4: the padding after every record (and the >= 128 trailing bases) is flagged invalid ----
__global__ void __launch_bounds__(256)
fasta_padding_kernel(const unsigned long long* __restrict__ scaf_off, const unsigned long long* __restrict__ rec_len,
                     uint64_t n_rec, uint64_t padded_len, uint32_t* __restrict__ inv) {
    const uint64_t r = (uint64_t)blockIdx.x * 256u + threadIdx.x;
    if (r >= (n_rec ? n_rec : 1)) return;
    uint64_t a = n_rec ? scaf_off[r] + rec_len[r] : 0;                  // first padding base
    const uint64_t b = (r + 1 < n_rec) ? scaf_off[r + 1] : padded_len;  // a multiple of 128
    if (a & 31u) {
        atomicOr(&inv[a >> 5], 0xffffffffu >> (uint32_t)(a & 31u));     // shares its word with the record's last bases
        a = (a | 31u) + 1u;
    }
    for (; a < b; a += 32) inv[a >> 5] = 0xffffffffu;
}

// ---- streamed planes: the layout of the records seen so far, on the device ---------------------------------------
// Runs behind the tile passes of every chunk.  Records [lay[kLayAssigned], n_rec) get their offset (frisk_b200_pack_layout's
// rule: next = align_up(off + len + 1, 128) -- the length of every record but the open one is final); out[0..1] = the plane
// words that are final now and not yet counted (the count kernel looks one word ahead), out[2..3] = the records whose padding
// can be flagged now.  final: the text is complete, the open record closes and the trailing padding is part of the ranges.
__device__ __forceinline__ unsigned long long align128(unsigned long long v) { return (v + 127ull) & ~127ull; }

__global__ void __launch_bounds__(kST)
fasta_layout_kernel(const unsigned long long* __restrict__ rec_len, unsigned long long* __restrict__ scaf_off,
                    unsigned long long* __restrict__ counters, unsigned long long* __restrict__ out, uint64_t rec_cap,
                    uint64_t plane_cap, int count_now, int final) {
    __shared__ unsigned long long sm[64];
    const unsigned long long n_rec = counters[kCtrRecords];
    if (counters[kCtrOverflow] != 0ull || n_rec > rec_cap) {           // (uniform: nothing below is touched)
        if (threadIdx.x == 0) {
            counters[kCtrOverflow] = 1ull;
            out[0] = out[1] = counters[kLayCounted];
            out[2] = out[3] = counters[kLayPadded];
        }
        return;
    }
    const unsigned long long r0 = counters[kLayAssigned];
    const unsigned long long base = r0 ? scaf_off[r0 - 1] : 0ull;
    const unsigned long long cnt = n_rec - r0;
    const unsigned long long per = (cnt + kST - 1) / kST;
    const unsigned long long lo = r0 + min((unsigned long long)threadIdx.x * per, cnt), hi = min(lo + per, n_rec);
    unsigned long long sum = 0;
    for (unsigned long long e = lo; e < hi; ++e) sum += e ? align128(rec_len[e - 1] + 1ull) : 0ull;
    unsigned long long total;
    unsigned long long run = base + block_excl_sum<unsigned long long>(sum, sm, &total);
    for (unsigned long long e = lo; e < hi; ++e) {
        run += e ? align128(rec_len[e - 1] + 1ull) : 0ull;
        scaf_off[e] = run;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long pos_hi = n_rec ? scaf_off[n_rec - 1] + counters[kCtrBaseCarry] : 0ull;   // first unwritten base
        const unsigned long long padded = (n_rec ? align128(pos_hi + 1ull) : 0ull) + 128ull;
        const unsigned long long w_lo = counters[kLayCounted], p_lo = counters[kLayPadded];
        if (padded > plane_cap) {
            counters[kCtrOverflow] = 1ull;
            out[0] = out[1] = w_lo; out[2] = out[3] = p_lo;
            return;
        }
        counters[kLayAssigned] = n_rec;
        unsigned long long w_hi = w_lo, p_hi = final ? n_rec : max(p_lo, n_rec ? n_rec - 1ull : 0ull);
        if (final) w_hi = padded / 32ull - 1ull;
        else if (count_now && (pos_hi >> 5) >= 1ull) w_hi = max(w_lo, (pos_hi >> 5) - 1ull);
        if (!count_now) w_hi = w_lo;
        out[0] = w_lo; out[1] = w_hi; counters[kLayCounted] = w_hi;
        out[2] = p_lo; out[3] = p_hi; counters[kLayPadded] = p_hi;
        if (final) counters[kLayPaddedLen] = padded;
    }
}

// padding of records [range[0], range[1]) (device-side range of fasta_layout_kernel); the last record of a finished text is
// padded up to the layout's padded length
__global__ void __launch_bounds__(256)
fasta_padding_range_kernel(const unsigned long long* __restrict__ scaf_off, const unsigned long long* __restrict__ rec_len,
                           const unsigned long long* __restrict__ range, const unsigned long long* __restrict__ counters, int final,
                           uint32_t* __restrict__ inv) {
    const unsigned long long n_rec = counters[kCtrRecords];
    if (final && n_rec == 0ull && counters[kCtrOverflow] == 0ull) {     // no record at all: 128 invalid bases
        if (blockIdx.x == 0 && threadIdx.x < 4) inv[threadIdx.x] = 0xffffffffu;
        return;
    }
    for (unsigned long long r = range[0] + (unsigned long long)blockIdx.x * 256u + threadIdx.x; r < range[1];
         r += (unsigned long long)gridDim.x * 256u) {
        unsigned long long a = scaf_off[r] + rec_len[r];
        const unsigned long long b = (r + 1 < n_rec) ? scaf_off[r + 1] : counters[kLayPaddedLen];
        if (a & 31u) {
            atomicOr(&inv[a >> 5], 0xffffffffu >> (uint32_t)(a & 31u));
            a = (a | 31u) + 1u;
        }
        for (; a < b; a += 32) inv[a >> 5] = 0xffffffffu;
    }
}

}  // namespace

struct frisk_b200_fasta {
    uint64_t n = 0, n_tiles = 0, n_rec = 0, padded_len = 128;
    uint64_t stats[3] = {0, 0, 0};
    uint8_t* d_text = nullptr;
    uint32_t *d_nhdr = nullptr, *d_rec_base = nullptr, *d_pre = nullptr, *d_post = nullptr;
    uint8_t *d_key = nullptr, *d_carry = nullptr;
    unsigned long long *d_base_in = nullptr, *d_hdr_pos = nullptr, *d_len = nullptr, *d_scaf_off = nullptr, *d_counters = nullptr;
    uint32_t *d_codes = nullptr, *d_inv = nullptr, *d_low = nullptr;     // planes built by the open itself (streamed planes)
    bool planes_ready = false;
    std::vector<uint64_t> name_off, seq_len, scaf_off, hdr_pos;
    std::vector<uint32_t> name_len;
};

namespace {
int free_all(frisk_b200_fasta* h, cudaStream_t st) {
    void* ptrs[] = {h->d_text, h->d_nhdr, h->d_rec_base, h->d_pre, h->d_post, h->d_key, h->d_carry,
                    h->d_base_in, h->d_hdr_pos, h->d_len, h->d_scaf_off, h->d_counters, h->d_codes, h->d_inv, h->d_low};
    for (void* p : ptrs)
        if (p) FRISK_CK(cudaFreeAsync(p, st));
    h->d_text = nullptr; h->d_nhdr = h->d_rec_base = h->d_pre = h->d_post = nullptr; h->d_key = h->d_carry = nullptr;
    h->d_base_in = h->d_hdr_pos = h->d_len = h->d_scaf_off = h->d_counters = nullptr;
    h->d_codes = h->d_inv = h->d_low = nullptr;
    h->planes_ready = false;
    return FRISK_OK;
}

// ---- the upload lane: text chunks travel on a second stream while the tile passes of the previous chunk run -------------
constexpr int kMaxChunks = 8;
constexpr uint64_t kMinChunkTiles = 512;            // 2 MiB of text: below that a chunk's copy is shorter than its launches
constexpr int kRetryExact = -1000;                  // internal: the chunked / speculative open could not decide, open again
constexpr uint64_t kFirstFetch = 4096;              // records read back with the counters (one synchronisation when R <= this)

struct UploadLane {
    cudaStream_t copy = nullptr;
    cudaEvent_t ready = nullptr, done[kMaxChunks] = {};
};
std::mutex g_lane_mu;
std::atomic<uint64_t> g_open_stats[2];
UploadLane g_lane[64];

int upload_lane(UploadLane** out) {
    int dev = 0;
    FRISK_CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return FRISK_E_UNSUPPORTED;
    std::lock_guard<std::mutex> lk(g_lane_mu);
    UploadLane& l = g_lane[dev];
    if (!l.copy) {
        FRISK_CK(cudaStreamCreateWithFlags(&l.copy, cudaStreamNonBlocking));
        FRISK_CK(cudaEventCreateWithFlags(&l.ready, cudaEventDisableTiming));
        for (auto& e : l.done) FRISK_CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    *out = &l;
    return FRISK_OK;
}

// exact = false: the text goes up in chunks and every chunk is tokenised while the next one is on the bus; the record table is
//   sized by a guess (one record per 64 bytes of text + 4096).  A line decision that needs a byte of a later chunk, or more
//   records than the guess, returns kRetryExact.
// exact = true: one copy, records counted (one more synchronisation) before the record table is allocated.  Cannot fail that way.
// sink (nullable, !exact only): the PLANES are built during the open as well -- allocated by the handle for an upper bound of
//   the layout (text bytes + 128 per guessed record), each chunk laid out (fasta_layout_kernel), packed and padded behind its
//   tile passes, and the plane words that became final handed to sink->on_range while the next chunk is still on the bus.
int open_impl(frisk_b200_fasta* h, const char* text, uint64_t n, cudaStream_t st, bool exact, frisk_internal::IngestSink* sink) {
    int rc = frisk_internal::pool_ready();
    if (rc) return rc;
    if (exact) sink = nullptr;
    h->n = n;
    h->n_tiles = n ? (n + kTile - 1) / kTile : 0;
    if (h->n_tiles > 0x7fffffffull) return FRISK_E_UNSUPPORTED;         // 8 TB of text per call
    const uint64_t T = h->n_tiles;
    if (!T && sink) return kRetryExact;                                 // (empty text: nothing to stream)
    if (T) {
        const uint64_t padded_text = T * kTile + 16;
        FRISK_CK(cudaMallocAsync((void**)&h->d_text, padded_text, st));
        FRISK_CK(cudaMemsetAsync(h->d_text + n, '\n', padded_text - n, st));
        FRISK_CK(cudaMallocAsync((void**)&h->d_nhdr, T * 4, st));
        FRISK_CK(cudaMallocAsync((void**)&h->d_rec_base, T * 4, st));
        FRISK_CK(cudaMallocAsync((void**)&h->d_pre, T * 4, st));
        FRISK_CK(cudaMallocAsync((void**)&h->d_post, T * 4, st));
        FRISK_CK(cudaMallocAsync((void**)&h->d_key, T, st));
        FRISK_CK(cudaMallocAsync((void**)&h->d_carry, T, st));
        FRISK_CK(cudaMallocAsync((void**)&h->d_base_in, T * 8, st));
        constexpr size_t kCtrWords = kCtrCount + (size_t)kRangeSlots * kMaxChunks;
        FRISK_CK(cudaMallocAsync((void**)&h->d_counters, kCtrWords * 8, st));
        FRISK_CK(cudaMemsetAsync(h->d_counters, 0, kCtrWords * 8, st));
        uint32_t* const d_ambig = reinterpret_cast<uint32_t*>(h->d_counters + kCtrAmbig);
        uint32_t* const d_key_carry = reinterpret_cast<uint32_t*>(h->d_counters + kCtrKeyCarry);

        uint64_t cap = 0, plane_cap = 0;
        auto alloc_records = [&](uint64_t R) -> int {
            cap = R;
            FRISK_CK(cudaMallocAsync((void**)&h->d_hdr_pos, R * 8, st));
            FRISK_CK(cudaMallocAsync((void**)&h->d_len, R * 8, st));
            FRISK_CK(cudaMallocAsync((void**)&h->d_scaf_off, R * 8, st));
            FRISK_CK(cudaMemsetAsync(h->d_len, 0, R * 8, st));
            return FRISK_OK;
        };
        auto tile_passes = [&](uint64_t t0, uint64_t t1, uint64_t avail) {
            fasta_tile_kernel<0><<<(unsigned)(t1 - t0), kTT, 0, st>>>(h->d_text, n, h->d_rec_base, h->d_carry, h->d_hdr_pos, h->d_len,
                                                                      h->d_pre, h->d_post, h->d_counters, nullptr, nullptr, nullptr,
                                                                      nullptr, nullptr, t0, avail, cap);
            fasta_scan2_kernel<<<1, kST, 0, st>>>(h->d_nhdr, h->d_pre, h->d_post, t0, t1, h->d_base_in,
                                                  h->d_counters + kCtrBaseCarry);
        };
        const uint64_t min_chunk_tiles = frisk_internal::g_ingest_chunk_tiles > 0 ? (uint64_t)frisk_internal::g_ingest_chunk_tiles : kMinChunkTiles;
        const int n_chunks = exact ? 1 : (int)std::max<uint64_t>(1, std::min<uint64_t>(kMaxChunks, T / min_chunk_tiles));
        if (!exact && (rc = alloc_records(n / 64 + 4096))) return rc;
        if (sink) {
            // every base is a byte of the text; every record adds < 128 + 128 bases of padding: room for one record per KiB
            plane_cap = ((n + 127) & ~127ull) + 256ull * (n / 1024 + 4096) + 256;
            FRISK_CK(cudaMallocAsync((void**)&h->d_codes, plane_cap / 4, st));
            FRISK_CK(cudaMallocAsync((void**)&h->d_inv, plane_cap / 8, st));
            FRISK_CK(cudaMallocAsync((void**)&h->d_low, plane_cap / 8, st));
            FRISK_CK(cudaMemsetAsync(h->d_codes, 0, plane_cap / 4, st));
            FRISK_CK(cudaMemsetAsync(h->d_inv, 0, plane_cap / 8, st));
            FRISK_CK(cudaMemsetAsync(h->d_low, 0, plane_cap / 8, st));
        }
        UploadLane* lane = nullptr;
        if (n_chunks > 1 || sink) {
            if ((rc = upload_lane(&lane))) return rc;
            FRISK_CK(cudaEventRecord(lane->ready, st));                 // the text buffer exists (stream-ordered allocation)
            FRISK_CK(cudaStreamWaitEvent(lane->copy, lane->ready, 0));
        }
        const uint64_t per = (T + n_chunks - 1) / n_chunks;
        const int count_every = n_chunks > 4 ? 2 : 1;                   // (a count launch has a fixed cost: fewer, larger ranges)
        int last_chunk = 0;
        for (int c = 0; c < n_chunks; ++c)
            if (std::min<uint64_t>((uint64_t)c * per, T) < T) last_chunk = c;
        for (int c = 0; c <= last_chunk; ++c) {
            const uint64_t t0 = (uint64_t)c * per, t1 = std::min<uint64_t>(t0 + per, T);
            const uint64_t b0 = t0 * kTile, b1 = std::min<uint64_t>(t1 * kTile, n);
            const bool final = c == last_chunk;
            if (lane) {
                FRISK_CK(cudaMemcpyAsync(h->d_text + b0, text + b0, b1 - b0, cudaMemcpyHostToDevice, lane->copy));
                FRISK_CK(cudaEventRecord(lane->done[c], lane->copy));
                if (final && sink && sink->uploaded_mark) FRISK_CK(cudaEventRecord(sink->uploaded_mark, lane->copy));
                FRISK_CK(cudaStreamWaitEvent(st, lane->done[c], 0));
            } else {
                FRISK_CK(cudaMemcpyAsync(h->d_text + b0, text + b0, b1 - b0, cudaMemcpyHostToDevice, st));
            }
            fasta_lines_kernel<<<(unsigned)(t1 - t0), kTT, 0, st>>>(h->d_text, n, h->d_nhdr, h->d_key, t0, b1, d_ambig);
            fasta_scan1_kernel<<<1, kST, 0, st>>>(h->d_nhdr, h->d_key, t0, t1, h->d_rec_base, h->d_carry, h->d_counters, d_key_carry);
            if (!exact) tile_passes(t0, t1, b1);
            if (sink) {
                unsigned long long* const d_range = h->d_counters + kCtrCount + (size_t)kRangeSlots * c;
                const bool count_now = final || (c % count_every) == count_every - 1;
                fasta_layout_kernel<<<1, kST, 0, st>>>(h->d_len, h->d_scaf_off, h->d_counters, d_range, cap, plane_cap,
                                                       (int)(count_now && (bool)sink->on_range), (int)final);
                if (final) {                                            // the record table is final: the host can have it now,
                    FRISK_CK(cudaEventRecord(lane->ready, st));         // while the last chunk is still being packed and counted
                    FRISK_CK(cudaStreamWaitEvent(lane->copy, lane->ready, 0));
                }
                fasta_tile_kernel<1><<<(unsigned)(t1 - t0), kTT, 0, st>>>(h->d_text, n, h->d_rec_base, h->d_carry, nullptr, nullptr,
                                                                          nullptr, nullptr, h->d_counters, h->d_base_in, h->d_scaf_off,
                                                                          h->d_codes, h->d_inv, h->d_low, t0, b1, 0);
                fasta_padding_range_kernel<<<final ? 32 : 8, 256, 0, st>>>(h->d_scaf_off, h->d_len, d_range + 2, h->d_counters, (int)final,
                                                                           h->d_inv);
                FRISK_CK(cudaGetLastError());
                if (count_now && sink->on_range) {
                    const uint64_t words_hint = ((b1 - (uint64_t)(c / count_every) * count_every * per * kTile) >> 5) + 8;
                    if ((rc = sink->on_range(h->d_codes, h->d_inv, h->d_low, d_range, words_hint, st))) return rc;
                }
            }
        }
        FRISK_CK(cudaGetLastError());
        unsigned long long ctr[kCtrCount] = {};
        if (exact) {
            FRISK_CK(cudaMemcpyAsync(ctr, h->d_counters, 8, cudaMemcpyDeviceToHost, st));
            FRISK_CK(cudaStreamSynchronize(st));
            if ((rc = alloc_records(ctr[kCtrRecords] ? ctr[kCtrRecords] : 1))) return rc;
            tile_passes(0, T, n);
            FRISK_CK(cudaGetLastError());
        }
        cudaStream_t back = sink ? lane->copy : st;                     // the stream the record table comes back on
        const uint64_t first = std::min<uint64_t>(cap, kFirstFetch);
        h->seq_len.resize(first);
        h->hdr_pos.resize(first);
        FRISK_CK(cudaMemcpyAsync(h->seq_len.data(), h->d_len, first * 8, cudaMemcpyDeviceToHost, back));
        FRISK_CK(cudaMemcpyAsync(h->hdr_pos.data(), h->d_hdr_pos, first * 8, cudaMemcpyDeviceToHost, back));
        FRISK_CK(cudaMemcpyAsync(ctr, h->d_counters, kCtrCount * 8, cudaMemcpyDeviceToHost, back));
        FRISK_CK(cudaStreamSynchronize(back));
        if (!exact && (ctr[kCtrAmbig] || ctr[kCtrOverflow])) return kRetryExact;
        if (ctr[kCtrWhitespace]) return FRISK_E_FORMAT;                // whitespace inside a sequence line
        const uint64_t n_rec = ctr[kCtrRecords];
        h->n_rec = n_rec;
        h->seq_len.resize(n_rec);
        h->hdr_pos.resize(n_rec);
        if (n_rec > first) {
            FRISK_CK(cudaMemcpyAsync(h->seq_len.data() + first, h->d_len + first, (n_rec - first) * 8, cudaMemcpyDeviceToHost, back));
            FRISK_CK(cudaMemcpyAsync(h->hdr_pos.data() + first, h->d_hdr_pos + first, (n_rec - first) * 8, cudaMemcpyDeviceToHost, back));
            FRISK_CK(cudaStreamSynchronize(back));
        }
        h->stats[1] = ctr[kCtrNonUpper];
        h->stats[2] = ctr[kCtrLower];
        if (sink) h->padded_len = ctr[kLayPaddedLen];                  // (checked against the host's layout below)
    }
    // names (F:156) and the 128-base aligned layout
    const uint64_t R = h->n_rec;
    h->name_off.resize(R); h->name_len.resize(R); h->scaf_off.resize(R);
    const unsigned char* tt = reinterpret_cast<const unsigned char*>(text);
    uint64_t total = 0;
    for (uint64_t r = 0; r < R; ++r) {
        if (!frisk_internal::parse_header_name(tt, n, h->hdr_pos[r], &h->name_off[r], &h->name_len[r])) return FRISK_E_FORMAT;
        total += h->seq_len[r];
    }
    h->stats[0] = total;
    const uint64_t device_padded = h->padded_len;
    rc = frisk_b200_pack_layout(h->seq_len.data(), R, h->scaf_off.data(), &h->padded_len);
    if (rc) return rc;
    if (sink) {
        if (device_padded != h->padded_len) return kRetryExact;         // (cannot happen: both sides apply the same rule)
        h->planes_ready = true;
        sink->counted = (bool)sink->on_range;
    } else if (R) {
        FRISK_CK(cudaMemcpyAsync(h->d_scaf_off, h->scaf_off.data(), R * 8, cudaMemcpyHostToDevice, st));
    }
    return FRISK_OK;
}

int pack_impl(frisk_b200_fasta* h, uint32_t* d_codes, uint32_t* d_inv, uint32_t* d_low, cudaStream_t st) {
    const uint64_t P = h->padded_len;
    FRISK_CK(cudaMemsetAsync(d_codes, 0, P / 4, st));
    FRISK_CK(cudaMemsetAsync(d_inv, 0, P / 8, st));
    if (d_low) FRISK_CK(cudaMemsetAsync(d_low, 0, P / 8, st));
    if (h->n_tiles && h->n_rec)
        fasta_tile_kernel<1><<<(unsigned)h->n_tiles, kTT, 0, st>>>(h->d_text, h->n, h->d_rec_base, h->d_carry, nullptr, nullptr,
                                                                   nullptr, nullptr, h->d_counters, h->d_base_in, h->d_scaf_off,
                                                                   d_codes, d_inv, d_low, 0, h->n, 0);
    const uint64_t R = h->n_rec ? h->n_rec : 1;
    fasta_padding_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(h->d_scaf_off, h->d_len, h->n_rec, P, d_inv);
    FRISK_CK(cudaGetLastError());
    return FRISK_OK;
}

// open with retry; sink (nullable): planes built by the open and owned by the handle (streamed, or -- after a retry --
// packed in one piece behind the exact open; sink->counted tells which)
int open_any(frisk_b200_fasta* h, const char* text, uint64_t n, cudaStream_t st, frisk_internal::IngestSink* sink) {
    const bool exact = frisk_internal::g_ingest_exact != 0;
    if (sink) sink->counted = false;
    int rc = open_impl(h, text, n, st, exact, sink);
    if (rc == FRISK_OK && !exact) g_open_stats[0].fetch_add(1);
    if (rc == kRetryExact) {
        if (n) g_open_stats[1].fetch_add(1);
        if (sink && sink->abandon && (rc = sink->abandon(st))) return rc;
        if ((rc = free_all(h, st))) return rc;
        *h = frisk_b200_fasta();
        rc = open_impl(h, text, n, st, true, nullptr);
    }
    if (rc == FRISK_OK && sink && !h->planes_ready) {
        const uint64_t P = h->padded_len;
        FRISK_CK(cudaMallocAsync((void**)&h->d_codes, P / 4, st));
        FRISK_CK(cudaMallocAsync((void**)&h->d_inv, P / 8, st));
        if (h->stats[2]) FRISK_CK(cudaMallocAsync((void**)&h->d_low, P / 8, st));
        if ((rc = pack_impl(h, h->d_codes, h->d_inv, h->d_low, st))) return rc;
        h->planes_ready = true;
        sink->counted = false;
    }
    return rc;
}
}  // namespace

int frisk_internal::fasta_open_planes(const char* text, uint64_t n, cudaStream_t st, IngestSink* sink, frisk_b200_fasta** out) {
    if (!out || !sink || (!text && n)) return FRISK_E_INVALID;
    *out = nullptr;
    frisk_b200_fasta* h = new (std::nothrow) frisk_b200_fasta();
    if (!h) return FRISK_E_INVALID;
    int rc;
    try {
        rc = open_any(h, text, n, st, sink);
    } catch (...) {
        rc = FRISK_E_CAPACITY;
    }
    if (rc) {
        free_all(h, st);
        delete h;
        return rc;
    }
    *out = h;
    return FRISK_OK;
}

extern "C" {

int frisk_b200_fasta_info(const frisk_b200_fasta* h, uint64_t* n_records, uint64_t* padded_len, uint64_t stats[3]) {
    if (!h) return FRISK_E_INVALID;
    if (n_records) *n_records = h->n_rec;
    if (padded_len) *padded_len = h->padded_len;
    if (stats) { stats[0] = h->stats[0]; stats[1] = h->stats[1]; stats[2] = h->stats[2]; }
    return FRISK_OK;
}

int frisk_b200_fasta_open(const char* text, uint64_t n, void* stream, frisk_b200_fasta** out, uint64_t* n_records,
                          uint64_t* padded_len, uint64_t stats[3]) {
    if (!out || (!text && n)) return FRISK_E_INVALID;
    *out = nullptr;
    if (frisk_b200_device_count() <= 0) return FRISK_E_NO_DEVICE;
    frisk_b200_fasta* h = new (std::nothrow) frisk_b200_fasta();
    if (!h) return FRISK_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    try {
        rc = open_any(h, text, n, st, nullptr);
    } catch (...) {                                   // std::bad_alloc of the record table: nothing crosses the ABI
        rc = FRISK_E_CAPACITY;
    }
    if (rc) {
        free_all(h, st);
        delete h;
        return rc;
    }
    if (n_records) *n_records = h->n_rec;
    if (padded_len) *padded_len = h->padded_len;
    if (stats) { stats[0] = h->stats[0]; stats[1] = h->stats[1]; stats[2] = h->stats[2]; }
    *out = h;
    return FRISK_OK;
}

int frisk_b200_fasta_records(const frisk_b200_fasta* h, uint64_t* name_off, uint32_t* name_len, uint64_t* seq_len,
                             uint64_t* scaf_off) {
    if (!h) return FRISK_E_INVALID;
    const size_t R = (size_t)h->n_rec;
    if (name_off && R) memcpy(name_off, h->name_off.data(), R * 8);
    if (name_len && R) memcpy(name_len, h->name_len.data(), R * 4);
    if (seq_len && R) memcpy(seq_len, h->seq_len.data(), R * 8);
    if (scaf_off && R) memcpy(scaf_off, h->scaf_off.data(), R * 8);
    return FRISK_OK;
}

int frisk_b200_fasta_pack(frisk_b200_fasta* h, uint32_t* d_codes, uint32_t* d_inv, uint32_t* d_low, void* stream) {
    if (!h || !d_codes || !d_inv) return FRISK_E_INVALID;
    return pack_impl(h, d_codes, d_inv, d_low, (cudaStream_t)stream);
}

int frisk_b200_fasta_planes(const frisk_b200_fasta* h, const uint32_t** d_codes, const uint32_t** d_inv, const uint32_t** d_low) {
    if (!h || !h->planes_ready) return FRISK_E_INVALID;
    if (d_codes) *d_codes = h->d_codes;
    if (d_inv) *d_inv = h->d_inv;
    if (d_low) *d_low = h->stats[2] ? h->d_low : nullptr;            // no lower-case base: no plane (as frisk_b200_pack reports it)
    return FRISK_OK;
}

int frisk_b200_fasta_open_stats(uint64_t out[2]) {
    if (!out) return FRISK_E_INVALID;
    out[0] = g_open_stats[0].load();
    out[1] = g_open_stats[1].load();
    return FRISK_OK;
}

int frisk_b200_fasta_close(frisk_b200_fasta* h, void* stream) {
    if (!h) return FRISK_OK;
    const int rc = free_all(h, (cudaStream_t)stream);
    delete h;
    return rc;
}

}  // extern "C"
