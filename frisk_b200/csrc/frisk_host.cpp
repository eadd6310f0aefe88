// frisk_b200 host code: FASTA scanning, 2-bit packing into (pinned) planes, window enumeration.
// Replaces iterFasta / countN / crawlGenome's enumeration (reference frisk/__init__.py, "F:").
// Pure CPU; no CUDA calls here.
#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <string>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/frisk_b200.h"
#include "frisk_internal.h"

namespace {

// class of a byte: 0..3 = A,T,G,C upper-case; 4..7 = a,t,g,c lower-case; 8 = other; 9 = whitespace
struct ByteClass {
    uint8_t t[256];
    ByteClass() {
        for (int i = 0; i < 256; ++i) t[i] = 8;
        t[(int)'A'] = 0; t[(int)'T'] = 1; t[(int)'G'] = 2; t[(int)'C'] = 3;      // F:70 order
        t[(int)'a'] = 4; t[(int)'t'] = 5; t[(int)'g'] = 6; t[(int)'c'] = 7;
        // characters removed by the reference's line.strip() / blank-line skip (F:149-151)
        t[(int)' '] = t[(int)'\t'] = t[(int)'\n'] = t[(int)'\r'] = t[(int)'\v'] = t[(int)'\f'] = 9;
    }
};
const ByteClass kClass;

inline bool is_space(unsigned char c) { return kClass.t[c] == 9; }

inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

// Pack one scaffold.  dst planes are indexed in absolute packed coordinates; `base` is the
// scaffold's first base offset (multiple of 128), so whole words belong to this scaffold only.
int pack_one(const unsigned char* s, const unsigned char* e, uint64_t expect, uint64_t base, uint64_t next_base,
             uint32_t* codes, uint32_t* inv, uint32_t* low, uint64_t stats[3]) {
    uint64_t n = 0, non = 0, nlow = 0;
    uint32_t cw = 0, iw = 0, lw = 0;
    uint64_t pos = base;
    for (const unsigned char* p = s; p < e; ++p) {
        // eight plain upper-case bases at once (a 60-column line is 7 such groups + 4 bytes + the newline): is every byte one of
        // A, C, G, T?  then 2 bits each straight from the ASCII codes -- bits 2..1 of A/C/T/G are 00/01/10/11, the F:70 order
        // A,T,G,C is (b0, b1^b0) -- gathered by one multiply per four characters.  No mask bits to set.
        while (p + 8 <= e && n + 8 <= expect) {
            uint64_t x;
            memcpy(&x, p, 8);
            const uint64_t m = 0x7F7F7F7F7F7F7F7Full;
            const uint64_t xa = x ^ 0x4141414141414141ull, xc = x ^ 0x4343434343434343ull, xg = x ^ 0x4747474747474747ull,
                           xt = x ^ 0x5454545454545454ull;
            if ((((xa & m) + m) | xa) & (((xc & m) + m) | xc) & (((xg & m) + m) | xg) & (((xt & m) + m) | xt) & 0x8080808080808080ull) break;
            const uint64_t y = x >> 1;
            const uint64_t c2 = ((y & 0x0101010101010101ull) << 1) | ((y ^ (y >> 1)) & 0x0101010101010101ull);
            const uint32_t bits16 = ((((uint32_t)c2 * 0x40100401u) >> 24) << 8) | (((uint32_t)(c2 >> 32) * 0x40100401u) >> 24);
            const uint32_t j = (uint32_t)(pos & 15);
            if (j <= 8) {
                cw |= bits16 << (16 - 2 * j);
                if (j == 8) { codes[pos >> 4] = cw; cw = 0; }
            } else {
                cw |= bits16 >> (2 * (j - 8));
                codes[pos >> 4] = cw;
                cw = bits16 << (32 - 2 * (j - 8));
            }
            const uint64_t np = pos + 8;
            if ((np ^ pos) >> 5) {                       // a mask word was completed (the eight bases add no mask bits)
                inv[pos >> 5] = iw; iw = 0;
                if (low) low[pos >> 5] = lw;
                lw = 0;
            }
            pos = np; n += 8; p += 8;
        }
        if (p >= e) break;
        const uint8_t c = kClass.t[*p];
        if (c == 9) continue;
        if (n == expect) return FRISK_E_FORMAT;
        const uint32_t j32 = (uint32_t)(pos & 31);
        if (c < 4) cw |= (uint32_t)c << (30 - 2 * (j32 & 15));
        else if (c < 8) { cw |= (uint32_t)(c - 4) << (30 - 2 * (j32 & 15)); lw |= 0x80000000u >> j32; ++nlow; ++non; }
        else { iw |= 0x80000000u >> j32; ++non; }
        ++pos; ++n;
        if ((pos & 15) == 0) { codes[(pos >> 4) - 1] = cw; cw = 0; }
        if ((pos & 31) == 0) {
            inv[(pos >> 5) - 1] = iw; iw = 0;
            if (low) low[(pos >> 5) - 1] = lw;
            lw = 0;
        }
    }
    if (n != expect) return FRISK_E_FORMAT;
    // padding up to the next scaffold (or the end): invalid, code 0
    while (pos < next_base) {
        const uint32_t j32 = (uint32_t)(pos & 31);
        if (j32 == 0 && next_base - pos >= 32 && (pos & 15) == 0) {   // whole mask word of padding
            codes[pos >> 4] = 0; codes[(pos >> 4) + 1] = 0;
            inv[pos >> 5] = 0xffffffffu;
            if (low) low[pos >> 5] = 0;
            pos += 32;
            continue;
        }
        iw |= 0x80000000u >> j32;
        ++pos;
        if ((pos & 15) == 0) { codes[(pos >> 4) - 1] = cw; cw = 0; }
        if ((pos & 31) == 0) {
            inv[(pos >> 5) - 1] = iw; iw = 0;
            if (low) low[(pos >> 5) - 1] = lw;
            lw = 0;
        }
    }
    stats[0] += n; stats[1] += non; stats[2] += nlow;
    return FRISK_OK;
}

}  // namespace

bool frisk_internal::parse_header_name(const unsigned char* t, uint64_t n, uint64_t line_start, uint64_t* name_off,
                                       uint32_t* name_len) {
    const void* nl = memchr(t + line_start, '\n', n - line_start);
    uint64_t a = line_start, b = nl ? (uint64_t)((const unsigned char*)nl - t) : n;
    while (a < b && is_space(t[a])) ++a;            // line.strip()
    while (b > a && is_space(t[b - 1])) --b;
    // name = line.strip('>').split()[0]  (F:156)
    uint64_t x = a, y = b;
    while (x < y && t[x] == '>') ++x;
    while (y > x && t[y - 1] == '>') --y;
    while (x < y && is_space(t[x])) ++x;
    uint64_t z = x;
    while (z < y && !is_space(t[z])) ++z;
    if (z == x) return false;
    *name_off = x;
    *name_len = (uint32_t)(z - x);
    return true;
}

extern "C" {

uint64_t frisk_b200_table_size(int kmin, int kmax) {
    uint64_t s = 0;
    for (int k = kmin; k <= kmax; ++k) s += 1ull << (2 * k);
    return s;
}

int frisk_b200_fasta_scan(const char* text, uint64_t n, uint64_t cap, uint64_t* name_off, uint32_t* name_len,
                          uint64_t* body_off, uint64_t* body_end, uint64_t* seq_len, uint64_t* n_records) {
    if (!n_records || (!text && n)) return FRISK_E_INVALID;
    const unsigned char* t = reinterpret_cast<const unsigned char*>(text);
    uint64_t rec = 0;      // records emitted (a record is emitted when it has a non-empty name, F:153/F:161)
    bool have = false;     // a named record is open
    uint64_t cur_len = 0, cur_name = 0, cur_body = 0;
    uint32_t cur_nlen = 0;
    auto emit = [&](uint64_t end) {
        if (rec < cap) {
            if (name_off) name_off[rec] = cur_name;
            if (name_len) name_len[rec] = cur_nlen;
            if (body_off) body_off[rec] = cur_body;
            if (body_end) body_end[rec] = end;
            if (seq_len) seq_len[rec] = cur_len;
        }
        ++rec;
    };
    uint64_t i = 0;
    while (i < n) {
        // one line [i, j)
        const void* nl = memchr(t + i, '\n', n - i);
        const uint64_t j = nl ? (uint64_t)((const unsigned char*)nl - t) : n;
        uint64_t a = i, b = j;
        while (a < b && is_space(t[a])) ++a;        // line.strip()
        while (b > a && is_space(t[b - 1])) --b;
        if (a < b) {
            if (t[a] == '>') {
                if (have) emit(i);
                if (!frisk_internal::parse_header_name(t, n, i, &cur_name, &cur_nlen))
                    return FRISK_E_FORMAT;          // the reference raises IndexError on an empty header
                cur_body = nl ? j + 1 : n; cur_len = 0;
                have = true;                         // a non-empty name is truthy
            } else if (have) {
                // The reference strips a line only at its ends (F:149): whitespace INSIDE a sequence line stays in its
                // string and counts towards totalLen / nnTotal / window coordinates.  Real FASTA never has it; rather
                // than silently shifting every coordinate after it, such input is refused.
                uint64_t p = a;
                for (; p + 8 <= b; p += 8) {             // eight bytes at a time: is any of them < 0x21 ?  (exact test below)
                    uint64_t x;
                    memcpy(&x, t + p, 8);
                    if (~(((x | 0x8080808080808080ull) - 0x2121212121212121ull) | x) & 0x8080808080808080ull) {
                        for (uint64_t q = p; q < p + 8; ++q)
                            if (is_space(t[q])) return FRISK_E_FORMAT;
                    }
                }
                for (; p < b; ++p)
                    if (is_space(t[p])) return FRISK_E_FORMAT;
                cur_len += b - a;
            }
            // sequence lines before the first header are dropped by the reference (name is None)
        }
        i = nl ? j + 1 : n;
    }
    if (have) emit(n);
    *n_records = rec;
    return rec > cap && (name_off || name_len || body_off || body_end || seq_len) ? FRISK_E_CAPACITY : FRISK_OK;
}

int frisk_b200_pack_layout(const uint64_t* scaf_len, uint64_t n, uint64_t* scaf_off, uint64_t* padded_len) {
    if ((!scaf_len && n) || (!scaf_off && n) || !padded_len) return FRISK_E_INVALID;
    uint64_t pos = 0;
    for (uint64_t i = 0; i < n; ++i) {
        scaf_off[i] = pos;
        pos = align_up(pos + scaf_len[i] + 1, 128);   // >= 1 padding base after every scaffold
    }
    *padded_len = pos + 128;                          // >= 128 trailing padding bases (kernel look-ahead)
    return FRISK_OK;
}

int frisk_b200_pack(const char* src, const uint64_t* src_off, const uint64_t* src_end, const uint64_t* scaf_len,
                    const uint64_t* scaf_off, uint64_t n, uint64_t padded_len, uint32_t* codes, uint32_t* inv,
                    uint32_t* low, uint64_t stats[3], int threads) {
    if (!codes || !inv || !stats || (padded_len & 127) || (n && (!src || !src_off || !src_end || !scaf_len || !scaf_off)))
        return FRISK_E_INVALID;
    stats[0] = stats[1] = stats[2] = 0;
    const uint64_t tail_start = n ? align_up(scaf_off[n - 1] + scaf_len[n - 1] + 1, 128) : 0;
    if (tail_start + 128 > padded_len) return FRISK_E_INVALID;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    if ((uint64_t)threads > n) threads = n ? (int)n : 1;
    std::atomic<uint64_t> next(0);
    std::atomic<int> err(FRISK_OK);
    std::vector<uint64_t> part((size_t)threads * 3, 0);
    auto work = [&](int t) {
        uint64_t* st = &part[(size_t)t * 3];
        for (;;) {
            const uint64_t i = next.fetch_add(1);
            if (i >= n) break;
            const uint64_t nb = (i + 1 < n) ? scaf_off[i + 1] : tail_start;
            const int rc = pack_one(reinterpret_cast<const unsigned char*>(src) + src_off[i],
                                    reinterpret_cast<const unsigned char*>(src) + src_end[i], scaf_len[i], scaf_off[i],
                                    nb, codes, inv, low, st);
            if (rc) err.store(rc);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    // trailing padding
    for (uint64_t pos = tail_start; pos < padded_len; pos += 32) {
        codes[pos >> 4] = 0; codes[(pos >> 4) + 1] = 0;
        inv[pos >> 5] = 0xffffffffu;
        if (low) low[pos >> 5] = 0;
    }
    for (int t = 0; t < threads; ++t)
        for (int k = 0; k < 3; ++k) stats[k] += part[(size_t)t * 3 + k];
    if (err.load()) return err.load();
    if (!low && stats[2]) return FRISK_E_FORMAT;
    return FRISK_OK;
}

}  // extern "C"

namespace {
// str(float) of Python 3 / numpy (repr, shortest round-trip digits; exponent form when the decimal
// point would sit before the 4th leading zero or after the 16th digit; integral values get ".0")
inline void append_pyfloat(std::string& out, double x) {
    if (std::isnan(x)) { out += "nan"; return; }
    if (std::isinf(x)) { out += x < 0 ? "-inf" : "inf"; return; }
    char buf[40];
    const auto r = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);   // [-]d[.ddd]e[+-]XX, shortest
    const char* p = buf;
    if (*p == '-') { out += '-'; ++p; }
    const char* e = p;
    while (e < r.ptr && *e != 'e') ++e;
    char digits[24];
    int nd = 0;
    for (const char* q = p; q < e; ++q)
        if (*q != '.') digits[nd++] = *q;
    int ex = 0;
    {
        const char* q = e + 1;
        const bool neg = *q == '-';
        if (*q == '-' || *q == '+') ++q;
        for (; q < r.ptr; ++q) ex = ex * 10 + (*q - '0');
        if (neg) ex = -ex;
    }
    const int decpt = ex + 1;                             // value = 0.d1d2... * 10^decpt
    if (nd == 1 && digits[0] == '0') { out += "0.0"; return; }
    if (decpt <= -4 || decpt > 16) {                      // exponent notation, at least two exponent digits
        out += digits[0];
        if (nd > 1) { out += '.'; out.append(digits + 1, nd - 1); }
        out += 'e';
        int x10 = decpt - 1;
        out += x10 < 0 ? '-' : '+';
        if (x10 < 0) x10 = -x10;
        char eb[8];
        int ne = 0;
        do { eb[ne++] = (char)('0' + x10 % 10); x10 /= 10; } while (x10);
        if (ne < 2) eb[ne++] = '0';
        while (ne) out += eb[--ne];
    } else if (decpt <= 0) {
        out += "0.";
        out.append((size_t)(-decpt), '0');
        out.append(digits, nd);
    } else if (decpt >= nd) {
        out.append(digits, nd);
        out.append((size_t)(decpt - nd), '0');
        out += ".0";
    } else {
        out.append(digits, decpt);
        out += '.';
        out.append(digits + decpt, nd - decpt);
    }
}

inline void append_int(std::string& out, int64_t v) {
    char buf[24];
    const auto r = std::to_chars(buf, buf + sizeof(buf), v);
    out.append(buf, r.ptr - buf);
}
}  // namespace

extern "C" {

int frisk_b200_format_rows(const char* names, const uint64_t* name_off, const uint32_t* name_len, const uint32_t* row_name,
                           const int64_t* start, const int64_t* stop, const double* rows, uint64_t n_rows, int n_values,
                           char* out, uint64_t cap, uint64_t* n_bytes, int threads) {
    if (!n_bytes || n_values < 0 || n_values > 5 || (n_rows && (!names || !name_off || !name_len || !row_name || !start || !stop || !rows)))
        return FRISK_E_INVALID;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    if ((uint64_t)threads > n_rows / 4096 + 1) threads = (int)(n_rows / 4096 + 1);
    std::vector<std::string> parts((size_t)threads);
    auto work = [&](int t) {
        const uint64_t a = n_rows * (uint64_t)t / (uint64_t)threads, b = n_rows * (uint64_t)(t + 1) / (uint64_t)threads;
        std::string& o = parts[(size_t)t];
        o.reserve((size_t)(b - a) * 96);
        for (uint64_t i = a; i < b; ++i) {
            const uint32_t s = row_name[i];
            o.append(names + name_off[s], name_len[s]);
            o += '\t'; append_int(o, start[i]);
            o += '\t'; append_int(o, stop[i]);
            for (int c = 0; c < n_values; ++c) { o += '\t'; append_pyfloat(o, rows[i * 5 + (uint64_t)c]); }
            o += '\n';
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    uint64_t total = 0;
    for (const auto& p : parts) total += p.size();
    *n_bytes = total;
    if (!out) return FRISK_OK;                            // size query
    if (total > cap) return FRISK_E_CAPACITY;
    uint64_t pos = 0;
    for (const auto& p : parts) { memcpy(out + pos, p.data(), p.size()); pos += p.size(); }
    return FRISK_OK;
}

// ---- 2-state 1-D Gaussian HMM (the fallback of downstream.fit_hmm when hmmlearn is absent) -------------
// Same algorithm and the same order of operations as downstream.GaussianHMM2 (log-space forward/backward,
// Baum-Welch updates, Viterbi), in C so that a million window scores take a fraction of a second.
}  // extern "C"

namespace {
inline double logaddexp2(double a, double b) {
    if (a == b) return a + 0.6931471805599453;                  // numpy: log(2) added when equal (also inf handling)
    const double d = a - b;
    return d > 0 ? a + std::log1p(std::exp(-d)) : b + std::log1p(std::exp(d));
}
inline void log_gauss(const double* x, uint64_t n, const double mean[2], const double var[2], double* logb) {
    const double c0 = std::log(2 * M_PI * var[0]), c1 = std::log(2 * M_PI * var[1]);
    for (uint64_t t = 0; t < n; ++t) {
        const double d0 = x[t] - mean[0], d1 = x[t] - mean[1];
        logb[2 * t] = -0.5 * (c0 + d0 * d0 / var[0]);
        logb[2 * t + 1] = -0.5 * (c1 + d1 * d1 / var[1]);
    }
}
}  // namespace

extern "C" {

int frisk_b200_hmm2_fit(const double* x, uint64_t n, int n_iter, double tol, double min_covar, double start[2],
                        double trans[4], double mean[2], double var[2]) {
    if (!x || n < 2 || !start || !trans || !mean || !var || n_iter < 1) return FRISK_E_INVALID;
    std::vector<double> sorted(x, x + n), logb(2 * n), la(2 * n), lb(2 * n);
    std::sort(sorted.begin(), sorted.end());
    const uint64_t half = n / 2;
    double m0 = 0, m1 = 0, mu = 0;
    for (uint64_t i = 0; i < half; ++i) m0 += sorted[i];
    for (uint64_t i = half; i < n; ++i) m1 += sorted[i];
    mean[0] = m0 / (double)half; mean[1] = m1 / (double)(n - half);
    for (uint64_t i = 0; i < n; ++i) mu += x[i];
    mu /= (double)n;
    double v = 0;
    for (uint64_t i = 0; i < n; ++i) v += (x[i] - mu) * (x[i] - mu);
    var[0] = var[1] = v / (double)n + min_covar;
    start[0] = start[1] = 0.5;
    trans[0] = trans[1] = trans[2] = trans[3] = 0.5;
    double prev = -INFINITY;
    for (int it = 0; it < n_iter; ++it) {
        log_gauss(x, n, mean, var, logb.data());
        const double lt[4] = {std::log(trans[0]), std::log(trans[1]), std::log(trans[2]), std::log(trans[3])};
        la[0] = std::log(start[0]) + logb[0]; la[1] = std::log(start[1]) + logb[1];
        for (uint64_t t = 1; t < n; ++t) {
            la[2 * t] = logb[2 * t] + logaddexp2(la[2 * t - 2] + lt[0], la[2 * t - 1] + lt[2]);
            la[2 * t + 1] = logb[2 * t + 1] + logaddexp2(la[2 * t - 2] + lt[1], la[2 * t - 1] + lt[3]);
        }
        lb[2 * n - 2] = lb[2 * n - 1] = 0.0;
        for (uint64_t t = n - 1; t-- > 0;) {
            const double b0 = logb[2 * t + 2] + lb[2 * t + 2], b1 = logb[2 * t + 3] + lb[2 * t + 3];
            lb[2 * t] = logaddexp2(lt[0] + b0, lt[1] + b1);
            lb[2 * t + 1] = logaddexp2(lt[2] + b0, lt[3] + b1);
        }
        const double ll = logaddexp2(la[2 * n - 2], la[2 * n - 1]);
        double w[2] = {0, 0}, sx[2] = {0, 0}, xi[4] = {0, 0, 0, 0};
        for (uint64_t t = 0; t < n; ++t) {
            const double g0 = std::exp(la[2 * t] + lb[2 * t] - ll), g1 = std::exp(la[2 * t + 1] + lb[2 * t + 1] - ll);
            w[0] += g0; w[1] += g1; sx[0] += g0 * x[t]; sx[1] += g1 * x[t];
            if (t + 1 < n) {
                const double b0 = logb[2 * t + 2] + lb[2 * t + 2], b1 = logb[2 * t + 3] + lb[2 * t + 3];
                xi[0] += std::exp(la[2 * t] + lt[0] + b0 - ll); xi[1] += std::exp(la[2 * t] + lt[1] + b1 - ll);
                xi[2] += std::exp(la[2 * t + 1] + lt[2] + b0 - ll); xi[3] += std::exp(la[2 * t + 1] + lt[3] + b1 - ll);
            }
        }
        const double g00 = std::exp(la[0] + lb[0] - ll), g01 = std::exp(la[1] + lb[1] - ll);
        start[0] = g00 / (g00 + g01); start[1] = g01 / (g00 + g01);
        trans[0] = xi[0] / (xi[0] + xi[1]); trans[1] = xi[1] / (xi[0] + xi[1]);
        trans[2] = xi[2] / (xi[2] + xi[3]); trans[3] = xi[3] / (xi[2] + xi[3]);
        mean[0] = sx[0] / w[0]; mean[1] = sx[1] / w[1];
        double sv[2] = {0, 0};
        for (uint64_t t = 0; t < n; ++t) {
            const double g0 = std::exp(la[2 * t] + lb[2 * t] - ll), g1 = std::exp(la[2 * t + 1] + lb[2 * t + 1] - ll);
            sv[0] += g0 * (x[t] - mean[0]) * (x[t] - mean[0]); sv[1] += g1 * (x[t] - mean[1]) * (x[t] - mean[1]);
        }
        var[0] = sv[0] / w[0] + min_covar; var[1] = sv[1] / w[1] + min_covar;
        if (ll - prev < tol) break;
        prev = ll;
    }
    return FRISK_OK;
}

int frisk_b200_hmm2_viterbi(const double* x, uint64_t n, const double start[2], const double trans[4], const double mean[2],
                            const double var[2], int32_t* path) {
    if (!x || !n || !start || !trans || !mean || !var || !path) return FRISK_E_INVALID;
    std::vector<double> logb(2 * n);
    std::vector<uint8_t> back(2 * n);
    log_gauss(x, n, mean, var, logb.data());
    const double lt[4] = {std::log(trans[0]), std::log(trans[1]), std::log(trans[2]), std::log(trans[3])};
    double d0 = std::log(start[0]) + logb[0], d1 = std::log(start[1]) + logb[1];
    for (uint64_t t = 1; t < n; ++t) {
        const double a0 = d0 + lt[0], a1 = d1 + lt[2], b0 = d0 + lt[1], b1 = d1 + lt[3];
        back[2 * t] = a1 > a0;          // argmax keeps the first maximum, like numpy
        back[2 * t + 1] = b1 > b0;
        d0 = (a1 > a0 ? a1 : a0) + logb[2 * t];
        d1 = (b1 > b0 ? b1 : b0) + logb[2 * t + 1];
    }
    int s = d1 > d0;
    path[n - 1] = s;
    for (uint64_t t = n - 1; t > 0; --t) {
        s = back[2 * t + (uint64_t)s];
        path[t - 1] = s;
    }
    return FRISK_OK;
}

int frisk_b200_plane_sparse(const uint32_t* plane, uint64_t n_words, uint64_t cap, uint32_t* idx, uint32_t* val,
                            uint64_t* n_nonzero) {
    if (!n_nonzero || (!plane && n_words) || n_words > 0xffffffffull) return FRISK_E_INVALID;
    uint64_t n = 0;
    const uint64_t* p64 = reinterpret_cast<const uint64_t*>(plane);     // planes are 16-byte aligned (padded_len % 128 == 0)
    for (uint64_t w = 0; w + 1 < n_words; w += 2) {
        if (p64[w >> 1] == 0) continue;                                  // two words at a time: most of the plane is zero
        for (uint64_t k = w; k < w + 2; ++k) {
            if (!plane[k]) continue;
            if (n < cap && idx && val) { idx[n] = (uint32_t)k; val[n] = plane[k]; }
            ++n;
        }
    }
    if ((n_words & 1) && plane[n_words - 1]) {
        if (n < cap && idx && val) { idx[n] = (uint32_t)(n_words - 1); val[n] = plane[n_words - 1]; }
        ++n;
    }
    *n_nonzero = n;
    return (n > cap && idx && val) ? FRISK_E_CAPACITY : FRISK_OK;
}

int frisk_b200_windows(const uint64_t* scaf_len, const uint64_t* scaf_off, uint64_t n_scaf, int w, int step,
                       int scaffolds_all, uint64_t cap, uint64_t* win_off, uint32_t* win_len, uint32_t* win_scaf,
                       int64_t* win_start, int64_t* win_stop, uint64_t* n_windows) {
    if (!n_windows || w <= 0 || step <= 0 || (n_scaf && (!scaf_len || !scaf_off))) return FRISK_E_INVALID;
    uint64_t n = 0;
    bool too_long = false;
    auto put = [&](uint64_t s, uint64_t off, uint64_t len, int64_t a, int64_t b) {
        if (len > FRISK_B200_MAX_WINDOW) too_long = true;
        if (n < cap) {
            if (win_off) win_off[n] = scaf_off[s] + off;
            if (win_len) win_len[n] = (uint32_t)len;
            if (win_scaf) win_scaf[n] = (uint32_t)s;
            if (win_start) win_start[n] = a;
            if (win_stop) win_stop[n] = b;
        }
        ++n;
    };
    // F:211/F:222: size <= w + ((w * 0.75) - i), evaluated in double like the reference
    const double min_size = (double)w + (((double)w * 0.75) - (double)step);
    for (uint64_t s = 0; s < n_scaf; ++s) {
        const uint64_t size = scaf_len[s];
        if ((double)size <= min_size) {
            if (scaffolds_all && size > 0) put(s, 0, size, 1, (int64_t)size);   // F:219 (30 % rule applied on device)
            continue;
        }
        // Once j + w overshoots it does so for every later j too, so the reference's never-cleared
        // `jumpback` flag (F:232) just means: every remaining j re-emits the last w bases (F:243).
        // the regular part of the grid -- j + w <= size -- without a call per window: plain array fills
        uint64_t j0 = 0;
        if (size >= (uint64_t)w && (uint64_t)w <= FRISK_B200_MAX_WINDOW) {
            // j = 0, step, ... while j + step <= size (the xrange) and j + w <= size (no jump-back yet); none if size < step
            const uint64_t n_reg = size >= (uint64_t)step ? std::min(size - (uint64_t)w, size - (uint64_t)step) / (uint64_t)step + 1 : 0;
            if (n + n_reg <= cap) {
                const uint64_t base = scaf_off[s];
                if (win_off) for (uint64_t k = 0; k < n_reg; ++k) win_off[n + k] = base + k * (uint64_t)step;
                if (win_len) for (uint64_t k = 0; k < n_reg; ++k) win_len[n + k] = (uint32_t)w;
                if (win_scaf) for (uint64_t k = 0; k < n_reg; ++k) win_scaf[n + k] = (uint32_t)s;
                if (win_start) for (uint64_t k = 0; k < n_reg; ++k) win_start[n + k] = (int64_t)(k * (uint64_t)step) + 1;
                if (win_stop) for (uint64_t k = 0; k < n_reg; ++k) win_stop[n + k] = (int64_t)(k * (uint64_t)step + (uint64_t)w);
                n += n_reg;
                j0 = n_reg * (uint64_t)step;
            }
        }
        for (uint64_t j = j0; j + (uint64_t)step <= size; j += (uint64_t)step) {  // xrange(0, size - i + 1, i)
            if (j + (uint64_t)w > size) {                                                          // F:230-232, F:243
                // size < w (possible when 0.75 w < step): seq[size - w : size] has a NEGATIVE start, which Python
                // counts from the end -- the slice is the last min(w - size, size) bases; the coordinates stay
                // (size - w, size).  Signed arithmetic throughout.
                const int64_t lead = (int64_t)size - (int64_t)w;
                if (lead >= 0) put(s, (uint64_t)lead, (uint64_t)w, lead, (int64_t)size);
                else {
                    const uint64_t d = (uint64_t)(-lead);
                    if (d <= size) { if (d > 0) put(s, size - d, d, lead, (int64_t)size); }
                    else put(s, 0, size, lead, (int64_t)size);
                }
            } else put(s, j, w, (int64_t)j + 1, (int64_t)(j + w));                                  // F:245
        }
    }
    *n_windows = n;
    if (too_long) return FRISK_E_UNSUPPORTED;
    const bool wants = win_off || win_len || win_scaf || win_start || win_stop;
    return (n > cap && wants) ? FRISK_E_CAPACITY : FRISK_OK;
}

}  // extern "C"
