/*
 * CPU oracle (plain C) for frisk's hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A restatement of the algorithm in /root/reference/frisk/__init__.py ("F:" below) with
 * dict lookups replaced by direct array indexing (table index = base-4 number with digits
 * A<T<G<C, first base most significant: exactly the key order of F:253-274).  The loop
 * structure (per order, per position, F:327-351) and the order of every floating-point
 * operation in the scorer (F:394-454, F:466-470) follow the reference, so scores agree with
 * the reference's py3 execution to the last bit or two.
 *
 * PARITY PIN: checked by tests/test_oracle.py against the fixtures in tests/golden/ that
 * tests/golden/make_golden.py produced by executing the reference's own source text
 * (oracle/ref_exec.py).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may load this library.
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC -pthread)
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FO_MAX_K 12

/* A,T,G,C -> 0,1,2,3 (F:70); anything else -> -1.  Upper-case only. */
static inline int base_code(unsigned char c) {
    switch (c) {
        case 'A': return 0;
        case 'T': return 1;
        case 'G': return 2;
        case 'C': return 3;
        default: return -1;
    }
}

static inline unsigned char up(unsigned char c) { return (c >= 'a' && c <= 'z') ? (unsigned char)(c - 32) : c; }

static size_t table_offset(int kmin, int k) { /* offset of order-k table in the concatenation */
    size_t off = 0;
    for (int x = kmin; x < k; ++x) off += (size_t)1 << (2 * x);
    return off;
}

size_t frisk_oracle_table_size(int kmin, int kmax) { return table_offset(kmin, kmax + 1); }

/* F:106-118 countN: number of chars NOT in upper-case ATGC */
static uint64_t count_non_atgc(const unsigned char *s, size_t n) {
    uint64_t bad = 0;
    for (size_t j = 0; j < n; ++j) bad += base_code(s[j]) < 0;
    return bad;
}

/* F:276-278: index of the reverse complement of the k-mer with index idx.
 * complement A<->T, G<->C is code ^ 1 with the 0,1,2,3 = A,T,G,C coding. */
static uint64_t revcomp_index(uint64_t idx, int k) {
    uint64_t r = 0;
    for (int t = 0; t < k; ++t) {
        r = (r << 2) | ((idx & 3) ^ 1);
        idx >>= 2;
    }
    return r;
}

/*
 * F:280-367 computeKmers over ONE sequence, accumulating into tables/meta.
 *   both_strands = genomeMode or sym (F:350); force_upper = 0 only for the genome pass
 *   under --maskHost (F:334-337).  meta = {totalLen, exMax, nnTotal} (F:356-359).
 */
/* Words starting at positions [lo, hi) only; meta[1] (exMax) accumulates, meta[0]/meta[2] untouched. */
static void count_range(const unsigned char *seq, size_t size, size_t lo, size_t hi, int kmin, int kmax,
                        int both_strands, int force_upper, uint64_t *tables, uint64_t meta[3]) {
    for (int k = kmin; k <= kmax; ++k) {   /* F:327 */
        uint64_t *tab = tables + table_offset(kmin, k);
        if (size < (size_t)k) continue;
        size_t end = size - (size_t)k + 1;  /* xrange(size - i + 1), F:329 */
        if (end > hi) end = hi;
        for (size_t j = lo; j < end; ++j) {
            uint64_t idx = 0;
            int ok = 1;
            for (int t = 0; t < k; ++t) {  /* the word seq[j:j+k] (F:331) */
                unsigned char c = seq[j + t];
                if (force_upper) c = up(c); /* F:334-337 */
                int b = base_code(c);
                if (b < 0) { ok = 0; break; }
                idx = (idx << 2) | (uint64_t)b;
            }
            if (!ok) {                      /* F:341-346 */
                if (k == kmax) meta[1] += 1;
                continue;
            }
            tab[idx] += 1;                  /* F:348 */
            if (both_strands) tab[revcomp_index(idx, k)] += 1; /* F:350-351 */
        }
    }
}

void frisk_oracle_count(const unsigned char *seq, size_t size, int kmin, int kmax, int both_strands,
                        int force_upper, uint64_t *tables, uint64_t meta[3]) {
    meta[0] += size;                       /* F:323 */
    meta[2] += count_non_atgc(seq, size);  /* F:324-325 */
    count_range(seq, size, 0, size, kmin, kmax, both_strands, force_upper, tables, meta);
}

/* error bits in a row's status */
#define FO_ERR_KLD_ZERODIV 1u   /* ZeroDivisionError inside IvomBuild (F:401-437, F:454) */
#define FO_ERR_GC_ZERODIV 2u    /* ZeroDivisionError in calcGC (F:136) */
#define FO_ERR_LOG_DOMAIN 4u    /* ValueError from math.log of a non-positive ratio (F:470) */

/*
 * F:369-457 IvomBuild for all kmax-mers present in the window; raw (un-normalised) values
 * are written to out[] in table order for the nonzero window bins, count returned.
 * tabs/space are the window's or the genome's (isGenomeIVOM).  *err gets FO_ERR_* bits.
 */
static size_t ivom_raw(const uint64_t *wtab_kmax, const uint64_t *tabs, int64_t space, int kmin, int kmax,
                       double *out, double *total_out, unsigned *err) {
    size_t nk = (size_t)1 << (2 * kmax);
    size_t offs[FO_MAX_K + 2];
    for (int x = kmin; x <= kmax; ++x) offs[x] = table_offset(kmin, x);
    size_t n = 0;
    double total = 0; /* F:382 (int 0 + float) */
    for (size_t kappa = 0; kappa < nk; ++kappa) { /* F:386, dict order == index order */
        if (wtab_kmax[kappa] == 0) continue;      /* F:388 */
        double ivom = 0.0;
        uint64_t running = 0;
        for (int x = kmin; x <= kmax; ++x) {
            uint64_t c = tabs[offs[x] + (kappa >> (2 * (kmax - x)))]; /* count of prefix k[0:x] */
            uint64_t weight = c << (2 * x);                           /* c * 4**x */
            int64_t den = (space - (int64_t)(x - 1)) * 2;
            if (den == 0) { *err |= FO_ERR_KLD_ZERODIV; return 0; }
            double prob = (double)c / (double)den;
            running += weight;                                        /* F:426-432 */
            if (running == 0) { *err |= FO_ERR_KLD_ZERODIV; return 0; }
            double alpha = (double)weight / (double)running;         /* F:437 */
            if (x == kmin) ivom = alpha * prob;                       /* F:442 */
            else ivom = alpha * prob + ((1 - alpha) * ivom);          /* F:444-446 */
        }
        out[n++] = ivom;
        total += ivom;                                                /* F:450 */
    }
    *total_out = total;
    return n;
}

/*
 * One window: F:1480-1488.  gtabs = genome tables (orders kmin..kmax), gmeta = {totalLen, exMax,
 * nnTotal}.  out = {KLD, GC, PI, SI, CRI}.  wtabs_out (nullable) receives the window tables and
 * wmeta_out (nullable) the window's meta.  Returns the FO_ERR_* status.
 */
static unsigned window_ws(const unsigned char *win, size_t len, int kmin, int kmax, const uint64_t *gtabs,
                          const uint64_t gmeta[3], int rip, double out[5], uint64_t *wtabs_out, uint64_t wmeta_out[3],
                          uint64_t *wt, double *gi, double *wi) {
    /* wt/gi/wi: caller-provided scratch (tsz u64, 4^kmax doubles x2) so worker threads do not
     * fight over the allocator; wt is re-zeroed here like the reference's deepcopy of the blank map (F:317) */
    size_t tsz = frisk_oracle_table_size(kmin, kmax);
    memset(wt, 0, tsz * sizeof(uint64_t));
    uint64_t wmeta[3] = {0, 0, 0};
    unsigned err = 0;
    frisk_oracle_count(win, len, kmin, kmax, 0, 1, wt, wmeta); /* F:1480 */
    int64_t gspace = (int64_t)gmeta[0] - (int64_t)gmeta[2];   /* F:379 */
    int64_t wspace = (int64_t)wmeta[0] - (int64_t)wmeta[2];   /* F:380 */
    const uint64_t *wmax = wt + table_offset(kmin, kmax);
    double gsum = 0, wsum = 0;
    size_t n = ivom_raw(wmax, gtabs, gspace, kmin, kmax, gi, &gsum, &err); /* F:1481 */
    size_t n2 = err ? 0 : ivom_raw(wmax, wt, wspace, kmin, kmax, wi, &wsum, &err); /* F:1482 */
    double kld = 0; /* F:465 */
    if (!err && n) {
        if (gsum == 0 || wsum == 0) err |= FO_ERR_KLD_ZERODIV; /* F:454 */
        for (size_t t = 0; t < n && !err; ++t) {
            double w = wi[t] / wsum; /* F:454 */
            double g = gi[t] / gsum;
            if (g != 0) {            /* F:469 */
                double ratio = w / g;
                if (!(ratio > 0)) { err |= FO_ERR_LOG_DOMAIN; break; }
                kld += w * (log(ratio) / log(2.0)); /* F:470: math.log(x, 2) == log(x)/log(2) */
            }
        }
    }
    (void)n2;
    out[0] = kld;
    /* F:120-137 calcGC on the raw (case-sensitive) window */
    uint64_t gc = 0, at = 0;
    for (size_t j = 0; j < len; ++j) {
        int b = base_code(win[j]);
        if (b >= 2) gc++;
        else if (b >= 0) at++;
    }
    if (gc + at == 0) { err |= FO_ERR_GC_ZERODIV; out[1] = NAN; }
    else out[1] = (double)gc / (double)(gc + at);
    /* F:474-495 calcRIP from the dinucleotide table */
    out[2] = out[3] = out[4] = NAN;
    if (rip && kmin <= 2 && kmax >= 2) {
        const uint64_t *di = wt + table_offset(kmin, 2);
        /* index = 4*first + second with A,T,G,C = 0,1,2,3 */
        uint64_t AT = di[1], TA = di[4], AC = di[3], GT = di[9], CA = di[12], TG = di[6];
        double pi = AT > 0 ? (double)TA / (double)AT : NAN;
        double si = (AC + GT) > 0 ? (double)(CA + TG) / (double)(AC + GT) : NAN;
        /* F:491 "if PI and SI": 0.0 is falsy, NaN is truthy */
        double cri = (pi != 0.0 && si != 0.0) ? pi - si : NAN;
        out[2] = pi; out[3] = si; out[4] = cri;
    }
    if (wtabs_out) memcpy(wtabs_out, wt, tsz * sizeof(uint64_t));
    if (wmeta_out) memcpy(wmeta_out, wmeta, sizeof(wmeta));
    return err;
}

unsigned frisk_oracle_window(const unsigned char *win, size_t len, int kmin, int kmax, const uint64_t *gtabs,
                             const uint64_t gmeta[3], int rip, double out[5], uint64_t *wtabs_out,
                             uint64_t wmeta_out[3]) {
    size_t tsz = frisk_oracle_table_size(kmin, kmax);
    size_t nk = (size_t)1 << (2 * kmax);
    uint64_t *wt = (uint64_t *)malloc(tsz * sizeof(uint64_t));
    double *gi = (double *)malloc(nk * sizeof(double));
    double *wi = (double *)malloc(nk * sizeof(double));
    unsigned err = window_ws(win, len, kmin, kmax, gtabs, gmeta, rip, out, wtabs_out, wmeta_out, wt, gi, wi);
    free(wt); free(gi); free(wi);
    return err;
}

/*
 * F:194-251 crawlGenome for ONE scaffold of length `size` at `seq`.
 * Emits (offset within scaffold, length, start, stop) per yielded window into the out arrays
 * (capacity cap); returns the number of windows (may exceed cap: call again with more room).
 */
size_t frisk_oracle_crawl(const unsigned char *seq, size_t size, int w, int step, int scaffolds_all,
                          uint64_t *off, uint32_t *len, int64_t *start, int64_t *stop, size_t cap) {
    size_t n = 0;
    int jumped = 0;
    int small = (double)size <= (double)w + (((double)w * 0.75) - (double)step); /* F:211, F:222 */
    if (small && scaffolds_all) {
        if ((double)count_non_atgc(seq, size) >= 0.3 * (double)size) return 0; /* F:213 */
        if (n < cap) { off[n] = 0; len[n] = (uint32_t)size; start[n] = 1; stop[n] = (int64_t)size; }
        return 1;
    }
    if (small) return 0;
    for (size_t j = 0; j + step <= size; j += step) { /* xrange(0, size - i + 1, i), F:228 */
        size_t o, l = (size_t)w;
        if (j + w > size) { /* F:230-232; size < w: Python's negative slice start counts from the end */
            jumped = 1;
            if (size >= (size_t)w) o = size - w;
            else if ((size_t)w - size <= size) { l = (size_t)w - size; o = size - l; }
            else { o = 0; l = size; }
        } else o = j;
        if ((double)count_non_atgc(seq + o, l) >= 0.3 * (double)l) continue; /* F:237-241 */
        if (n < cap) {
            off[n] = o; len[n] = (uint32_t)l;
            if (jumped) { start[n] = (int64_t)size - (int64_t)w; stop[n] = (int64_t)size; } /* F:243 */
            else { start[n] = (int64_t)j + 1; stop[n] = (int64_t)(j + w); }        /* F:245 */
        }
        ++n;
    }
    return n;
}

/* ------------------------------------------------------------------ threaded whole-path driver */
typedef struct {
    const unsigned char *seq;
    const uint64_t *scaf_off; /* nscaf + 1 offsets into seq */
    size_t nscaf;
    int kmin, kmax, mask_host, t, threads;
    uint64_t *tables; /* private */
    uint64_t meta[3];
} bg_job;

static void *bg_worker(void *p) {
    bg_job *j = (bg_job *)p;
    for (size_t s = 0; s < j->nscaf; ++s) { /* thread t takes the t-th slice of every scaffold */
        size_t size = (size_t)(j->scaf_off[s + 1] - j->scaf_off[s]);
        size_t lo = size * (size_t)j->t / (size_t)j->threads, hi = size * (size_t)(j->t + 1) / (size_t)j->threads;
        count_range(j->seq + j->scaf_off[s], size, lo, hi, j->kmin, j->kmax, 1, !j->mask_host, j->tables, j->meta);
    }
    return NULL;
}

/* Background tables of a whole genome (F:1442); word start positions split over `threads` workers. */
void frisk_oracle_background(const unsigned char *seq, const uint64_t *scaf_off, size_t nscaf, int kmin, int kmax,
                             int mask_host, int threads, uint64_t *tables, uint64_t meta[3]) {
    size_t tsz = frisk_oracle_table_size(kmin, kmax);
    if (threads < 1) threads = 1;
    bg_job *jobs = (bg_job *)calloc(threads, sizeof(bg_job));
    pthread_t *tid = (pthread_t *)calloc(threads, sizeof(pthread_t));
    for (int t = 0; t < threads; ++t) {
        jobs[t].seq = seq; jobs[t].scaf_off = scaf_off; jobs[t].nscaf = nscaf;
        jobs[t].kmin = kmin; jobs[t].kmax = kmax; jobs[t].mask_host = mask_host;
        jobs[t].t = t; jobs[t].threads = threads;
        jobs[t].tables = (uint64_t *)calloc(tsz, sizeof(uint64_t));
        pthread_create(&tid[t], NULL, bg_worker, &jobs[t]);
    }
    for (size_t s = 0; s < nscaf; ++s) {
        size_t size = (size_t)(scaf_off[s + 1] - scaf_off[s]);
        meta[0] += size;                                        /* F:323 */
        meta[2] += count_non_atgc(seq + scaf_off[s], size);     /* F:324-325 */
    }
    for (int t = 0; t < threads; ++t) {
        pthread_join(tid[t], NULL);
        for (size_t b = 0; b < tsz; ++b) tables[b] += jobs[t].tables[b];
        meta[1] += jobs[t].meta[1];
        free(jobs[t].tables);
    }
    free(jobs); free(tid);
}

typedef struct {
    const unsigned char *seq;
    const uint64_t *win_off; /* absolute offsets into seq */
    const uint32_t *win_len;
    size_t first, last;
    int kmin, kmax, rip;
    const uint64_t *gtabs;
    const uint64_t *gmeta;
    double *rows;      /* n x 5 */
    uint32_t *status;  /* n */
} win_job;

static void *win_worker(void *p) {
    win_job *j = (win_job *)p;
    size_t tsz = frisk_oracle_table_size(j->kmin, j->kmax);
    size_t nk = (size_t)1 << (2 * j->kmax);
    uint64_t *wt = (uint64_t *)malloc(tsz * sizeof(uint64_t));
    double *gi = (double *)malloc(nk * sizeof(double));
    double *wi = (double *)malloc(nk * sizeof(double));
    for (size_t i = j->first; i < j->last; ++i)
        j->status[i] = window_ws(j->seq + j->win_off[i], j->win_len[i], j->kmin, j->kmax, j->gtabs, j->gmeta, j->rip,
                                 j->rows + 5 * i, NULL, NULL, wt, gi, wi);
    free(wt); free(gi); free(wi);
    return NULL;
}

/* Score n windows (F:1478-1494), split over `threads` workers. */
void frisk_oracle_score(const unsigned char *seq, const uint64_t *win_off, const uint32_t *win_len, size_t n, int kmin,
                        int kmax, int rip, const uint64_t *gtabs, const uint64_t gmeta[3], int threads, double *rows,
                        uint32_t *status) {
    if (threads < 1) threads = 1;
    win_job *jobs = (win_job *)calloc(threads, sizeof(win_job));
    pthread_t *tid = (pthread_t *)calloc(threads, sizeof(pthread_t));
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (win_job){seq, win_off, win_len, n * (size_t)t / threads, n * (size_t)(t + 1) / threads,
                            kmin, kmax, rip, gtabs, gmeta, rows, status};
        pthread_create(&tid[t], NULL, win_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
    free(jobs); free(tid);
}
