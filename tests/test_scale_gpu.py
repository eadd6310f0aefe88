"""BASELINE configs C4 (3 Gbp, 24 long scaffolds, Mbp-scale N runs) and C5 (14 Gbp, 1 M short scaffolds,
many N gaps, --scaffoldsAll) at FULL size.  The oracle cannot run 3-14 Gbp in test time, so the
genomes are drawn directly in packed form on the GPU (torch is only the random source and the
buffer owner) and checked through size-independent properties plus sampled parity:

  * books: strand symmetry; sum of order-k counts = 2 x valid k-words; valid + exMax = every
    kmax-word start; order-1 total = 2 x resolved bases
  * additivity: the background of two halves of the base range sums to the background of the whole
  * sampled parity: the bases of ~300 windows (short scaffolds, N-run neighbours and > 2^32 offsets
    included) are decoded from the planes and scored by the C oracle against the GPU's genome tables
  * one whole small scaffold's background, bit-exact against the C oracle
"""
import ctypes as C

import numpy as np
import pytest

from tests.helpers import assert_rows_close, max_rel_err

pytestmark = pytest.mark.gpu



from frisk_b200.synth_device import LETTERS, build_device_genome, decode, c4_spec, c5_spec  # noqa: E402,F401


def check_books(res, g, kmax=8):
    off, sums = 0, []
    for k in range(1, kmax + 1):
        t = res.tables[off:off + 4 ** k].astype(np.int64)
        idx = np.arange(4 ** k)
        rc = np.zeros_like(idx)
        tmp = idx.copy()
        for _ in range(k):
            rc = (rc << 2) | ((tmp & 3) ^ 1)
            tmp >>= 2
        assert np.array_equal(t, t[rc]), "order %d not reverse-complement symmetric" % k
        sums.append(int(t.sum()))
        off += 4 ** k
    assert all(a >= b for a, b in zip(sums, sums[1:]))
    possible = int(np.maximum(g.scaf_len.astype(np.int64) - kmax + 1, 0).sum())
    assert sums[-1] // 2 + res.meta[1] == possible                      # valid + exMax (F:344)
    assert sums[0] // 2 == g.total_len - g.nn_total                     # every resolved base once per strand


def check_additivity(eng, dg, kmax=8):
    import torch
    P = dg.host.padded_len
    mid = (P // 2) & ~31
    whole = eng.background(dg, kmax)
    parts = eng.background(dg, kmax, first_base=0, last_base=mid)
    eng.background(dg, kmax, d_fwd=parts, first_base=mid, last_base=P - 32)
    assert torch.equal(whole, parts), "background of the halves must sum to the background of the whole"


def check_sampled_windows(eng, dg, res, wins, pick, kmax=8):
    from oracle import c_oracle
    cand = res.win_index[pick]
    seqs = [decode(dg, int(wins.off[c]), int(wins.length[c])) for c in cand]
    lens = np.array([len(s) for s in seqs], np.uint32)
    woff = np.concatenate([[0], np.cumsum(lens[:-1], dtype=np.uint64)]).astype(np.uint64)
    seq = np.ascontiguousarray(np.concatenate(seqs))
    meta = np.array(res.meta, dtype=np.uint64)
    rows, status = c_oracle.score(seq, woff, lens, np.ascontiguousarray(res.tables), meta, 1, kmax, True, threads=8)
    assert np.array_equal(status & 7, res.status[pick] & 7)
    ok = status == 0
    assert ok.sum() > 0.9 * len(pick)
    assert_rows_close(res.rows[pick][ok], rows[ok], rtol_kld=1e-6, rtol_other=1e-15, what="sampled windows")
    assert max_rel_err(res.rows[pick][ok, 0], rows[ok, 0]) < 1e-10


def check_scaffold_background(eng, dg, s, kmax=8):
    """Background of scaffold s alone (its own base range) vs the C oracle on its decoded bases."""
    from oracle import c_oracle
    g = dg.host
    a = int(g.scaf_off[s])
    b = (a + int(g.scaf_len[s]) + 1 + 127) & ~127           # next scaffold's start: the padding closes every word
    d_fwd = eng.background(dg, kmax, first_base=a, last_base=b)
    d_tables, _ = eng.finalize(d_fwd, kmax)
    bases = decode(dg, a, int(g.scaf_len[s]))
    tabs, _ = c_oracle.background(bases, np.array([0, len(bases)], np.uint64), 1, kmax, False, threads=8)
    assert np.array_equal(d_tables.cpu().numpy().view(np.uint64), tabs)


def test_c4_human_scale_3gbp():
    """C4: 3 Gbp in 24 scaffolds of 50-250 Mbp (+2 small ones), AT-rich isochore blocks, one 3 Mbp N
    run per long scaffold; default k = 1..8, w = 5000, step = 2500: ~1.2 M windows on one B200."""
    import torch
    from frisk_b200 import engine as eng
    rng = np.random.Generator(np.random.PCG64(4004))
    lens, runs = c4_spec()
    dg = build_device_genome(eng, lens, runs, seed=44, at_rich_block=300_000 // 16)
    g = dg.host
    assert g.padded_len > 2 ** 31
    pipe = eng.Pipeline(dg)
    pipe.enqueue()
    res = pipe.result()
    wins = pipe.wins
    assert len(wins) > 1_150_000
    assert len(res.rows) == int((pipe.d_status.cpu().numpy().view(np.uint32) & 8 == 0).sum())
    check_books(res, g)
    assert np.all(np.isfinite(res.rows[res.status == 0, 0])) and np.all(res.rows[res.status == 0, 0] >= 0)
    # the N runs remove windows: every candidate fully inside a 3 Mbp run is excluded (F:238)
    assert len(wins) - len(res.rows) >= 24 * (3_000_000 // 2500 - 3)
    check_additivity(eng, dg)
    pick = np.sort(rng.choice(len(res.rows), 260, replace=False))
    near_runs = np.nonzero(res.rows[:, 1] != res.rows[:, 1])[0][:10]         # GC undefined: none expected
    assert near_runs.size == 0
    # windows that straddle an N-run edge (partly unresolved but kept) are the interesting ones
    st = pipe.d_status.cpu().numpy().view(np.uint32)
    kept = np.nonzero((st & 8) == 0)[0]
    edge = np.nonzero(np.diff(kept) > 1)[0][:40]                              # rows just before an excluded stretch
    pick = np.unique(np.concatenate([pick, edge, np.arange(len(res.rows) - 20, len(res.rows))]))
    check_sampled_windows(eng, dg, res, wins, pick)
    check_scaffold_background(eng, dg, 25)
    check_scaffold_background(eng, dg, 24)
    del pipe, dg
    torch.cuda.empty_cache()


def test_c5_wheat_scale_14gbp_fragmented():
    """C5: 14 Gbp in 1,000,000 scaffolds (lognormal, median ~9 kbp, min 500), 30 % with 1-3 N runs of
    10-2,000 bp; --scaffoldsAll so that short scaffolds become variable-length windows (F:211-221)."""
    import torch
    from frisk_b200 import engine as eng
    rng = np.random.Generator(np.random.PCG64(5005))
    n = 1_000_000
    lens, runs = c5_spec(n)
    dg = build_device_genome(eng, lens, runs, seed=55)
    g = dg.host
    assert g.padded_len > 14_000_000_000
    pipe = eng.Pipeline(dg, scaffolds_all=True)
    pipe.enqueue()
    res = pipe.result()
    wins = pipe.wins
    assert len(wins) > 4_500_000 and wins.max_len <= 6250
    check_books(res, g)
    good = res.status == 0
    assert good.mean() > 0.999
    assert np.all(np.isfinite(res.rows[good, 0])) and np.all(res.rows[good, 0] >= 0)
    check_additivity(eng, dg)
    # sample: random rows, rows of the shortest scaffolds, and rows far beyond 2^32 bases
    short = np.nonzero(wins.length[res.win_index] < 1500)[0][:60]
    far = np.nonzero(wins.off[res.win_index] > np.uint64(2 ** 33))[0][-60:]
    pick = np.unique(np.concatenate([rng.choice(len(res.rows), 200, replace=False), short, far]))
    check_sampled_windows(eng, dg, res, wins, pick)
    check_scaffold_background(eng, dg, int(np.argmax(lens)))
    check_scaffold_background(eng, dg, n - 1)
    del pipe, dg
    torch.cuda.empty_cache()


def test_c3_plant_scale_k_sweep_120mbp():
    """C3 at FULL size: 120 Mbp in 12 chromosomes with 30 % TE-like repeats, scored for every kmax' = 1..8
    (eight reference runs) from ONE background pass.  Background bit-exact against the C oracle; per k' the
    order-k' tables are the prefix of the kmax = 8 tables, the books close, and 150 sampled windows agree
    with the oracle scoring them against the GPU's tables."""
    from frisk_b200 import _lib, engine as eng, synth
    from oracle import c_oracle
    sc = synth.make("C3", 1.0)
    g = eng.PackedGenome.from_scaffolds(sc)
    assert g.total_len == 120_000_000
    sweep = eng.run_sweep(g, kmaxes=range(1, 9))
    seq, off = c_oracle.concat(sc)
    tabs8, meta8 = c_oracle.background(seq, off, 1, 8, False, threads=16)
    assert np.array_equal(sweep[8].tables, tabs8) and list(sweep[8].meta) == [int(x) for x in meta8]
    _, woff, wlen, st, sp = c_oracle.crawl(seq, off)
    rng = np.random.Generator(np.random.PCG64(33))
    pick = np.sort(rng.choice(len(woff), 150, replace=False))
    for k, res in sweep.items():
        assert len(res.rows) == len(woff) == 48_000 and np.array_equal(res.coords, np.stack([st, sp], 1))
        tsz = _lib.table_size(1, k)
        assert np.array_equal(res.tables, tabs8[:tsz])                       # x-word counts do not depend on kmax
        possible = int(np.maximum(g.scaf_len.astype(np.int64) - k + 1, 0).sum())
        assert int(res.tables[_lib.table_size(1, k - 1) if k > 1 else 0:].sum()) // 2 + res.meta[1] == possible
        meta = np.array(res.meta, dtype=np.uint64)
        rows, status = c_oracle.score(seq, woff[pick], wlen[pick], np.ascontiguousarray(res.tables), meta, 1, k, True, threads=16)
        assert np.all(status == 0) and np.all(res.status[pick] == 0)
        assert_rows_close(res.rows[pick], rows, rtol_kld=1e-6, rtol_other=1e-15, what="C3 k=%d" % k)
        assert max_rel_err(res.rows[pick, 0], rows[:, 0]) < 1e-10
        assert np.all(np.isfinite(res.rows[:, 0])) and np.all(res.rows[:, 0] >= 0)
    # more context can only sharpen the contrast on repeats: the mean score grows with k'
    means = [float(sweep[k].rows[:, 0].mean()) for k in range(1, 9)]
    assert all(a < b for a, b in zip(means, means[1:]))
