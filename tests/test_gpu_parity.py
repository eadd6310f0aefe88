"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(libfrisk_b200.so); the oracle (oracle/) and the golden fixtures are only the checker."""
import numpy as np
import pytest

from tests.helpers import Golden, SMALL_CASES, assert_rows_close, max_rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from frisk_b200 import _lib, engine
    _lib.require_device()          # fail loudly: no CPU fallback exists
    return engine


def _run_case(eng, g, host_path=False, dump=False):
    q = eng.PackedGenome.from_scaffolds(g.scaffolds(), pinned=host_path)
    h = eng.PackedGenome.from_scaffolds(g.host(), pinned=host_path) if g.host() is not None else None
    kw = g.kwargs()
    if host_path:
        return eng.run_host(q, h, **kw)
    return eng.run(q, h, dump=dump, **kw)


def _check_against_golden(res, g, what):
    from frisk_b200 import _lib
    assert np.array_equal(res.tables, g.tables), what + ": genome tables must be bit-exact"
    assert list(res.meta) == [int(x) for x in g.genome_meta], what + ": totalLen/exMax/nnTotal"
    assert res.names == g.names, what
    assert np.array_equal(res.coords, g.coords), what
    ref = g.vals.copy()
    zd = np.isinf(ref[:, 0])          # rows on which the reference raised ZeroDivisionError
    assert np.array_equal(zd, (res.status & _lib.ROW_KLD_ZERODIV) != 0), what + ": ZeroDivisionError rows"
    rows = res.rows.copy()
    ref[zd, 0] = 0.0
    rows[zd, 0] = 0.0
    # north_star tolerance: 1e-6 relative on KLD; GC/PI/SI/CRI are single divisions -> exact
    assert_rows_close(rows, ref, rtol_kld=1e-6, rtol_other=1e-15, what=what)
    return max_rel_err(rows[:, 0], ref[:, 0])


@pytest.mark.parametrize("case", SMALL_CASES)
def test_staged_path_matches_reference_golden(eng, case):
    g = Golden(case)
    err = _check_against_golden(_run_case(eng, g), g, case)
    assert err < 1e-10, "KLD agreement is expected ~1e-13, got %.2e" % err


@pytest.mark.parametrize("case", ["edge_default", "edge_scaffoldsAll", "edge_maskHost", "c2_small_query_vs_c1_host",
                                  "edge_k2_5_w1000_i250", "edge_k9_10_w2500_i2500"])
def test_host_buffer_path_matches_reference_golden(eng, case):
    g = Golden(case)
    res = _run_case(eng, g, host_path=True)
    _check_against_golden(res, g, case + "[run_host]")
    staged = _run_case(eng, g)
    assert np.array_equal(res.rows, staged.rows, equal_nan=True), "both entry paths run the same kernels"
    # the invalid plane uploaded dense / as its non-zero words: same bits
    q = eng.PackedGenome.from_scaffolds(g.scaffolds(), pinned=True)
    h = eng.PackedGenome.from_scaffolds(g.host(), pinned=True) if g.host() is not None else None
    for sparse in (False, True):
        alt = eng.run_host(q, h, sparse=sparse, **g.kwargs())
        assert np.array_equal(alt.rows, res.rows, equal_nan=True) and np.array_equal(alt.tables, res.tables), sparse
        assert alt.meta == res.meta and np.array_equal(alt.status, res.status)


@pytest.mark.parametrize("case", ["edge_default", "edge_k2_5_w1000_i250", "edge_scaffoldsAll", "edge_k1_1", "edge_k1_9"])
def test_window_tables_bit_exact(eng, case):
    """Window k-mer tables (debug dump of the shared-memory histograms) vs the reference's."""
    from oracle import c_oracle
    g = Golden(case)
    res = _run_case(eng, g, dump=True)
    kw = g.kwargs()
    for slot, idx in enumerate(g.win_pick):
        assert np.array_equal(res.win_tables[idx].astype(np.uint32), g.win_tables[slot]), (case, int(idx))
    # and every window against the C oracle
    sc = g.scaffolds()
    seq, off = c_oracle.concat(sc)
    _, woff, wlen, _, _ = c_oracle.crawl(seq, off, kw["w"], kw["step"], kw["scaffolds_all"])
    assert len(woff) == len(res.rows)
    for i in range(len(woff)):
        win = seq[int(woff[i]):int(woff[i]) + int(wlen[i])]
        _, _, wt, _ = c_oracle.window_tables(win, g.tables, g.genome_meta, kw["kmin"], kw["kmax"])
        assert np.array_equal(res.win_tables[i].astype(np.uint64), wt), (case, i)


@pytest.mark.parametrize("case", ["edge_default", "edge_scaffoldsAll", "edge_k4_8", "c1_small", "edge_k2_5_w1000_i250"])
def test_dense_table_kernel_matches_reference_golden(eng, case):
    """frisk_b200_score has two kernels: the bucketed one (default for 4 <= kmax <= 8 and windows
    <= 8192) and the dense-table one (everything else).  Force the dense one on the same cases."""
    from frisk_b200 import _lib
    g = Golden(case)
    _lib.check(_lib.lib().frisk_b200_set_option(b"force_dense_kernel", 1), "set_option")
    try:
        res = _run_case(eng, g, dump=True)
    finally:
        _lib.lib().frisk_b200_set_option(b"force_dense_kernel", 0)
    _check_against_golden(res, g, case + "[dense]")
    default = _run_case(eng, g, dump=True)
    assert np.array_equal(res.win_tables, default.win_tables)
    assert max_rel_err(res.rows[:, 0], default.rows[:, 0]) < 1e-12


@pytest.mark.parametrize("case", ["edge_default", "edge_scaffoldsAll", "edge_maskHost", "edge_k2_5_w1000_i250", "edge_k1_1",
                                  "c2_small_query_vs_c1_host"])
def test_general_kernel_matches_reference_golden(eng, case):
    """The general (global-memory, run-time K) score kernel serves kmax 9..12 and windows longer than
    65,535 bases; forced here on the default-range cases it must agree with the reference and with the
    shared-memory kernels, and be bit-reproducible."""
    from frisk_b200 import _lib
    g = Golden(case)
    _lib.check(_lib.lib().frisk_b200_set_option(b"force_general_kernel", 1), "set_option")
    try:
        res = _run_case(eng, g, dump=True)
        again = _run_case(eng, g)
    finally:
        _lib.lib().frisk_b200_set_option(b"force_general_kernel", 0)
    _check_against_golden(res, g, case + "[general]")
    assert np.array_equal(res.rows, again.rows, equal_nan=True)
    default = _run_case(eng, g, dump=True)
    assert np.array_equal(res.win_tables, default.win_tables)
    assert np.array_equal(res.status, default.status)
    ok = res.status == 0
    assert max_rel_err(res.rows[ok, 0], default.rows[ok, 0]) < 1e-12


@pytest.mark.parametrize("kw", [dict(kmin=1, kmax=10), dict(kmin=3, kmax=12, w=3000, step=1500, scaffolds_all=True),
                                dict(kmin=11, kmax=11, mask_host=True)])
def test_word_sizes_beyond_8_against_c_oracle(eng, kw):
    """--maxWordSize 9..12 (the reference takes any; its default is 8): background, finalise, genome IVOM
    and scoring all run in the general path.  Tables bit-exact, rows to 1e-10."""
    from frisk_b200 import synth
    from oracle import c_oracle
    sc = synth.make("edge") + synth.make("C2", 0.004, seed=9)
    full = dict(kmin=1, kmax=8, w=5000, step=2500, mask_host=False, scaffolds_all=False, rip=True)
    full.update(kw)
    ref = c_oracle.run(sc, threads=8, **full)
    g = eng.PackedGenome.from_scaffolds(sc, pinned=True)
    for res in (eng.run(g, **full), eng.run_host(g, **full)):
        assert np.array_equal(res.tables, ref["tables"])
        assert list(res.meta) == [int(x) for x in ref["meta"]]
        assert res.names == ref["names"] and np.array_equal(res.coords, ref["coords"])
        assert np.array_equal(res.status & 7, ref["status"] & 7)
        ok = ref["status"] == 0
        assert_rows_close(res.rows[ok], ref["rows"][ok], rtol_kld=1e-6, rtol_other=1e-15, what=str(kw))
        assert max_rel_err(res.rows[ok, 0], ref["rows"][ok, 0]) < 1e-10


def test_windows_longer_than_65535_bases(eng):
    """--windowlen beyond the 16-bit counters of the shared-memory kernels, default k = 1..8 and a
    narrow k = 2..5 (tables entirely in shared memory), with N runs and soft-masked bases inside."""
    from frisk_b200 import synth
    from oracle import c_oracle
    sc = synth.make("C1", 0.2, seed=31) + synth.make("edge")
    name, seq = sc[0]
    seq = seq.copy()
    seq[200_000:230_000] = ord("N")
    seq[400_000:420_000] |= 0x20                       # lower case
    sc[0] = (name, seq)
    for kw in (dict(kmin=1, kmax=8, w=100_000, step=40_000), dict(kmin=2, kmax=5, w=70_001, step=70_001, scaffolds_all=True)):
        full = dict(mask_host=False, scaffolds_all=False, rip=True)
        full.update(kw)
        ref = c_oracle.run(sc, threads=8, **full)
        res = eng.run(eng.PackedGenome.from_scaffolds(sc), **full)
        assert len(res.rows) == len(ref["rows"]) > 0
        assert np.array_equal(res.tables, ref["tables"])
        assert np.array_equal(res.coords, ref["coords"])
        assert np.array_equal(res.status & 7, ref["status"] & 7)
        ok = ref["status"] == 0
        assert_rows_close(res.rows[ok], ref["rows"][ok], rtol_kld=1e-6, rtol_other=1e-15, what="long windows " + str(kw))
        assert max_rel_err(res.rows[ok, 0], ref["rows"][ok, 0]) < 1e-10


@pytest.mark.parametrize("case", ["edge_k2_5_w1000_i250", "edge_k1_3_w3000_i1000", "edge_k1_1"])
def test_small_word_sizes_three_kernels_agree(eng, case):
    """kmax <= 6 is served by the small-K kernel (no sorting); the bucketed kernel (kmax 4..6, forced) and
    the dense-table kernel (forced) must give the same window tables and the same rows to 1e-12."""
    from frisk_b200 import _lib
    g = Golden(case)
    default = _run_case(eng, g, dump=True)
    _check_against_golden(default, g, case + "[small-K]")
    again = _run_case(eng, g)
    assert np.array_equal(default.rows, again.rows, equal_nan=True)             # bit-reproducible
    for opt in (b"force_bucket_kernel", b"force_dense_kernel"):
        _lib.check(_lib.lib().frisk_b200_set_option(opt, 1), "set_option")
        try:
            other = _run_case(eng, g, dump=True)
        finally:
            _lib.lib().frisk_b200_set_option(opt, 0)
        _check_against_golden(other, g, case + "[%s]" % opt.decode())
        assert np.array_equal(other.win_tables, default.win_tables)
        ok = default.status == 0
        assert np.array_equal(other.status, default.status)
        assert max_rel_err(other.rows[ok, 0], default.rows[ok, 0]) < 1e-12


def _with_option(opt, value, fn):
    from frisk_b200 import _lib
    _lib.check(_lib.lib().frisk_b200_set_option(opt, value), "set_option")
    try:
        return fn()
    finally:
        _lib.lib().frisk_b200_set_option(opt, 0)


@pytest.mark.parametrize("case", ["edge_default", "edge_scaffoldsAll", "edge_maskHost", "edge_k4_8", "c1_small",
                                  "c2_small_query_vs_c1_host"])
def test_direct_kernel_agrees_with_bucket_and_dense_kernels(eng, case):
    """kmax 7 and 8 run on the direct kernel (byte table, every position scores its own K-mer with weight
    1/count; frisk_direct.cu).  It must reproduce the reference goldens, be bit-reproducible, and agree with the
    bucketed and the dense-table kernels (forced): identical window tables and status, rows to 1e-12; the
    384-thread instantiation gives the same."""
    g = Golden(case)
    default = _run_case(eng, g, dump=True)
    err = _check_against_golden(default, g, case + "[direct]")
    assert err < 1e-10
    again = _run_case(eng, g)
    assert np.array_equal(default.rows, again.rows, equal_nan=True)
    ok = default.status == 0
    for opt, val in ((b"force_bucket_kernel", 1), (b"force_dense_kernel", 1), (b"force_direct_kernel", 1)):
        other = _with_option(opt, val, lambda: _run_case(eng, g, dump=True))
        _check_against_golden(other, g, case + "[%s]" % opt.decode())
        assert np.array_equal(other.win_tables, default.win_tables), opt
        assert np.array_equal(other.status, default.status), opt
        assert np.array_equal(np.isnan(other.rows), np.isnan(default.rows)), opt
        assert max_rel_err(other.rows[ok, 0], default.rows[ok, 0]) < 1e-12, opt
        assert np.array_equal(other.rows[:, 1:], default.rows[:, 1:], equal_nan=True), opt


@pytest.mark.parametrize("kw", [dict(kmax=7), dict(kmax=7, kmin=3, w=2000, step=700, scaffolds_all=True),
                                dict(kmax=8, kmin=6), dict(kmax=8, kmin=8, w=7000, step=3000), dict(kmax=7, kmin=7, mask_host=True)])
def test_direct_kernel_word_ranges_against_c_oracle(eng, kw):
    """kmax 7 (16 KiB byte table, order-4 atomics) and kmin > 1 on the direct kernel: window tables bit-exact
    against the C oracle for every window, rows to 1e-10, bucketed kernel (forced) to 1e-12."""
    from frisk_b200 import synth
    from oracle import c_oracle
    sc = synth.make("edge") + synth.make("C2", 0.004, seed=5)
    full = dict(kmin=1, kmax=8, w=5000, step=2500, mask_host=False, scaffolds_all=False, rip=True)
    full.update(kw)
    ref = c_oracle.run(sc, threads=8, **full)
    g = eng.PackedGenome.from_scaffolds(sc)
    res = eng.run(g, dump=True, **full)
    assert np.array_equal(res.tables, ref["tables"])
    assert np.array_equal(res.coords, ref["coords"])
    assert np.array_equal(res.status & 7, ref["status"] & 7)
    ok = ref["status"] == 0
    assert_rows_close(res.rows[ok], ref["rows"][ok], rtol_kld=1e-6, rtol_other=1e-15, what=str(kw))
    assert max_rel_err(res.rows[ok, 0], ref["rows"][ok, 0]) < 1e-10
    seq, off = c_oracle.concat(sc)
    for i in range(len(ref["rows"])):
        if ref["status"][i] & 8:
            continue
        win = seq[int(ref["win_off"][i]):int(ref["win_off"][i]) + int(ref["win_len"][i])]
        _, _, wt, _ = c_oracle.window_tables(win, ref["tables"], ref["meta"], full["kmin"], full["kmax"])
        assert np.array_equal(res.win_tables[i].astype(np.uint64), wt), (kw, i)
    for opt in (b"force_bucket_kernel", b"force_direct_kernel", b"force_nibble_kernel"):
        other = _with_option(opt, 1, lambda: eng.run(g, dump=True, **full))
        assert np.array_equal(other.win_tables, res.win_tables), opt
        assert np.array_equal(other.status, res.status), opt
        assert max_rel_err(other.rows[ok, 0], res.rows[ok, 0]) < 1e-12, opt


@pytest.mark.parametrize("opt", [None, b"force_direct_kernel", b"force_nibble_kernel"])
def test_direct_kernel_hands_over_what_a_byte_cannot_hold(eng, opt):
    """Windows with a K-mer seen 256+ times (byte wrap; the nibble kernel: 16+ times, detected through the grand
    total of its counters) or with more than 64 words cut short at K-1 / K-2 bases (many N boundaries) are marked
    for the bucketed kernel and re-done there: rows and tables still exact, and no internal marker survives in the
    status words."""
    from frisk_b200 import synth
    from oracle import c_oracle
    rng = np.random.Generator(np.random.PCG64(8))
    a = synth.iid_bases(rng, 60_000, 0.45)
    a[2_000:2_300] = ord("A")                                  # 293 x AAAAAAAA: just over a byte
    a[7_600:7_860] = ord("C")                                  # 253 x CCCCCCCC: just under
    a[12_000:12_600] = np.tile(np.frombuffer(b"ACG", dtype=np.uint8), 200)
    for p in range(20_000, 24_000, 40):                        # an N every 40 bases: 100 boundaries per window
        a[p] = ord("N")
    for p in range(30_000, 32_400, 80):                        # 30 boundaries: stays on the direct kernel (60 side words)
        a[p] = ord("N")
    a[40_000:40_255 + 7] = ord("T")                            # exactly 255 x TTTTTTTT
    a[45_000:45_256 + 7] = ord("G")                            # exactly 256 x GGGGGGGG
    a[50_000:50_015 + 7] = ord("A")                            # exactly 15 x AAAAAAAA: the most a nibble holds
    a[52_000:52_016 + 7] = ord("T")                            # exactly 16: wraps a nibble (carry into the neighbour)
    a[54_000:54_008] = np.frombuffer(b"CCCCCCCC", dtype=np.uint8)
    a[56_000:56_000 + 17 * 9] = np.tile(np.frombuffer(b"CCCCCCCCA", dtype=np.uint8), 17)   # 17 x the TOP nibble of its word: carry lost
    sc = [("handover", a)]
    for kw in (dict(), dict(kmax=7), dict(w=3000, step=1000, kmin=2)):
        full = dict(kmin=1, kmax=8, w=5000, step=2500, mask_host=False, scaffolds_all=False, rip=True)
        full.update(kw)
        ref = c_oracle.run(sc, threads=4, **full)
        run = lambda **k: eng.run(eng.PackedGenome.from_scaffolds(sc), **k, **full)
        res = run(dump=True) if opt is None else _with_option(opt, 1, lambda: run(dump=True))
        assert np.array_equal(res.tables, ref["tables"])
        assert np.array_equal(res.status & 7, ref["status"] & 7)
        assert not np.any(res.status & 0x80000000), "internal redo marker must not survive"
        ok = ref["status"] == 0
        assert_rows_close(res.rows[ok], ref["rows"][ok], rtol_kld=1e-6, rtol_other=1e-15, what="handover")
        assert max_rel_err(res.rows[ok, 0], ref["rows"][ok, 0]) < 1e-10
        seq, off = c_oracle.concat(sc)
        for i in range(len(ref["rows"])):
            win = seq[int(ref["win_off"][i]):int(ref["win_off"][i]) + int(ref["win_len"][i])]
            _, _, wt, _ = c_oracle.window_tables(win, ref["tables"], ref["meta"], full["kmin"], full["kmax"])
            assert np.array_equal(res.win_tables[i].astype(np.uint64), wt), (kw, i)
        nodump = run() if opt is None else _with_option(opt, 1, run)
        assert np.array_equal(nodump.rows, res.rows, equal_nan=True)
        # one C call from pinned host buffers: rows and status are written over PCIe, the hand-over marks stay on the device
        hostrun = lambda: eng.run_host(eng.PackedGenome.from_scaffolds(sc, pinned=True), **full)
        viahost = hostrun() if opt is None else _with_option(opt, 1, hostrun)
        assert np.array_equal(viahost.rows, res.rows, equal_nan=True) and np.array_equal(viahost.status, res.status)


def test_long_windows_use_segments(eng):
    """Windows longer than the bucketed kernel's 8192-base buffer (and longer than the dense
    kernel's 8192-entry k-mer list) against the C oracle."""
    from frisk_b200 import synth
    from oracle import c_oracle
    sc = synth.make("C1", 0.06, seed=77)
    kw = dict(kmin=1, kmax=8, w=30000, step=12000, mask_host=False, scaffolds_all=False, rip=True)
    ref = c_oracle.run(sc, threads=8, **kw)
    res = eng.run(eng.PackedGenome.from_scaffolds(sc), **kw)
    assert np.array_equal(res.tables, ref["tables"])
    assert np.array_equal(res.coords, ref["coords"])
    assert_rows_close(res.rows, ref["rows"], rtol_kld=1e-6, rtol_other=1e-15, what="long windows")
    assert max_rel_err(res.rows[:, 0], ref["rows"][:, 0]) < 1e-10


def test_low_complexity_windows(eng):
    """Homopolymer / dinucleotide-repeat windows: one bucket holds thousands of entries (the slow
    path of the bucketed kernel, maximal atomic contention in both kernels)."""
    from frisk_b200 import synth
    from oracle import c_oracle
    rng = np.random.Generator(np.random.PCG64(3))
    a = synth.iid_bases(rng, 40_000, 0.5)
    a[5_000:11_000] = ord("A")
    a[15_000:21_000] = np.tile(np.frombuffer(b"AT", dtype=np.uint8), 3000)
    a[25_000:30_000] = np.tile(np.frombuffer(b"ACGTTGCA", dtype=np.uint8), 625)
    a[31_000:31_040] = ord("N")
    sc = [("lowcomplex", a)]
    ref = c_oracle.run(sc, threads=4)
    res = eng.run(eng.PackedGenome.from_scaffolds(sc), dump=True)
    assert np.array_equal(res.tables, ref["tables"])
    print("low-complexity KLDs:", ref["rows"][:, 0], "abs diff:", np.abs(res.rows[:, 0] - ref["rows"][:, 0]))
    assert_rows_close(res.rows, ref["rows"], rtol_kld=1e-6, rtol_other=1e-15, what="low complexity")
    assert max_rel_err(res.rows[:, 0], ref["rows"][:, 0]) < 1e-10
    seq, off = c_oracle.concat(sc)
    for i in range(len(ref["rows"])):
        win = seq[int(ref["win_off"][i]):int(ref["win_off"][i]) + int(ref["win_len"][i])]
        _, _, wt, _ = c_oracle.window_tables(win, ref["tables"], ref["meta"], 1, 8)
        assert np.array_equal(res.win_tables[i].astype(np.uint64), wt), i


def test_bit_reproducible(eng):
    g = Golden("c1_small")
    a = _run_case(eng, g)
    b = _run_case(eng, g)
    assert np.array_equal(a.rows, b.rows, equal_nan=True)
    assert np.array_equal(a.tables, b.tables)


@pytest.mark.parametrize("config,scale,kw", [
    ("C1", 0.2, {}),
    ("C2", 0.05, dict(scaffolds_all=True)),
    ("C3", 0.01, dict(kmin=1, kmax=6)),
    ("C5", 0.0002, dict(scaffolds_all=True)),
    ("C4", 0.0005, dict(w=2000, step=500)),
    ("C2", 0.02, dict(w=1000, step=500, scaffolds_all=True)),          # short windows: the 2-round instantiation
    ("C1", 0.05, dict(w=2040, step=1020, kmin=3)),
    ("C1", 0.06, dict(w=8000, step=3000)),                              # the 8-round instantiation (5,115..8,186 bases)
    ("C2", 0.03, dict(w=8186, step=8186, scaffolds_all=True, kmax=5)),
    ("C2", 0.02, dict(kmax=6, w=60000, step=20000)),                    # small-K kernel, windows near its 65,535 limit
    ("C2", 0.02, dict(kmax=4, kmin=2)),
    ("C2", 0.02, dict(kmax=2)),
    ("C5", 0.002, dict(scaffolds_all=True)),                            # > 4,096 windows of mixed length: two length-class launches
])
def test_against_c_oracle_on_fresh_genomes(eng, config, scale, kw):
    from frisk_b200 import synth
    from oracle import c_oracle
    sc = synth.make(config, scale)
    full = dict(kmin=1, kmax=8, w=5000, step=2500, mask_host=False, scaffolds_all=False, rip=True)
    full.update(kw)
    ref = c_oracle.run(sc, threads=8, **full)
    res = eng.run(eng.PackedGenome.from_scaffolds(sc), **full)
    kmin, kmax = full["kmin"], full["kmax"]
    assert np.array_equal(res.tables, ref["tables"])
    assert list(res.meta) == [int(x) for x in ref["meta"]]
    assert res.names == ref["names"]
    assert np.array_equal(res.coords, ref["coords"])
    assert np.array_equal(res.status & 7, ref["status"] & 7)
    ok = ref["status"] == 0
    assert_rows_close(res.rows[ok], ref["rows"][ok], rtol_kld=1e-6, rtol_other=1e-15, what=config)
    assert max_rel_err(res.rows[ok, 0], ref["rows"][ok, 0]) < 1e-10


def test_full_size_c2_properties_and_sampled_parity(eng):
    """BASELINE config C2 at full size (40 Mbp, ~16 k windows): size-independent properties of the
    tables plus parity of the background and of 400 sampled windows against the C oracle."""
    from frisk_b200 import _lib, synth
    from oracle import c_oracle
    sc = synth.make("C2", 1.0)
    g = eng.PackedGenome.from_scaffolds(sc)
    res = eng.run(g)
    K = 8
    off = 0
    sums = []
    for k in range(1, K + 1):
        t = res.tables[off:off + 4 ** k].astype(np.int64)
        # strand symmetry: count(kmer) == count(revcomp(kmer))
        idx = np.arange(4 ** k)
        rc = np.zeros_like(idx)
        tmp = idx.copy()
        for _ in range(k):
            rc = (rc << 2) | ((tmp & 3) ^ 1)
            tmp >>= 2
        assert np.array_equal(t, t[rc]), "order %d not reverse-complement symmetric" % k
        sums.append(int(t.sum()))
        off += 4 ** k
    # both strands: total of order k = 2 * (#valid k-words) which is non-increasing in k, and
    # exMax closes the books for kmax: valid + excluded = all kmax-word start positions
    assert all(a >= b for a, b in zip(sums, sums[1:]))
    possible = int(np.maximum(g.scaf_len.astype(np.int64) - K + 1, 0).sum())
    assert sums[-1] // 2 + res.meta[1] == possible
    assert sums[0] // 2 == g.total_len - (g.nn_total - g.n_lower)
    # background vs C oracle (bit-exact), windows sampled
    seq, soff = c_oracle.concat(sc)
    tabs, meta = c_oracle.background(seq, soff, 1, 8, False, threads=8)
    assert np.array_equal(res.tables, tabs)
    assert list(res.meta) == [int(x) for x in meta]
    sidx, woff, wlen, st, sp = c_oracle.crawl(seq, soff)
    assert len(woff) == len(res.rows)
    assert np.array_equal(res.coords, np.stack([st, sp], 1))
    rng = np.random.Generator(np.random.PCG64(5))
    pick = np.sort(rng.choice(len(woff), 400, replace=False))
    rows, status = c_oracle.score(seq, woff[pick], wlen[pick], tabs, meta, threads=8)
    assert np.all(status == 0) and np.all(res.status[pick] == 0)
    assert_rows_close(res.rows[pick], rows, rtol_kld=1e-6, rtol_other=1e-15, what="C2 full")
    assert max_rel_err(res.rows[pick, 0], rows[:, 0]) < 1e-10
    assert np.all(np.isfinite(res.rows[:, 0])) and np.all(res.rows[:, 0] >= 0)


def test_hmm_anomaly_calls_identical_on_full_c1(eng):
    """north_star: "HMM anomaly calls on those scores must be identical".  Scores of the full-size C1
    config (5 Mbp, 2,000 rows) from the GPU vs the reference's own (golden c1_full, produced by executing
    the reference source): fit the 2-state HMM on each and compare the decoded state paths (tests/hmm_ref.py
    documents the hmmlearn stand-in)."""
    from tests import hmm_ref
    g = Golden("c1_full")
    res = _run_case(eng, g)
    err = _check_against_golden(res, g, "c1_full")
    assert err < 1e-10
    ref_kld, gpu_kld = g.vals[:, 0], res.rows[:, 0]
    m_ref, m_gpu = hmm_ref.fit(ref_kld), hmm_ref.fit(gpu_kld)
    path_ref, path_gpu = hmm_ref.predict(m_ref, ref_kld), hmm_ref.predict(m_gpu, gpu_kld)
    assert np.array_equal(path_ref, path_gpu)
    hi = int(np.argmax(m_ref[2]))
    called = int((path_ref == hi).sum())
    assert 0 < called < len(path_ref), "the planted islands are called, the background is not"
    # the reference's own model when the box has it (F:1539-1541: GaussianHMM(n_components=2, covariance_type="full");
    # the reference leaves random_state unset, so "identical" needs it fixed): same state path from both score vectors
    try:
        from hmmlearn import hmm
    except ImportError:
        return
    paths = []
    for kld in (ref_kld, gpu_kld):
        model = hmm.GaussianHMM(n_components=2, covariance_type="full", random_state=0)
        data = kld[~np.isnan(kld)][:, np.newaxis]
        model.fit(data)
        paths.append(model.predict(data))
    assert np.array_equal(paths[0], paths[1]), "hmmlearn state paths differ between reference and GPU scores"


def test_addressing_beyond_2_pow_32_bases(eng):
    """A small genome placed behind > 2^32 padding bases (C4/C5-scale offsets): every kernel must use
    64-bit base offsets.  Same tables and the same rows as the genome on its own."""
    import torch
    from frisk_b200 import _lib, synth
    sc = synth.make("C2", 0.01, seed=21) + synth.make("edge")
    g = eng.PackedGenome.from_scaffolds(sc)
    ref = eng.run(g, scaffolds_all=True)
    shift = (1 << 32) + (1 << 20)                       # bases, multiple of 128
    big_len = shift + g.padded_len
    dev = torch.device("cuda:0")
    codes = torch.zeros(big_len // 16, dtype=torch.int32, device=dev)
    inv = torch.full((big_len // 32,), -1, dtype=torch.int32, device=dev)      # all padding
    codes[shift // 16:] = torch.from_numpy(g.codes.view(np.int32)).to(dev)
    inv[shift // 32:] = torch.from_numpy(g.inv.view(np.int32)).to(dev)
    low = None
    if g.low is not None:
        low = torch.zeros(big_len // 32, dtype=torch.int32, device=dev)
        low[shift // 32:] = torch.from_numpy(g.low.view(np.int32)).to(dev)

    class Shifted:      # a DeviceGenome whose planes start 2^32 bases in
        pass
    dg = Shifted()
    dg.codes, dg.inv, dg.low, dg.device = codes, inv, low, dev
    dg.host = type("H", (), {"padded_len": big_len})()
    d_fwd = eng.background(dg, 8, False)
    d_tables, d_valid = eng.finalize(d_fwd, 8)
    d_ig = eng.genome_ivom(d_tables, 1, 8, g.genome_space)
    wins = g.windows(5000, 2500, True)
    moved = eng.WindowList(wins.off + np.uint64(shift), wins.length, wins.scaf, wins.start, wins.stop)
    d_rows, d_status, _ = eng.score(dg, moved, d_ig, 1, 8, True)
    torch.cuda.synchronize()
    out = eng.assemble(g, g, wins, d_tables.cpu().numpy().view(np.uint64), int(d_valid.item()), d_rows.cpu().numpy(),
                       d_status.cpu().numpy().view(np.uint32), 1, 8)
    assert np.array_equal(out.tables, ref.tables)
    assert out.meta == ref.meta
    assert np.array_equal(out.rows, ref.rows, equal_nan=True)
    assert np.array_equal(out.status, ref.status)


def test_background_16bit_counter_wraps_are_exact(eng):
    """The background kernel counts into 16-bit shared-memory counters, two per word.  Low-complexity
    sequence makes single bins exceed 65,535 inside one CTA's share: the low half wraps (carry into its
    neighbour), the high half wraps, and with the period-9 repeat both halves of one word (AAAAAAAA,
    AAAAAAAT) reach 0xFFFF together.  Tables must stay bit-exact."""
    from oracle import c_oracle
    a = np.full(20_000_000, ord("A"), np.uint8)
    p8 = np.tile(np.frombuffer(b"AAAAAAAT", np.uint8), 10_000_000)
    p9 = np.tile(np.frombuffer(b"AAAAAAAAT", np.uint8), 10_000_000)
    sc = [("polyA", a), ("period8", p8), ("period9", p9)]
    g = eng.PackedGenome.from_scaffolds(sc)
    dg = eng.DeviceGenome(g)
    d_tables, d_valid = eng.finalize(eng.background(dg, 8), 8)
    seq, off = c_oracle.concat(sc)
    tabs, meta = c_oracle.background(seq, off, 1, 8, False, threads=16)
    got = d_tables.cpu().numpy().view(np.uint64)
    assert np.array_equal(got, tabs)
    assert int(got[-4 ** 8]) > 2 * 65536 * 148           # AAAAAAAA: far beyond what 148 16-bit counters hold
    assert g.ex_max(8, int(d_valid.item())) == int(meta[1])


def test_kmax_sweep_shares_one_background_pass(eng):
    """BASELINE config C3 (k sweep 1..8, here at 1 % scale with its TE-like repeats): every k' of the sweep
    equals a separate reference-style run with --maxWordSize k' (C oracle)."""
    from frisk_b200 import synth
    from oracle import c_oracle
    sc = synth.make("C3", 0.01)
    sweep = eng.run_sweep(eng.PackedGenome.from_scaffolds(sc), kmaxes=range(1, 9))
    assert sorted(sweep) == list(range(1, 9))
    for k, res in sweep.items():
        ref = c_oracle.run(sc, threads=8, kmin=1, kmax=k)
        assert np.array_equal(res.tables, ref["tables"]), k
        assert list(res.meta) == [int(x) for x in ref["meta"]], k
        assert np.array_equal(res.coords, ref["coords"])
        ok = ref["status"] == 0
        assert np.array_equal(res.status & 7, ref["status"] & 7)
        assert_rows_close(res.rows[ok], ref["rows"][ok], rtol_kld=1e-6, rtol_other=1e-15, what="sweep k=%d" % k)
        assert max_rel_err(res.rows[ok, 0], ref["rows"][ok, 0]) < 1e-10


def test_concurrent_streams_and_threads_do_not_share_scratch(eng):
    """Two host threads, each with its own CUDA stream, run different genomes at the same time through
    the staged entry points, the one-call entry point and the device ingest; every result equals the
    serial one bit for bit (scratch is stream-ordered, the cached workspace is locked per device)."""
    import threading
    import torch
    from frisk_b200 import synth
    genomes = [synth.make("C2", 0.02, seed=s) + synth.make("edge") for s in (101, 202)]
    packed = [eng.PackedGenome.from_scaffolds(sc, pinned=True) for sc in genomes]
    texts = [np.frombuffer(synth.fasta_bytes(sc), dtype=np.uint8) for sc in genomes]
    serial = [eng.run(g, scaffolds_all=True) for g in packed]
    results = [[None] * 3 for _ in genomes]
    errors = []

    def work(i):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for _ in range(5):
                    results[i][0] = eng.run(packed[i], scaffolds_all=True)
                    results[i][1] = eng.run_host(packed[i], scaffolds_all=True)
                    results[i][2] = eng.run_fasta(texts[i], scaffolds_all=True)
            stream.synchronize()
        except Exception as e:            # surfaced below: a thread must not die silently
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(genomes))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i, ref in enumerate(serial):
        for res in results[i]:
            assert np.array_equal(res.tables, ref.tables)
            assert np.array_equal(res.rows, ref.rows, equal_nan=True) and np.array_equal(res.status, ref.status)


def test_fused_finalise_launch_is_stable_and_equals_the_staged_kernels(eng):
    """The cooperative finalise + genome-IVOM launch (grid-wide barriers between its phases) against the five
    separate kernels of the staged entry points, and against itself over 300 back-to-back launches."""
    import torch
    from frisk_b200 import synth
    g = eng.PackedGenome.from_scaffolds(synth.make("C2", 0.05, seed=3) + synth.make("edge"))
    pipe = eng.Pipeline(g, scaffolds_all=True)
    pipe.enqueue()
    torch.cuda.synchronize()
    t0, ig0, rows0 = pipe.d_tables.clone(), pipe.d_ig.clone(), pipe.d_rows.clone()
    d_tables, d_valid = eng.finalize(eng.background(pipe.dq, 8), 8)          # staged: forward_totals/low/symmetrise kernels
    d_ig = eng.genome_ivom(d_tables, 1, 8, g.genome_space)
    assert torch.equal(d_tables, t0) and int(d_valid.item()) == int(pipe.d_valid.item())
    assert torch.equal(d_ig.view(torch.int64), ig0.view(torch.int64))        # same bits, NaN entries included
    for _ in range(300):
        pipe.enqueue()
    torch.cuda.synchronize()
    assert torch.equal(pipe.d_tables, t0) and torch.equal(pipe.d_ig.view(torch.int64), ig0.view(torch.int64))
    assert torch.equal(pipe.d_rows.view(torch.int64), rows0.view(torch.int64))


def test_no_silent_fallback_symbols_loaded(eng):
    """The product library is the thing that ran: it is loaded in this process and reports a GPU."""
    from frisk_b200 import _lib
    assert _lib.device_count() >= 1
    with open("/proc/self/maps") as fh:
        assert "libfrisk_b200.so" in fh.read()


def test_fused_k_sweep_matches_the_per_kmax_launches(eng):
    """frisk_b200_score_sweep (kmax' = 1..8 from one pass over every window, BASELINE config C3) == eight separate
    launches, on the edge genome (lower case, IUPAC, N runs, low-complexity stretches that overflow the 4-bit counters
    and are handed over per kmax') and a C3 slice; and == the oracle for every kmax'."""
    from frisk_b200 import synth
    from oracle import c_oracle
    rng = np.random.Generator(np.random.PCG64(21))
    extra = synth.iid_bases(rng, 30_000, 0.4)
    extra[5_000:5_040] = ord("A")                                   # 33 x AAAAAAAA: nibble overflow -> hand-over
    extra[9_000:9_006] = ord("N")
    for p in range(15_000, 19_000, 37):                              # > 64 cut words in one window -> hand-over
        extra[p] = ord("N")
    sc = synth.make("edge") + synth.make("C3", 0.002) + [("extra", extra)]
    g = eng.PackedGenome.from_scaffolds(sc)
    for kw in (dict(), dict(w=2000, step=500, scaffolds_all=True), dict(w=8000, step=4000)):
        fused = eng.run_sweep(g, fused=True, **kw)
        plain = eng.run_sweep(g, fused=False, **kw)
        assert sorted(fused) == sorted(plain) == list(range(1, 9))
        for k in range(1, 9):
            a, b = fused[k], plain[k]
            assert np.array_equal(a.tables, b.tables) and a.meta == b.meta and a.names == b.names
            assert np.array_equal(a.coords, b.coords) and np.array_equal(a.status, b.status), (kw, k)
            ok = b.status == 0
            assert max_rel_err(a.rows[ok, 0], b.rows[ok, 0]) < 1e-12, (kw, k)
            assert np.array_equal(a.rows[:, 1:], b.rows[:, 1:], equal_nan=True), (kw, k)
            ref = c_oracle.run(sc, threads=4, kmin=1, kmax=k, w=kw.get("w", 5000), step=kw.get("step", 2500),
                               scaffolds_all=kw.get("scaffolds_all", False))
            assert np.array_equal(a.coords, ref["coords"]) and np.array_equal(a.status & 7, ref["status"] & 7)
            okr = ref["status"] == 0
            assert_rows_close(a.rows[okr], ref["rows"][okr], rtol_kld=1e-6, rtol_other=1e-15, what="sweep k=%d" % k)
            assert max_rel_err(a.rows[okr, 0], ref["rows"][okr, 0]) < 1e-10
