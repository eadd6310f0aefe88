"""CPU: pin both oracles (Python restatement, C restatement) to the golden fixtures that
were produced by executing the reference's own source text (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import c_oracle, frisk_oracle, ref_exec
from tests.helpers import Golden, SMALL_CASES, assert_rows_close, max_rel_err


def _tables_array(maps, kmin, kmax):
    return np.concatenate([np.fromiter(maps[k - kmin].values(), dtype=np.uint64) for k in range(kmin, kmax + 1)])


@pytest.mark.parametrize("case", SMALL_CASES)
def test_c_oracle_matches_reference(case):
    g = Golden(case)
    out = c_oracle.run(g.scaffolds(), host=g.host(), threads=4, **g.kwargs())
    assert np.array_equal(out["tables"], g.tables), "genome tables must be bit-exact"
    assert np.array_equal(out["meta"], g.genome_meta), "totalLen / exMax / nnTotal"
    assert out["names"] == g.names
    assert np.array_equal(out["coords"], g.coords)
    ref = g.vals.copy()
    zd = np.isinf(ref[:, 0])
    assert np.array_equal(zd, (out["status"] & c_oracle.ERR_KLD_ZERODIV) != 0), "ZeroDivisionError rows"
    ref[zd, 0] = 0.0
    out["rows"][zd, 0] = 0.0
    # same operation order as the reference -> agreement far inside the 1e-6 contract
    assert_rows_close(out["rows"], ref, rtol_kld=1e-12, rtol_other=0.0, what=case)


@pytest.mark.parametrize("case", ["edge_default", "edge_k2_5_w1000_i250", "edge_scaffoldsAll"])
def test_c_oracle_window_tables(case):
    g = Golden(case)
    sc = g.scaffolds()
    kw = g.kwargs()
    seq, off = c_oracle.concat(sc)
    _, woff, wlen, _, _ = c_oracle.crawl(seq, off, kw["w"], kw["step"], kw["scaffolds_all"])
    for slot, idx in enumerate(g.win_pick):
        win = seq[int(woff[idx]):int(woff[idx]) + int(wlen[idx])]
        _, _, wt, wm = c_oracle.window_tables(win, g.tables, g.genome_meta, kw["kmin"], kw["kmax"])
        assert np.array_equal(wt, g.win_tables[slot].astype(np.uint64))
        assert np.array_equal(wm, g.win_meta[slot])


@pytest.mark.parametrize("case", ["edge_default", "edge_scaffoldsAll", "edge_maskHost", "edge_k1_3_w3000_i1000",
                                  "edge_k4_8"])
def test_python_oracle_matches_reference(case):
    g = Golden(case)
    sc = [(n, s.tobytes().decode()) for n, s in g.scaffolds()]
    kw = g.kwargs()
    genome, rows = frisk_oracle.score_windows(sc, **kw)
    kmin, kmax = kw["kmin"], kw["kmax"]
    assert np.array_equal(_tables_array(genome, kmin, kmax), g.tables)
    kr = kmax - kmin
    assert [genome[kr + 1]["totalLen"], genome[kr + 2]["exMax"], genome[kr + 3]["nnTotal"]] == list(g.genome_meta)
    assert [r[0] for r in rows] == g.names
    assert np.array_equal(np.array([[r[1], r[2]] for r in rows]).reshape(-1, 2), g.coords)
    vals = np.array([[np.inf if isinstance(v, str) else (np.nan if v is None else v) for v in r[3:]] for r in rows],
                    dtype=float).reshape(-1, 5)
    assert np.array_equal(np.isinf(vals[:, 0]), np.isinf(g.vals[:, 0]))
    # identical operation order in the same language: bit-identical expected
    fin = np.isfinite(g.vals)
    assert np.array_equal(vals[fin], g.vals[fin])


def test_python_and_c_oracle_agree_on_fresh_input():
    from frisk_b200 import synth
    sc = synth.make("C2", 0.002, seed=99)
    out = c_oracle.run(sc, threads=2)
    genome, rows = frisk_oracle.score_windows([(n, s.tobytes().decode()) for n, s in sc])
    assert np.array_equal(_tables_array(genome, 1, 8), out["tables"])
    assert len(rows) == len(out["rows"])
    assert max_rel_err([r[3] for r in rows], out["rows"][:, 0]) < 1e-12


@pytest.mark.skipif(not ref_exec.available(), reason="/root/reference only exists in the build container")
def test_live_reference_text_matches_golden_and_oracle():
    """Container-only: re-execute the reference's source text and compare (guards the fixtures)."""
    import os, tempfile
    from frisk_b200 import synth
    g = Golden("edge_k1_3_w3000_i1000")
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "q.fa")
    synth.write_fasta(g.scaffolds(), path)
    p = g.params
    args = ref_exec.make_args(path, querySeq=path, **p)
    genome, rows = ref_exec.run_hot_path(args, genomepickle=os.path.join(tmp, "g.p"))
    assert np.array_equal(_tables_array(genome, p["kmin"], p["kmax"]), g.tables)
    assert [r[0] for r in rows] == g.names
    assert np.array_equal(np.array([r[3] for r in rows], float), g.vals[:, 0])


def test_pca_feature_restatement_known_answers():
    """scrubMirrors / flattenKmerMap(prop=True) (F:797-831) on a sequence small enough to do by hand."""
    from oracle import frisk_oracle as fo
    # AACG: 1-mers fwd A2 C1 G1, + revcomp (CGTT) C1 G1 T2 -> A2 T2 G2 C2; kept A (T is its mirror), G (C its mirror)
    v = fo.region_features("AACG", 1, 1)
    assert v == [0.5, 0.5]
    # 2-mers of AACG: AA AC CG; revcomps TT GT CG -> AA1 AC1 CG2 TT1 GT1; kept in table order (A,T,G,C digits):
    # AA AT AG AC TA TG TC GG GC CG (10 of 16: GA is the mirror of TC, which comes first); counts AA1 AC1 CG2, others 0 -> total 4
    maps = fo.scrub_mirrors(fo.compute_kmers([("r", "AACG")], 2, 2, both_strands=True)[:1])
    assert list(maps[0]) == ["AA", "AT", "AG", "AC", "TA", "TG", "TC", "GG", "GC", "CG"]
    assert fo.region_features("AACG", 2, 2) == [0.25, 0, 0, 0.25, 0, 0, 0, 0, 0, 0.5]
    assert len(fo.region_features("ACGTTGCAACGT" * 5, 1, 6)) == 2 + 10 + 32 + 136 + 512 + 2080


def test_pca_feature_restatement_matches_reference_text_golden():
    """Row f3 pin: oracle.region_features == the vectors the reference's OWN computeKmers(pcaMode, sym) +
    scrubMirrors + flattenKmerMap(prop=True) text produced (tests/golden/pca_features.npz), bit for bit."""
    from oracle import frisk_oracle as fo
    from tests.helpers import pca_golden
    regions, gold = pca_golden()
    for (lo, hi), want in gold.items():
        for (name, seq), row in zip(regions, want):
            if np.isnan(row).all():
                with pytest.raises(ZeroDivisionError):
                    fo.region_features(seq.tobytes().decode(), lo, hi)
                continue
            if len(seq) > 20_000 and (lo, hi) != (1, 6):
                continue                                           # the long region once is enough for the Python loops
            assert np.array_equal(np.array(fo.region_features(seq.tobytes().decode(), lo, hi)), row), (name, lo, hi)


@pytest.mark.skipif(not ref_exec.available(), reason="reference tree not present (GPU box)")
def test_pca_golden_is_what_the_reference_text_produces():
    """Container-only: re-execute the reference's text on two regions and compare with the committed golden."""
    from tests.helpers import pca_golden
    regions, gold = pca_golden()
    pick = [0, 9, 12]
    vecs = ref_exec.run_pca_features([(regions[i][0], regions[i][1].tobytes().decode()) for i in pick], 2, 4)
    for i, v in zip(pick, vecs):
        assert np.array_equal(v, gold[(2, 4)][i])
