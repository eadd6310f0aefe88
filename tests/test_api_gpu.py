"""GPU tests of the reference-facing function surface (computeKmers / IvomBuild / KLD / main)."""
import pickle
import types

import os

import numpy as np
import pytest

from tests.helpers import Golden, assert_rows_close, max_rel_err

pytestmark = pytest.mark.gpu


def _args(path, g, tmp, **extra):
    p = g.params
    d = dict(hostSeq=path, querySeq=None, minWordSize=p["kmin"], maxWordSize=p["kmax"], windowlen=p["w"], increment=p["i"],
             maskHost=p["maskHost"], scaffoldsAll=p["scaffoldsAll"], RIP=p["RIP"], pcaMin=1, pcaMax=6, tempDir=str(tmp))
    d.update(extra)
    return types.SimpleNamespace(**d)


def _arr(maps, kmin, kmax):
    return np.concatenate([np.fromiter(maps[k - kmin].values(), dtype=np.uint64) for k in range(kmin, kmax + 1)])


@pytest.mark.parametrize("case", ["edge_default", "edge_maskHost", "edge_k2_5_w1000_i250"])
def test_compute_kmers_and_scoring_functions(tmp_path, case):
    import frisk
    from frisk_b200 import synth
    from oracle import frisk_oracle
    g = Golden(case)
    path = str(tmp_path / "g.fa")
    synth.write_fasta(g.scaffolds(), path)
    args = _args(path, g, tmp_path)
    kmin, kmax = args.minWordSize, args.maxWordSize
    blank = frisk.rangeMaps(kmin, kmax)
    pk = str(tmp_path / "genome.p")
    genome = frisk.computeKmers(args, genomepickle=pk, window=None, genomeMode=True, kmerMap=blank, getMeta=True)
    assert np.array_equal(_arr(genome, kmin, kmax), g.tables)
    kr = kmax - kmin
    assert [genome[kr + 1]["totalLen"], genome[kr + 2]["exMax"], genome[kr + 3]["nnTotal"]] == [int(x) for x in g.genome_meta]
    assert list(genome[0].keys()) == list(blank[0].keys())                      # reference dict order
    assert pickle.load(open(pk, "rb")) == genome                               # F:363
    # the per-window loop of the reference's main(), function by function (F:1478-1488), first 3 windows
    rows = []
    for n, (seq, name, start, stop) in enumerate(frisk.crawlGenome(args, path)):
        if n == 3:
            break
        win = frisk.computeKmers(args, window=[(name, seq)], genomeMode=False, kmerMap=blank, getMeta=True)
        ref_win = frisk_oracle.compute_kmers([(name, seq)], kmin, kmax, both_strands=False)
        assert win == ref_win
        gi = frisk.IvomBuild(win, args, genome, True)
        wi = frisk.IvomBuild(win, args, genome, False)
        ref_gi = frisk_oracle.ivom_build(ref_win, genome, kmin, kmax, True)
        assert list(gi.keys()) == list(ref_gi.keys())
        assert max_rel_err(list(gi.values()), list(ref_gi.values()), atol=0) < 1e-12
        kld = frisk.KLD(gi, wi, args)
        rows.append((name, start, stop, kld, frisk.calcGC(seq)) + tuple(frisk.calcRIP(win, args)))
    vals = np.array([r[3:] for r in rows], float)
    assert [r[0] for r in rows] == g.names[:3]
    assert_rows_close(vals, g.vals[:3], rtol_kld=1e-6, rtol_other=0.0, what=case + "[function API]")
    # symmetric (PCA-feature) counting: second caller of the counting kernel (F:1571-1578)
    seq, name = rows and frisk_oracle.iter_fasta(path).__next__()[1][:4000], "x"
    sym = frisk.computeKmers(types.SimpleNamespace(pcaMin=1, pcaMax=6, maskHost=False), window=[(name, seq)], pcaMode=True,
                             kmerMap=frisk.rangeMaps(1, 6), getMeta=False, sym=True)
    assert sym == frisk_oracle.compute_kmers([(name, seq)], 1, 6, both_strands=True)[:6]


@pytest.mark.parametrize("case,flags", [("edge_default", ["--RIP"]), ("edge_scaffoldsAll", ["--RIP", "--scaffoldsAll"]),
                                        ("edge_k2_5_w1000_i250", ["--RIP", "-m", "2", "-k", "5", "-w", "1000", "-i", "250"])])
def test_cli_main_writes_reference_products(tmp_path, case, flags, capsys):
    import frisk
    import pandas as pd
    from frisk_b200 import synth
    g = Golden(case)
    path = str(tmp_path / "genome.fa")
    synth.write_fasta(g.scaffolds(), path)
    tmp = tmp_path / "temp"
    argv = ["-H", path, "-t", str(tmp), "--exitAfter", "WindowKLD"] + flags
    with pytest.raises(SystemExit) as ei:
        frisk.main(argv)
    assert ei.value.code == 0
    out = capsys.readouterr().out
    # raw TSV (F:1475, F:1493)
    tsv = pd.read_csv(tmp / "raw_window_scores.bed", sep="\t", float_precision="round_trip")
    assert list(tsv.columns) == ["name", "start", "stop", "windowKLD", "GC", "PI", "SI", "CRI"]
    assert list(tsv["name"]) == g.names
    assert np.array_equal(tsv[["start", "stop"]].to_numpy(), g.coords)
    assert_rows_close(tsv[["windowKLD", "GC", "PI", "SI", "CRI"]].to_numpy(float), g.vals, rtol_kld=1e-6, rtol_other=1e-15,
                      what=case + "[TSV]")
    assert out.count("\n") >= len(g.names)                                     # rows echoed to stdout (F:1494)
    # caches (F:501, F:505)
    a = frisk.mainArgs(argv)
    genome = pickle.load(open(frisk.makePicklePath(a, space="genome"), "rb"))
    kmin, kmax = a.minWordSize, a.maxWordSize
    assert np.array_equal(_arr(genome, kmin, kmax), g.tables)
    kr = kmax - kmin
    assert [genome[kr + 1]["totalLen"], genome[kr + 2]["exMax"], genome[kr + 3]["nnTotal"]] == [int(x) for x in g.genome_meta]
    frame = pd.read_pickle(frisk.makePicklePath(a, space="window"))
    assert list(frame["name"]) == g.names and np.array_equal(frame["windowKLD"].to_numpy(), tsv["windowKLD"].to_numpy())
    # second run re-uses both caches (default: --recalc / --recalcWin not given) and returns the frame
    frame2 = frisk.main(["-H", path, "-t", str(tmp), "--quiet"] + flags)
    assert frame2.equals(frame)


def test_cli_hmm_stage_writes_state_intervals(tmp_path):
    """--hmmKLD (F:1536-1548): scores from the GPU -> 2-state HMM -> 2StateHmm.gff3; the state paths equal those
    obtained from the reference's own scores (golden c1_small)."""
    import frisk
    from frisk_b200 import downstream, synth
    g = Golden("c1_small")
    path = str(tmp_path / "genome.fa")
    synth.write_fasta(g.scaffolds(), path)
    tmp = tmp_path / "temp"
    frame = frisk.main(["-H", path, "-t", str(tmp), "--quiet", "--RIP", "--hmmKLD"])
    lines = open(tmp / "2StateHmm.gff3").read().splitlines()
    assert lines[0] == "##gff-version 3" and len(lines) >= 3
    assert "hmmState" in frame.columns
    ref_model = downstream.fit_hmm(g.vals[:, 0])
    if isinstance(ref_model, downstream.GaussianHMM2):
        assert np.array_equal(frame["hmmState"].to_numpy().astype(int), ref_model.predict(g.vals[:, 0]))


def test_batch_pca_features_match_the_reference_formulation():
    """frisk_b200_region_features (one CTA per region) vs the oracle's restatement of
    computeKmers(pcaMode, sym) + scrubMirrors + flattenKmerMap(prop=True) (F:1571-1591), and vs this
    package's own per-region dict path."""
    import argparse
    from frisk_b200 import api, engine, synth
    from oracle import frisk_oracle as fo
    rng = np.random.default_rng(8)
    edge = synth.make("edge")
    big = synth.make("C1", 0.02)[0][1]
    regions = [("r%d" % i, big[a:a + int(l)]) for i, (a, l) in enumerate(zip(rng.integers(0, 90_000, 12), rng.integers(300, 9000, 12)))]
    regions += [(n, s[:4000]) for n, s in edge[:3]]                       # lower case, N runs, IUPAC inside
    regions.append(("long", big[:100_000]))
    args = argparse.Namespace(pcaMin=1, pcaMax=6)
    labels, feats = api.pcaFeatures(args, regions)
    assert list(labels) == [n for n, _ in regions] and feats.shape == (len(regions), 2772)
    for (name, seq), row in zip(regions, feats):
        want = np.array(fo.region_features(seq.tobytes().decode(), 1, 6))
        assert np.array_equal(row, want), name                           # one exact division per entry
    # the dict-level API path (computeKmers on the GPU + host scrub/flatten) agrees
    name, seq = regions[0]
    args2 = argparse.Namespace(pcaMin=2, pcaMax=4, minWordSize=1, maxWordSize=8, maskHost=False, hostSeq="")
    maps = api.computeKmers(args2, window=[(name, seq.tobytes().decode())], pcaMode=True, kmerMap=api.rangeMaps(2, 4),
                            getMeta=False, sym=True)
    vec = api.flattenKmerMap(api.scrubMirrors(maps), window=5000, seqLen=len(seq), kmin=2, kmax=4, prop=True)
    _, f24 = api.pcaFeatures(argparse.Namespace(pcaMin=2, pcaMax=4), [regions[0]])
    assert np.array_equal(vec, f24[0])
    with pytest.raises(ZeroDivisionError):
        api.pcaFeatures(args, [("tiny", np.frombuffer(b"ACG", np.uint8))])   # no 4-mer at all: F:824 divides by zero


def test_batch_pca_features_match_reference_text_golden():
    """Row f3 pin on the device: frisk_b200_region_features == tests/golden/pca_features.npz, the vectors the
    reference's OWN text (computeKmers(pcaMode, sym) F:280-367, scrubMirrors F:797-811, flattenKmerMap F:813-831)
    produced for these regions, bit for bit; regions on which the reference raises ZeroDivisionError raise here."""
    import argparse
    from frisk_b200 import api
    from tests.helpers import pca_golden
    regions, gold = pca_golden()
    for (lo, hi), want in gold.items():
        ok = ~np.isnan(want).all(axis=1)
        args = argparse.Namespace(pcaMin=lo, pcaMax=hi)
        labels, feats = api.pcaFeatures(args, [r for r, k in zip(regions, ok) if k])
        assert feats.shape == want[ok].shape and np.array_equal(feats, want[ok]), (lo, hi)
        for r, k in zip(regions, ok):
            if not k:
                with pytest.raises(ZeroDivisionError):
                    api.pcaFeatures(args, [r])


def test_pca_features_straight_from_the_planes(tmp_path):
    """thresholdKLD intervals -> composition vectors read from the resident planes (no getFasta / getBEDSeq
    strings) == the vectors of the extracted sequences (reference slicing rule seq[start-1:stop])."""
    import argparse
    from frisk_b200 import api, downstream, engine, synth
    sc = synth.make("C1", 0.04) + synth.make("edge")[:2]
    text = synth.fasta_bytes(sc)
    dg = engine.DeviceGenome.from_fasta_bytes(text)
    res = engine.run(dg, scaffolds_all=True)
    frame = api.windows_frame(res, with_rip=True)
    targs = argparse.Namespace(findSelf=False, mergeDist=0, dimReduce="features")
    thr = np.percentile(np.log10(frame["windowKLD"][frame["windowKLD"] > 0]), 80.0)
    merged, _ = downstream.thresholdKLD(frame, thr, targs, merge=True)
    assert len(merged) >= 5
    merged = merged + [("no_such_scaffold", 1, 100, 0.1, 0.1, 0.1)]
    args = argparse.Namespace(pcaMin=1, pcaMax=6)
    labels, feats = api.pcaFeaturesFromIntervals(args, dg, merged)
    assert len(labels) == len(merged) - 1 and labels[0] == "%s:%d:%d" % merged[0][:3]
    seqs = dict(sc)
    regions = [(lab, seqs[rec[0]][int(rec[1]) - 1:int(rec[2])]) for lab, rec in zip(labels, merged)]
    labels2, feats2 = api.pcaFeatures(args, regions)
    assert list(labels2) == list(labels) and np.array_equal(feats, feats2)


def test_cli_runs_without_importing_torch(tmp_path):
    """The CLI path goes through the C ABI alone (native device buffers, one frisk_b200_run_resident call):
    PyTorch -- whose import alone takes seconds -- must not be loaded, and the products equal the golden ones."""
    import subprocess
    import sys
    import pandas as pd
    from frisk_b200 import synth
    g = Golden("c2_small")
    path = str(tmp_path / "genome.fa")
    synth.write_fasta(g.scaffolds(), path)
    tmp = tmp_path / "temp"
    code = ("import sys, frisk; frisk.main(['-H', %r, '-t', %r, '--quiet', '--RIP']); "
            "assert 'torch' not in sys.modules, 'torch was imported'; print('no torch')" % (path, str(tmp)))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, timeout=600)
    assert out.returncode == 0 and "no torch" in out.stdout, out.stdout[-1000:] + out.stderr[-3000:]
    tsv = pd.read_csv(tmp / "raw_window_scores.bed", sep="\t", float_precision="round_trip")
    assert list(tsv["name"]) == g.names and np.array_equal(tsv[["start", "stop"]].to_numpy(), g.coords)
    assert_rows_close(tsv[["windowKLD", "GC", "PI", "SI", "CRI"]].to_numpy(float), g.vals, rtol_kld=1e-6, rtol_other=1e-15,
                      what="c2_small[CLI, native buffers]")
