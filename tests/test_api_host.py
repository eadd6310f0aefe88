"""CPU tests of the reference-facing function surface (the string/dict helpers and the CLI parser)."""
import numpy as np
import pytest

import frisk
import frisk_b200
from frisk_b200 import synth
from oracle import frisk_oracle


def test_module_surface_matches_reference_names():
    # the names the reference's main() uses
    for name in ["computeKmers", "crawlGenome", "IvomBuild", "KLD", "calcGC", "calcRIP", "rangeMaps", "main", "mainArgs",
                 "iterFasta", "countN", "revComplement", "prepareMaps", "makePicklePath", "tempPathCheck", "LETTERS"]:
        assert hasattr(frisk, name) and hasattr(frisk_b200, name), name
    assert frisk.LETTERS == ("A", "T", "G", "C")


def test_string_helpers_match_oracle():
    rng = np.random.default_rng(3)
    seq = "".join(rng.choice(list("ACGTacgtNnRY-"), 5000))
    assert frisk.countN(seq) == frisk_oracle.count_n(seq)
    assert frisk.calcGC(seq) == frisk_oracle.calc_gc(seq)
    with pytest.raises(ZeroDivisionError):
        frisk.calcGC("nnnnacgt")
    assert frisk.revComplement("AACGT") == frisk_oracle.rev_complement("AACGT") == "ACGTT"
    maps = frisk.rangeMaps(2, 4)
    ref = frisk_oracle.range_maps(2, 4)
    assert [list(m.items()) for m in maps] == [list(m.items()) for m in ref]      # same keys, same order, zeros
    assert list(frisk.prepareMaps(0, 2, [""]).keys()) == list(ref[0].keys())


def test_calc_rip_rules():
    import types
    args = types.SimpleNamespace(minWordSize=1, maxWordSize=3)
    di = dict.fromkeys(frisk.rangeMaps(2, 2)[0], 0)
    di.update(AT=4, TA=2, AC=1, GT=1, CA=3, TG=1)
    win = [{}, di, {}]
    assert frisk.calcRIP(win, args) == frisk_oracle.calc_rip(win, 1, 3) == (0.5, 2.0, -1.5)
    di.update(TA=0)                                  # PI == 0.0 is falsy -> CRI NaN (F:491)
    pi, si, cri = frisk.calcRIP(win, args)
    assert pi == 0.0 and si == 2.0 and np.isnan(cri)
    di.update(AT=0)
    pi, si, cri = frisk.calcRIP(win, args)
    assert np.isnan(pi) and np.isnan(cri)
    with pytest.raises(ValueError):
        frisk.calcRIP(win, types.SimpleNamespace(minWordSize=3, maxWordSize=5))


@pytest.mark.parametrize("w,i,sa", [(5000, 2500, False), (3000, 1000, True), (1000, 250, False)])
def test_iter_fasta_and_crawl_genome_match_oracle(tmp_path, w, i, sa):
    import types
    sc = synth.make("edge")
    path = str(tmp_path / "edge.fa")
    synth.write_fasta(sc, path)
    recs = list(frisk.iterFasta(path))
    assert recs == list(frisk_oracle.iter_fasta(path))
    args = types.SimpleNamespace(windowlen=w, increment=i, scaffoldsAll=sa)
    got = list(frisk.crawlGenome(args, path))
    ref = list(frisk_oracle.crawl_genome(recs, w, i, sa))
    assert got == ref


def test_cli_parser_defaults_and_paths():
    args = frisk.mainArgs(["-H", "genomes/host.fa"])
    assert (args.minWordSize, args.maxWordSize, args.windowlen, args.increment) == (1, 8, 5000, 2500)
    assert args.recalc is True and args.recalcWin is True and args.exitAfter is None and args.tempDir == "temp"
    assert frisk.makePicklePath(args, space="genome") == "temp/host.fa_kmers_1_8_genome.p"
    assert frisk.makePicklePath(args, space="window") == "temp/host.fa_kmers_1_8_KLD_window_5000_increment_2500.p"
    args = frisk.mainArgs(["-H", "h.fa", "-Q", "q.fa", "--recalc", "-m", "2", "-k", "5", "-w", "1000", "-i", "250"])
    assert args.recalc is False
    assert frisk.makePicklePath(args, space="window") == "temp/q.fa_kmers_2_5_KLD_window_1000_increment_250.p"
    with pytest.raises(SystemExit):
        frisk.mainArgs(["-H", "h.fa", "-m", "5", "-k", "3"])


def test_scrub_mirrors_and_flatten_match_oracle():
    from frisk_b200 import _lib, api, engine
    from oracle import frisk_oracle as fo
    seq = synth.make("edge")[0][1][:3000].tobytes().decode().upper().replace("N", "A")
    maps = fo.compute_kmers([("r", seq)], 1, 4, both_strands=True)[:4]
    ours = api.flattenKmerMap(api.scrubMirrors(maps), kmin=1, kmax=4, prop=True)
    assert np.array_equal(ours, np.array(fo.flatten_props(fo.scrub_mirrors(maps))))
    counts = api.flattenKmerMap(maps, window=5000, seqLen=len(seq), kmin=2, kmax=3, prop=False)
    assert len(counts) == 16 + 64 and np.isclose(counts.sum(), (5000.0 / len(seq)) * (2 * (len(seq) - 1) + 2 * (len(seq) - 2)))
    slot, nf = engine.feature_slots(1, 6)
    assert nf == 2 + 10 + 32 + 136 + 512 + 2080 and slot.shape[0] == _lib.table_size(1, 6)
    # kept = first of each pair in table order
    keys = api._kmer_keys(3)
    kept = [k for k, sl in zip(keys, slot[20:84]) if sl >= 0]
    assert kept == list(fo.scrub_mirrors([dict.fromkeys(keys, 0)])[0])
