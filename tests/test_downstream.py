"""Host-side downstream glue: thresholds pinned to values produced by EXECUTING the
reference's own FDBins / otsu / setKLDThresh text (tests/golden/make_thresholds.py), interval building and
the HMM glue against hand-worked cases."""
import json
import os
import types

import numpy as np
import pandas as pd
import pytest

from frisk_b200 import downstream as ds
from tests import hmm_ref

HERE = os.path.dirname(os.path.abspath(__file__))


def _log_kld(case):
    kld = np.load(os.path.join(HERE, "golden", case + ".npz"))["row_vals"][:, 0]
    kld = kld[np.isfinite(kld) & (kld > 0)]
    return np.log10(kld[:, np.newaxis])          # (n, 1), as main() passes it (F:1530-1532)


@pytest.mark.parametrize("case", ["c1_full", "c1_small", "c2_small", "edge_default"])
def test_thresholds_match_the_reference_functions(case):
    gold = json.load(open(os.path.join(HERE, "golden", "thresholds.json")))["cases"][case]
    log = _log_kld(case)
    assert len(log) == gold["n"] and ds.FDBins(log) == gold["FDBins"]
    for key, kw in (("otsu", dict(threshTypeKLD="otsu")), ("percentile99.0", dict(threshTypeKLD="percentile", percentileKLD=99.0)),
                    ("percentile90.0", dict(threshTypeKLD="percentile", percentileKLD=90.0)),
                    ("force0.35", dict(threshTypeKLD=None, forceThresholdKLD="0.35"))):
        args = types.SimpleNamespace(threshTypeKLD=None, forceThresholdKLD=None, percentileKLD=99.0)
        args.__dict__.update(kw)
        thr, bins = ds.setKLDThresh(args, log)
        assert float(np.ravel(thr)[0]) == gold[key][0] and bins == gold[key][1], key       # same numpy calls: same bits
    assert float(np.ravel(ds.otsu(log, 40))[0]) == gold["otsu_40bins"]
    with pytest.raises(UnboundLocalError):                                               # the reference's behaviour
        ds.setKLDThresh(types.SimpleNamespace(threshTypeKLD=None, forceThresholdKLD=None, percentileKLD=99.0), log)


def test_find_base_ranges_and_merge():
    s = [0, 0, 1, 1, 1, 0, 1, 0, 0, 0]
    assert ds.findBaseRanges(s, 1) == [(2, 4), (6, 6)]
    assert ds.findBaseRanges(s, 0) == [(0, 1), (5, 5), (7, 9)]
    assert ds.findBaseRanges(s, 0, minlen=1) == [(0, 1), (7, 9)]
    assert ds.findBaseRanges("AANNNA", "N", name="chr1") == [("chr1", 2, 4)]
    recs = [("b", 10, 20, 1.0), ("a", 1, 5000, 0.5), ("a", 2501, 7500, 0.7), ("a", 7500, 9000, 0.1), ("a", 9006, 9100, 0.2)]
    assert ds.merge_intervals(recs, 0) == [("a", 1, 9000, 0.7, 0.1, (0.5 + 0.7 + 0.1) / 3), ("a", 9006, 9100, 0.2, 0.2, 0.2),
                                           ("b", 10, 20, 1.0, 1.0, 1.0)]
    assert ds.merge_intervals(recs, 6)[0][:3] == ("a", 1, 9100)


def _frame():
    rows = []
    for name, klds in (("s2", [0.01, 0.02, 0.5, 0.6, 0.02]), ("s1", [0.4, np.nan, 0.01, 0.01, 0.01, 0.7])):
        for i, k in enumerate(klds):
            rows.append((name, i * 2500 + 1, i * 2500 + 5000, k))
    return pd.DataFrame(rows, columns=["name", "start", "stop", "windowKLD"])


def test_threshold_kld_and_gff():
    df = _frame()
    args = types.SimpleNamespace(findSelf=False, mergeDist=0, dimReduce="features")
    anomalies, picked = ds.thresholdKLD(df, np.log10(0.3), args, merge=True)
    assert anomalies == [("s1", 1, 5000, 0.4, 0.4, 0.4), ("s1", 12501, 17500, 0.7, 0.7, 0.7), ("s2", 5001, 12500, 0.6, 0.5, 0.55)]
    assert list(picked["windowKLD"]) == [0.4, 0.7, 0.5, 0.6]
    lines = list(ds.anomaly2GFF(anomalies, args))
    assert lines[0] == "##gff-version 3\n"
    assert lines[3].split("\t")[:5] == ["s2", "frisk_" + ds.FRISK_VERSION, "Kmer-anomaly", "5001", "12500"]
    assert lines[3].rstrip().endswith("ID=Anomaly_3;maxKLD=0.6;minKLD=0.5;meanKLD=0.55")
    args.findSelf = True
    selfish, _ = ds.thresholdKLD(df, np.log10(0.015), args, merge=False)
    assert [r[:3] for r in selfish] == [("s1", 5001, 10000), ("s1", 7501, 12500), ("s1", 10001, 15000), ("s2", 1, 5000)]


def test_hmm_glue_on_planted_islands():
    """Full-size C1 scores (reference output, golden c1_full): fit, decode per scaffold, write GFF3; the built-in
    model is the tests' stand-in, and the planted islands come out as one of the two states."""
    g = np.load(os.path.join(HERE, "golden", "c1_full.npz"))
    kld = g["row_vals"][:, 0]
    df = pd.DataFrame({"name": [str(x) for x in g["row_names"]], "start": g["row_coords"][:, 0], "stop": g["row_coords"][:, 1],
                       "windowKLD": kld})
    model = ds.fit_hmm(kld)
    if isinstance(model, ds.GaussianHMM2):
        assert np.array_equal(model.predict(kld), hmm_ref.predict(hmm_ref.fit(kld), kld))
    bed, df2 = ds.hmm2BED(df, model)
    assert set(df2["hmmState"].dropna().unique()) == {0.0, 1.0}
    assert bed == sorted(bed, key=lambda r: (r[0], r[1], r[2])) and {r[3] for r in bed} == {"State1", "State2"}
    # intervals tile the scored windows: consecutive runs alternate states and cover first start .. last stop
    runs = sorted(bed, key=lambda r: int(r[1]))
    assert int(runs[0][1]) == int(df["start"].iloc[0]) and int(runs[-1][2]) == int(df["stop"].iloc[-1])
    assert all(a[3] != b[3] for a, b in zip(runs, runs[1:]))
    lines = list(ds.hmmBED2GFF(bed))
    assert lines[0] == "##gff-version 3\n" and len(lines) == len(bed) + 1
    f = lines[1].rstrip("\n").split("\t")
    assert f[1] == "frisk_" + ds.FRISK_VERSION and f[2] in ("State1", "State2") and f[8].startswith("ID=" + f[2] + "_")
    minority = min(("State1", "State2"), key=lambda s: sum(int(r[2]) - int(r[1]) for r in bed if r[3] == s))
    assert 10 <= sum(1 for r in bed if r[3] == minority) <= 60          # ~25 planted islands


def test_c_hmm_equals_the_numpy_formulation():
    """frisk_b200_hmm2_fit / _viterbi (C) against the numpy statement of the same algorithm: parameters to 1e-9,
    state paths identical, on the reference's own C1 scores and on a harder two-regime series."""
    import time
    rng = np.random.default_rng(4)
    series = [np.load(os.path.join(HERE, "golden", "c1_full.npz"))["row_vals"][:, 0],
              np.concatenate([rng.normal(0.2, 0.05, 700), rng.normal(0.5, 0.2, 90), rng.normal(0.2, 0.05, 1300),
                              rng.normal(0.6, 0.1, 40), rng.normal(0.25, 0.05, 500)])]
    for x in series:
        a, b = ds.GaussianHMM2().fit(x), ds.GaussianHMM2().fit_numpy(x)
        for name in ("startprob_", "transmat_", "means_", "vars_"):
            assert np.allclose(getattr(a, name), getattr(b, name), rtol=1e-9, atol=1e-12), name
        assert np.array_equal(a.predict(x), b.predict_numpy(x))
        assert np.array_equal(a.predict(x), hmm_ref.predict(hmm_ref.fit(x), x))
    big = np.abs(rng.normal(0.3, 0.1, 1_200_000)) + 1e-3
    t = time.time()
    m = ds.GaussianHMM2().fit(big)
    path = m.predict(big)
    assert time.time() - t < 20 and len(path) == len(big)           # a C4-sized table of window scores
