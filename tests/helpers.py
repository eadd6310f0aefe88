"""Shared helpers for the parity tests."""
from __future__ import annotations

import json
import os

import numpy as np

from frisk_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SMALL_CASES = ["edge_default", "edge_scaffoldsAll", "edge_maskHost", "edge_k2_5_w1000_i250",
               "edge_k1_3_w3000_i1000", "edge_k4_8", "edge_k1_1", "edge_k1_9", "edge_k9_10_w2500_i2500", "c1_small", "c2_small",
               "c2_small_query_vs_c1_host", "short_w1000_i800", "short_w1000_i800_all"]


class Golden:
    """One tests/golden/<case>.npz produced from the reference's own source (make_golden.py)."""

    def __init__(self, case: str):
        z = np.load(os.path.join(GOLDEN, case + ".npz"))
        self.meta = json.loads(str(z["meta"]))
        self.params = self.meta["params"]
        self.tables = z["genome_tables"]
        self.genome_meta = z["genome_meta"]
        self.names = [str(x) for x in z["row_names"]]
        self.coords = z["row_coords"]
        self.vals = z["row_vals"]                  # KLD, GC, PI, SI, CRI ; +inf = reference raised ZeroDivisionError
        self.win_pick = z["win_pick"]
        self.win_tables = z["win_tables"]
        self.win_meta = z["win_meta"]

    def scaffolds(self):
        sc = synth.make(*self.meta["genome"])
        assert synth.digest(sc) == self.meta["genome_digest"], "synthetic genome drifted from the golden run"
        return sc

    def host(self):
        if not self.meta["host"]:
            return None
        h = synth.make(*self.meta["host"])
        assert synth.digest(h) == self.meta["host_digest"]
        return h

    def kwargs(self):
        p = self.params
        return dict(kmin=p["kmin"], kmax=p["kmax"], w=p["w"], step=p["i"], mask_host=p["maskHost"],
                    scaffolds_all=p["scaffoldsAll"], rip=p["RIP"])


KLD_ATOL = 1e-13


def assert_rows_close(vals, ref, rtol_kld=1e-6, rtol_other=1e-12, what=""):
    """KLD within rtol_kld relative (north_star tolerance), GC/PI/SI/CRI within rtol_other,
    NaN positions identical.

    KLD also gets an absolute floor of 1e-13: a degenerate window (one or two distinct k-mers,
    e.g. a homopolymer run) has a true KLD of ~0 (1e-9 and below) which the reference itself only
    resolves to ~3e-16 absolute (it rounds w/G before the log), so no independent evaluation can
    match it to a relative 1e-6 there.  Real windows score 1e-3 .. 1 and are held to the relative
    tolerance."""
    vals = np.asarray(vals, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert vals.shape == ref.shape, (what, vals.shape, ref.shape)
    for col, name in enumerate(["KLD", "GC", "PI", "SI", "CRI"]):
        a, b = vals[:, col], ref[:, col]
        nan_a, nan_b = np.isnan(a), np.isnan(b)
        assert np.array_equal(nan_a, nan_b), "%s: NaN positions differ in %s" % (what, name)
        ok = ~nan_a & np.isfinite(b)
        tol = rtol_kld if col == 0 else rtol_other
        diff = np.abs(a[ok] - b[ok])
        if col == 0:
            diff = np.maximum(diff - KLD_ATOL, 0.0)
        err = diff / np.maximum(np.abs(b[ok]), 1e-300)
        assert err.size == 0 or err.max() <= tol, "%s: %s max rel err %.3e > %.1e" % (what, name, err.max(), tol)


def max_rel_err(a, b, atol=KLD_ATOL):
    a = np.asarray(a, float); b = np.asarray(b, float)
    ok = np.isfinite(a) & np.isfinite(b)
    if not ok.any():
        return 0.0
    diff = np.maximum(np.abs(a[ok] - b[ok]) - atol, 0.0)
    return float((diff / np.maximum(np.abs(b[ok]), 1e-300)).max())


def pca_golden():
    """tests/golden/pca_features.npz (make_pca_golden.py: the reference's own scrubMirrors / flattenKmerMap text):
    returns (regions [(name, uint8 array)], {(pcaMin, pcaMax): float64[n, F]}); NaN rows = ZeroDivisionError."""
    z = np.load(os.path.join(GOLDEN, "pca_features.npz"))
    meta = json.loads(str(z["meta"]))
    cache = {}
    regions = []
    for cfg, scale, s, a, l in meta["defs"]:
        if (cfg, scale) not in cache:
            cache[(cfg, scale)] = synth.make(cfg, scale)
            assert synth.digest(cache[(cfg, scale)]) == meta["digests"]["%s@%s" % (cfg, scale)], "synthetic genome drifted"
        regions.append(("%s_%d_%d_%d" % (cfg, s, a, l), cache[(cfg, scale)][s][1][a:a + l]))
    return regions, {(lo, hi): z["feat_%d_%d" % (lo, hi)] for lo, hi in meta["ranges"]}
