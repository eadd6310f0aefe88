"""Golden vectors for the PCA-feature path, produced by EXECUTING THE REFERENCE'S OWN TEXT:
computeKmers(pcaMode=True, sym=True) F:280-367 -> scrubMirrors F:797-811 -> flattenKmerMap(prop=True) F:813-831,
the loop at F:1571-1591.  Container-only (needs /root/reference; oracle/ref_exec.py).  Run from the repo root:

    python tests/golden/make_pca_golden.py

Writes tests/golden/pca_features.npz: region definitions (synth config, scale, scaffold, start, length) and one
feature matrix per (pcaMin, pcaMax) range; a row of NaN marks a region on which the reference raises
ZeroDivisionError (F:824).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from frisk_b200 import synth  # noqa: E402
from oracle import ref_exec  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
RANGES = [(1, 6), (2, 4), (1, 1), (3, 3)]


def region_defs():
    """(config, scale, scaffold index, start, length)"""
    defs = []
    edge = synth.make("edge")
    for s, (name, seq) in enumerate(edge):                       # lower case, IUPAC, '-', N runs, tiny scaffolds
        defs.append(("edge", 1.0, s, 0, min(len(seq), 4000)))
    defs.append(("edge", 1.0, 0, 2900, 700))                     # spans the soft-masked block
    defs.append(("edge", 1.0, 1, 4900, 1800))                    # mostly N
    rng = np.random.default_rng(8)
    n1 = len(synth.make("C1", 0.02)[0][1])
    for a, l in zip(rng.integers(0, n1 - 9000, 8), rng.integers(300, 9000, 8)):
        defs.append(("C1", 0.02, 0, int(a), int(l)))
    defs.append(("C1", 0.02, 0, 0, 60_000))                      # a long region (counts beyond 16 bits at order 1)
    return defs


def regions_from(defs):
    cache = {}
    out = []
    for cfg, scale, s, a, l in defs:
        sc = cache.setdefault((cfg, scale), synth.make(cfg, scale))
        out.append(("%s_%d_%d_%d" % (cfg, s, a, l), sc[s][1][a:a + l]))
    return out


def main():
    defs = region_defs()
    regions = regions_from(defs)
    strs = [(n, s.tobytes().decode()) for n, s in regions]
    arrays = {}
    for lo, hi in RANGES:
        vecs = ref_exec.run_pca_features(strs, lo, hi)
        width = max(len(v) for v in vecs if not isinstance(v, str))
        m = np.full((len(vecs), width), np.nan)
        for i, v in enumerate(vecs):
            if not isinstance(v, str):
                m[i] = v
        arrays["feat_%d_%d" % (lo, hi)] = m
        print("pca %d..%d: %d regions x %d features, %d ZeroDivisionError" %
              (lo, hi, len(vecs), width, sum(isinstance(v, str) for v in vecs)))
    meta = dict(defs=defs, ranges=RANGES, ref_sha256=ref_exec.REF_SHA256,
                digests={"%s@%s" % (c, s): synth.digest(synth.make(c, s)) for c, s in {(d[0], d[1]) for d in defs}})
    np.savez_compressed(os.path.join(HERE, "pca_features.npz"), meta=np.array(json.dumps(meta)), **arrays)


if __name__ == "__main__":
    main()
