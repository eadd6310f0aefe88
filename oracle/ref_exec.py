"""Container-only harness: run the reference's OWN hot-path source text under py3.

TEST INFRASTRUCTURE ONLY.  This module reads /root/reference (which exists only in
the build container, never on the GPU box) and is used exclusively by
``tests/golden/make_golden.py`` to generate the committed golden fixtures and by the
(auto-skipped when the reference is absent) cross-checks in ``tests/test_oracle.py``.
Nothing in the product package, bench.py or the ``-m gpu`` tests imports it.

No reference source is copied into this repository: the text is sliced by line range
from /root/reference/frisk/__init__.py at run time and
``exec``-ed with three py2->py3 shims:

  * ``xrange = range``                      (F:228, 272, 327, 329)
  * ``np.NaN = np.nan``                     (F:483, 489, 494)
  * ``kmerMap[0].itervalues()``             (F:291) via a dict subclass

The whole module cannot be imported under py3.12 (tuple-parameter lambda at F:95,
missing hmmlearn/pybedtools/seaborn), but the twelve hot-path definitions can.
"""
from __future__ import annotations

import hashlib
import os
import types

REF_FILE = "/root/reference/frisk/__init__.py"
# sha256 of the reference file the golden fixtures were generated from
REF_SHA256 = "21fa89eb7588f3bd8bbd2f054131131666df4dfcbdfbe326ec88da6f6bca0e69"

# (first, last) 1-based inclusive line ranges of the hot-path definitions
_RANGES = [
    (70, 70),     # LETTERS
    (106, 118),   # countN
    (120, 137),   # calcGC
    (139, 164),   # iterFasta
    (194, 251),   # crawlGenome
    (253, 265),   # prepareMaps
    (267, 274),   # rangeMaps
    (276, 278),   # revComplement
    (280, 367),   # computeKmers
    (369, 457),   # IvomBuild
    (459, 472),   # KLD
    (474, 495),   # calcRIP
]


def available() -> bool:
    return os.path.isfile(REF_FILE)


class _PyTwoDict(dict):
    """dict with the py2 ``itervalues`` method used at F:291."""

    def itervalues(self):
        return iter(self.values())


def load() -> types.SimpleNamespace:
    """Return a namespace holding the reference's hot-path functions."""
    raw = open(REF_FILE, "rb").read()
    digest = hashlib.sha256(raw).hexdigest()
    if digest != REF_SHA256:
        raise RuntimeError("reference file drifted: sha256 %s" % digest)
    lines = raw.decode().split("\n")
    text = "from __future__ import division\n"
    for a, b in _RANGES:
        text += "\n".join(lines[a - 1:b]) + "\n\n"
    import copy, gzip, logging, math, pickle, sys
    from collections import Counter
    import numpy as np

    class _NP:  # minimal stand-in exposing the removed alias np.NaN
        NaN = np.nan

    glb = {
        "xrange": range, "np": _NP, "copy": copy, "gzip": gzip, "logging": logging,
        "math": math, "pickle": pickle, "sys": sys, "Counter": Counter,
        "__name__": "frisk_reference_hot_path",
    }
    exec(compile(text, REF_FILE + "<hot-path slices>", "exec"), glb)
    ns = types.SimpleNamespace(**{k: v for k, v in glb.items() if not k.startswith("__")})

    def blank_map(kmin, kmax):
        maps = ns.rangeMaps(kmin, kmax)
        maps[0] = _PyTwoDict(maps[0])
        return maps

    ns.blank_map = blank_map
    return ns


def make_args(hostSeq, querySeq=None, kmin=1, kmax=8, w=5000, i=2500, maskHost=False,
              scaffoldsAll=False, RIP=True, pcaMin=1, pcaMax=6):
    return types.SimpleNamespace(hostSeq=hostSeq, querySeq=querySeq, minWordSize=kmin,
                                 maxWordSize=kmax, windowlen=w, increment=i,
                                 maskHost=maskHost, scaffoldsAll=scaffoldsAll, RIP=RIP,
                                 pcaMin=pcaMin, pcaMax=pcaMax)


def run_hot_path(args, genomepickle="/dev/null", want_tables=False):
    """The reference's main() stages 2+3 (F:1442 and F:1478-1494) as data.

    Returns (genomeKmers, rows) with rows = [(name, start, stop, KLD, GC, PI, SI, CRI)].
    A window on which the reference raises ZeroDivisionError is reported as a row whose
    KLD is the string 'ZeroDivisionError' (the reference itself would abort there).
    """
    ref = load()
    blank = ref.blank_map(args.minWordSize, args.maxWordSize)
    query = args.querySeq or args.hostSeq
    genome = ref.computeKmers(args, genomepickle=genomepickle, window=None, genomeMode=True,
                              kmerMap=blank, getMeta=True)
    rows = []
    tables = []
    do_rip = args.RIP and args.minWordSize <= 2 <= args.maxWordSize
    for seq, name, start, stop in ref.crawlGenome(args, query):
        win = ref.computeKmers(args, genomepickle=None, window=[(name, seq)], genomeMode=False,
                               kmerMap=blank, getMeta=True)
        try:
            gi = ref.IvomBuild(win, args, genome, True)
            wi = ref.IvomBuild(win, args, genome, False)
            kld = ref.KLD(gi, wi, args)
        except ZeroDivisionError:
            kld = "ZeroDivisionError"
        try:
            gc = ref.calcGC(seq)
        except ZeroDivisionError:
            gc = "ZeroDivisionError"
        pi, si, cri = ref.calcRIP(win, args) if do_rip else (None, None, None)
        rows.append((name, start, stop, kld, gc, pi, si, cri))
        if want_tables:
            tables.append(win)
    if want_tables:
        return genome, rows, tables
    return genome, rows


# threshold helpers of the downstream stage: runnable unmodified with the xrange shim
_THRESH_RANGES = [
    (508, 513),   # FDBins
    (515, 543),   # otsu
    (664, 690),   # setKLDThresh
]


def load_thresholds() -> types.SimpleNamespace:
    """The reference's FDBins / otsu / setKLDThresh (F:508-543, F:664-690), executed from its source text."""
    raw = open(REF_FILE, "rb").read()
    if hashlib.sha256(raw).hexdigest() != REF_SHA256:
        raise RuntimeError("reference file drifted")
    lines = raw.decode().split("\n")
    text = "from __future__ import division\n"
    for a, b in _THRESH_RANGES:
        text += "\n".join(lines[a - 1:b]) + "\n\n"
    import logging, math
    import numpy as np
    glb = {"xrange": range, "np": np, "math": math, "logging": logging, "__name__": "frisk_reference_thresholds"}
    exec(compile(text, REF_FILE + "<threshold slices>", "exec"), glb)
    return types.SimpleNamespace(**{k: v for k, v in glb.items() if not k.startswith("__")})


# PCA features of anomalous regions: F:797-811 scrubMirrors runs unmodified; F:813-831
# flattenKmerMap needs py2's list-returning dict.keys() (``d.keys()[0]`` at F:819) and ``itertools``
_PCA_RANGES = [
    (797, 811),   # scrubMirrors
    (813, 831),   # flattenKmerMap
]


class _PyTwoListKeysDict(dict):
    """dict whose ``keys()`` returns a list, as in py2 (F:819 indexes it); ``itervalues`` as above."""

    def keys(self):
        return list(dict.keys(self))

    def itervalues(self):
        return iter(self.values())


def load_pca() -> types.SimpleNamespace:
    """The hot-path namespace of ``load()`` plus the reference's scrubMirrors / flattenKmerMap, executed from
    its source text.  Inside those two functions the name ``dict`` resolves to the py2-style subclass, so the
    ``dict()`` that scrubMirrors builds (F:806) answers ``keys()[0]`` in flattenKmerMap."""
    ns = load()
    raw = open(REF_FILE, "rb").read()
    lines = raw.decode().split("\n")
    text = "from __future__ import division\n"
    for a, b in _PCA_RANGES:
        text += "\n".join(lines[a - 1:b]) + "\n\n"
    import itertools
    import numpy as np
    glb = {"np": np, "itertools": itertools, "dict": _PyTwoListKeysDict, "revComplement": ns.revComplement,
           "__name__": "frisk_reference_pca_features"}
    exec(compile(text, REF_FILE + "<pca slices>", "exec"), glb)
    ns.scrubMirrors = glb["scrubMirrors"]
    ns.flattenKmerMap = glb["flattenKmerMap"]
    return ns


def run_pca_features(regions, pcaMin=1, pcaMax=6, windowlen=5000):
    """The reference's loop F:1571-1591 over ``regions`` = [(name, sequence str)]: per region
    computeKmers(pcaMode=True, sym=True, getMeta=False) -> scrubMirrors -> flattenKmerMap(prop=True).
    Returns a list of 1-D float arrays; a region on which the reference raises ZeroDivisionError (F:824, no
    valid word of some order) yields the string 'ZeroDivisionError'."""
    ref = load_pca()
    args = make_args(None, pcaMin=pcaMin, pcaMax=pcaMax, w=windowlen)
    blank = ref.rangeMaps(pcaMin, pcaMax)
    blank[0] = _PyTwoDict(blank[0])
    out = []
    for name, target in regions:
        count_map = ref.computeKmers(args, genomepickle=None, window=[(name, target)], genomeMode=False, pcaMode=True,
                                     kmerMap=blank, getMeta=False, sym=True)
        uniq = ref.scrubMirrors(count_map)
        try:
            out.append(ref.flattenKmerMap(uniq, window=windowlen, seqLen=len(target), kmin=pcaMin, kmax=pcaMax, prop=True))
        except ZeroDivisionError:
            out.append("ZeroDivisionError")
    return out
