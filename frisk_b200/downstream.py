"""Host-side glue between the window scores and the reference's later stages:
KLD thresholds (Freedman-Diaconis bin count, Otsu split, percentile), interval building from
thresholded windows, and the 2-state HMM segmentation with its GFF3 writer.

These are py3 restatements of the reference's functions with the same names, arguments and return
shapes -- including their quirks, which downstream numbers depend on -- except that the two
``pybedtools.BedTool`` objects the reference returns (F:664, F:789) are plain sorted lists of tuples
here (pybedtools / bedtools are not a dependency of this package); the merge done there by
``bedtools merge -d D -c 4,4,4 -o max,min,mean`` (F:667) is ``merge_intervals`` below.

Nothing here touches the GPU: the inputs are the DataFrame / arrays the hot path produced.
"""
from __future__ import annotations

import logging
import math
from typing import Iterable, Iterator, List, Sequence, Tuple

import numpy as np

from .api import FRISK_VERSION

__all__ = ["FDBins", "otsu", "setKLDThresh", "findBaseRanges", "range2interval", "hmm2BED", "hmmBED2GFF",
           "merge_intervals", "thresholdKLD", "anomaly2GFF", "GaussianHMM2", "fit_hmm"]


# ---------------------------------------------------------------------- thresholds
def FDBins(data) -> int:
    """F:508-513.  Named after Freedman-Diaconis, but what it returns -- and what every caller uses as a
    BIN COUNT -- is ``round(2 * IQR * n^(1/3))`` (the textbook rule divides by the cube root)."""
    q75, q25 = np.percentile(data, [75, 25])
    return int(round(2 * (q75 - q25) * math.pow(len(data), 1.0 / 3.0)))


def otsu(data, optBins: int) -> float:
    """F:515-543: Otsu's split of the histogram of log10(KLD).  The values are divided by
    ``-max|x|`` first (log10 KLDs are negative, so this maps them to positive numbers <= 1), the
    histogram is normalised by its tallest bin, and the class "means"/"variances" are taken over the
    normalised bin HEIGHTS (not over the values), exactly as the reference does.  Returns the left
    edge of the chosen bin mapped back to log10(KLD)."""
    raw = data
    x = np.atleast_1d(data)
    x = x[~np.isnan(x)]
    x = x / (max(abs(x)) * -1.0)
    hist, edges = np.histogram(x, bins=optBins)
    h = hist * 1.0
    h = h.ravel() / h.max()
    cum = h.cumsum()
    best, split = np.inf, -1
    for i in range(1, optBins):
        left, right = h[:i], h[i:]
        q1, q2 = cum[i - 1], cum[optBins - 1] - cum[i - 1]
        m1, m2 = q1 / len(left), q2 / len(right)
        v1 = np.sum(np.square(left - m1)) / len(left)
        v2 = np.sum(np.square(right - m2)) / len(right)
        cost = v1 * q1 + v2 * q2
        if cost < best:
            best, split = cost, i
    logging.info("OTSU selected bin %s as threshold position" % str(split))
    return edges[split] * max(abs(raw)) * -1.0


def setKLDThresh(args, logKLD):
    """F:664-690 -> (KLDthreshold, optBins).  At least 30 bins; ``--forceThresholdKLD`` wins over
    ``--threshTypeKLD`` (otsu | percentile).  Like the reference, raises UnboundLocalError when neither
    is given (its callers always pass one)."""
    opt = max(FDBins(logKLD), 30)
    if args.threshTypeKLD or args.forceThresholdKLD:
        if args.forceThresholdKLD:
            thr = np.log10(float(args.forceThresholdKLD))
            logging.info("Forcing log10(KLD) threshold = %s" % str(thr))
        elif args.threshTypeKLD == "otsu":
            logging.info("Calculating optimal KLD threshold by Otsu binarization.")
            if FDBins(logKLD) < 10:
                logging.warning("[WARNING] Low variance in log10(KLD) data: Review data distribution, "
                                "consider percentile or manual thresholding.")
            thr = otsu(logKLD, opt)
            logging.info("Optimal log10(KLD) threshold = %s" % str(thr))
        elif args.threshTypeKLD == "percentile":
            thr = np.percentile(logKLD, args.percentileKLD)
            logging.info("Setting threshold at %s percentile of log10(KLD)= %s" % (str(args.percentileKLD), str(thr)))
    return thr, opt


# ---------------------------------------------------------------------- intervals
def merge_intervals(records: Sequence[Tuple[str, int, int, float]], dist: int = 0) -> List[tuple]:
    """``bedtools merge -d dist -c 4,4,4 -o max,min,mean`` on (chrom, start, end, value) records:
    records are sorted by (chrom, start); a record joins the current block when it starts no further
    than ``dist`` past the block's end (book-ended records merge at dist = 0)."""
    out: List[tuple] = []
    block = None
    for chrom, start, end, val in sorted(records, key=lambda r: (r[0], r[1])):
        if block is not None and block[0] == chrom and start <= block[2] + dist:
            block[2] = max(block[2], end)
            block[3].append(val)
        else:
            if block is not None:
                out.append((block[0], block[1], block[2], max(block[3]), min(block[3]), sum(block[3]) / len(block[3])))
            block = [chrom, start, end, [val]]
    if block is not None:
        out.append((block[0], block[1], block[2], max(block[3]), min(block[3]), sum(block[3]) / len(block[3])))
    return out


def thresholdKLD(intervalList, threshold, args, threshCol="windowKLD", merge=True):
    """F:647-662 -> (anomalies, tItems): the windows whose log10(score) is on the anomalous side of
    the threshold (>= ; <= with --findSelf), optionally merged (--mergeDist).  ``anomalies`` is a list of
    (name, start, stop, KLD) or, merged, (name, start, stop, maxKLD, minKLD, meanKLD) tuples."""
    frame = intervalList.loc[~np.isnan(intervalList["windowKLD"])].sort_values(["name", "start", "stop"])
    score = np.log10(frame[threshCol])
    picked = frame.loc[score <= threshold] if args.findSelf else frame.loc[score >= threshold]
    picked = picked.copy()
    picked[["start", "stop"]] = picked[["start", "stop"]].astype(int)
    records = [(str(n), int(a), int(b), float(k)) for n, a, b, k in
               zip(picked["name"], picked["start"], picked["stop"], picked["windowKLD"])]
    anomalies = merge_intervals(records, args.mergeDist) if merge else records
    return anomalies, picked


def anomaly2GFF(anomBED, args, **kwargs) -> Iterator[str]:
    """F:553-567: GFF3 lines for thresholded windows ('windows' mode: KLD=) or merged features
    (maxKLD=, minKLD=, meanKLD=)."""
    kind = kwargs.get("category", "Kmer-anomaly")
    width = len(str(len(anomBED)))
    for n, rec in enumerate(anomBED, start=1):
        ident = "ID=Anomaly_" + str(n).zfill(width)
        if args.dimReduce == "windows":
            attrs = ";".join([ident, "KLD=" + str(rec[3])])
        else:
            attrs = ";".join([ident, "maxKLD=" + str(rec[3]), "minKLD=" + str(rec[4]), "meanKLD=" + str(rec[5])])
        if n == 1:
            yield "##gff-version 3\n"
        yield "\t".join([str(rec[0]), "frisk_" + FRISK_VERSION, kind, str(rec[1]), str(rec[2]), ".", "+", ".", attrs]) + "\n"


# ---------------------------------------------------------------------- HMM segmentation
def findBaseRanges(s, ch, name=None, minlen=0):
    """F:91-104: maximal runs of ``ch`` in ``s`` as inclusive (first, last) index pairs (prefixed by
    ``name`` when given); a run is kept when last - first >= minlen."""
    out = []
    first = prev = None
    for i, item in enumerate(s):
        if item == ch:
            if first is None:
                first = i
            prev = i
        elif first is not None:
            if prev - first >= minlen:
                out.append((name, first, prev) if name else (first, prev))
            first = None
    if first is not None and prev - first >= minlen:
        out.append((name, first, prev) if name else (first, prev))
    return out


def range2interval(rangeList, scaffoldWindows, state) -> Iterator[Tuple[str, str, str, str]]:
    """F:787-795: window-index runs -> (name, start of the first window, stop of the last, state)."""
    win = scaffoldWindows.reset_index(drop=True)
    for block in rangeList:
        yield (str(win["name"][0]), str(int(win["start"][block[0]])), str(int(win["stop"][block[1]])), str(state))


def hmm2BED(allWindows, model, dataCol="windowKLD"):
    """F:757-785: decode every scaffold's window scores with the fitted 2-state model, store the state
    per window in ``allWindows['hmmState']`` and return the runs of equal state as
    (name, start, stop, 'State1' | 'State2') tuples sorted like the reference sorts them (as strings)."""
    intervals: List[tuple] = []
    allWindows["hmmState"] = np.nan
    for name in dict.fromkeys(allWindows["name"]):              # scaffold names in order of first appearance
        rows = allWindows.loc[(allWindows["name"] == name) & ~np.isnan(allWindows["windowKLD"])]
        if len(rows) == 0:
            continue
        states = np.asarray(model.predict(rows[[dataCol]].to_numpy()))
        allWindows.loc[rows.index, "hmmState"] = states.astype(float)
        rows = rows.assign(hmmState=states.tolist())
        for value, label in ((0, "State1"), (1, "State2")):
            runs = findBaseRanges(states, value)
            if runs:
                intervals.extend(range2interval(runs, rows, label))
    intervals.sort(key=lambda r: (r[0], r[1], r[2]))
    return intervals, allWindows


def hmmBED2GFF(hmmBED) -> Iterator[str]:
    """F:589-596."""
    width = len(str(len(hmmBED)))
    for n, rec in enumerate(hmmBED, start=1):
        if n == 1:
            yield "##gff-version 3\n"
        yield "\t".join([rec[0], "frisk_" + FRISK_VERSION, str(rec[3]), str(rec[1]), str(rec[2]), ".", "+", ".",
                         "ID=" + rec[3] + "_" + str(n).zfill(width)]) + "\n"


class GaussianHMM2:
    """Deterministic 2-state, 1-D Gaussian HMM (Baum-Welch + Viterbi) used when ``hmmlearn`` is not
    installed.  Follows hmmlearn's GaussianHMM defaults where they are deterministic (uniform start and
    transition initialisation, min_covar 1e-3, n_iter 10, tol 1e-2) and replaces its unseeded k-means
    initialisation of the means by the means of the lower and upper half of the sorted data."""

    def __init__(self, n_iter: int = 10, tol: float = 1e-2, min_covar: float = 1e-3):
        self.n_iter, self.tol, self.min_covar = n_iter, tol, min_covar

    @staticmethod
    def _log_gauss(x, mean, var):
        return -0.5 * (np.log(2 * np.pi * var)[None, :] + (x[:, None] - mean[None, :]) ** 2 / var[None, :])

    def fit(self, X):
        """Baum-Welch in C (frisk_b200_hmm2_fit; a million observations in a fraction of a second)."""
        import ctypes as C
        from . import _lib
        x = np.ascontiguousarray(np.asarray(X, float).reshape(-1))
        start, trans, mean, var = np.zeros(2), np.zeros(4), np.zeros(2), np.zeros(2)
        p = lambda a: C.c_void_p(a.ctypes.data)
        _lib.check(_lib.lib().frisk_b200_hmm2_fit(p(x), len(x), self.n_iter, self.tol, self.min_covar, p(start), p(trans), p(mean),
                                                  p(var)), "frisk_b200_hmm2_fit")
        self.startprob_, self.transmat_, self.means_, self.vars_ = start, trans.reshape(2, 2), mean, var
        return self

    def predict(self, X):
        """Viterbi path in C (frisk_b200_hmm2_viterbi)."""
        import ctypes as C
        from . import _lib
        x = np.ascontiguousarray(np.asarray(X, float).reshape(-1))
        path = np.zeros(len(x), np.int32)
        if len(x) == 0:
            return path.astype(int)
        p = lambda a: C.c_void_p(a.ctypes.data)
        tr = np.ascontiguousarray(self.transmat_, float).reshape(-1)
        _lib.check(_lib.lib().frisk_b200_hmm2_viterbi(p(x), len(x), p(np.ascontiguousarray(self.startprob_, float)), p(tr),
                                                      p(np.ascontiguousarray(self.means_, float)),
                                                      p(np.ascontiguousarray(self.vars_, float)), p(path)), "frisk_b200_hmm2_viterbi")
        return path.astype(int)

    def fit_numpy(self, X):
        """The same algorithm in numpy (the cross-check of the C routine in tests/test_downstream.py)."""
        x = np.asarray(X, float).reshape(-1)
        order = np.sort(x)
        half = len(x) // 2
        mean = np.array([order[:half].mean(), order[half:].mean()])
        var = np.full(2, x.var() + self.min_covar)
        start = np.full(2, 0.5)
        trans = np.full((2, 2), 0.5)
        prev = -np.inf
        n = len(x)
        for _ in range(self.n_iter):
            logb = self._log_gauss(x, mean, var)
            la = np.zeros((n, 2))
            lb = np.zeros((n, 2))
            lt = np.log(trans)
            la[0] = np.log(start) + logb[0]
            for t in range(1, n):
                la[t] = logb[t] + np.logaddexp(la[t - 1, 0] + lt[0], la[t - 1, 1] + lt[1])
            for t in range(n - 2, -1, -1):
                lb[t] = np.logaddexp(lt[:, 0] + logb[t + 1, 0] + lb[t + 1, 0], lt[:, 1] + logb[t + 1, 1] + lb[t + 1, 1])
            ll = np.logaddexp(la[-1, 0], la[-1, 1])
            gamma = np.exp(la + lb - ll)
            xi = np.exp(la[:-1, :, None] + lt[None] + (logb[1:] + lb[1:])[:, None, :] - ll).sum(0)
            start = gamma[0] / gamma[0].sum()
            trans = xi / xi.sum(1, keepdims=True)
            w = gamma.sum(0)
            mean = (gamma * x[:, None]).sum(0) / w
            var = (gamma * (x[:, None] - mean[None]) ** 2).sum(0) / w + self.min_covar
            if ll - prev < self.tol:
                break
            prev = ll
        self.startprob_, self.transmat_, self.means_, self.vars_ = start, trans, mean, var
        return self

    def predict_numpy(self, X):
        x = np.asarray(X, float).reshape(-1)
        logb = self._log_gauss(x, self.means_, self.vars_)
        lt = np.log(self.transmat_)
        n = len(x)
        delta = np.log(self.startprob_) + logb[0]
        back = np.zeros((n, 2), int)
        for t in range(1, n):
            cand = delta[:, None] + lt
            back[t] = cand.argmax(0)
            delta = cand.max(0) + logb[t]
        path = np.zeros(n, int)
        path[-1] = int(delta.argmax())
        for t in range(n - 1, 0, -1):
            path[t - 1] = back[t, path[t]]
        return path


def fit_hmm(allKLD):
    """F:1539-1541: the 2-state model fitted on every non-NaN window KLD as one sequence.  hmmlearn's
    GaussianHMM(n_components=2, covariance_type="full") when it is installed, GaussianHMM2 otherwise."""
    data = np.asarray(allKLD, float).reshape(-1)
    data = data[~np.isnan(data)][:, np.newaxis]
    try:
        from hmmlearn import hmm
        model = hmm.GaussianHMM(n_components=2, covariance_type="full")
    except ImportError:
        model = GaussianHMM2()
    model.fit(data)
    return model
