"""Generate the committed golden fixtures by EXECUTING THE REFERENCE'S OWN SOURCE TEXT.

Container-only (needs /root/reference; see oracle/ref_exec.py).  Run from the repo root:

    python tests/golden/make_golden.py [--full-c1] [--jobs 8]

Each case is a seeded synthetic genome (frisk_b200/synth.py) + run parameters; outputs
are the reference's genome tables (uint64, orders kmin..kmax concatenated in the
reference's A,T,G,C dict order), its three meta counters, every emitted window row, and
for a few windows the full window tables.  Stored as tests/golden/<case>.npz.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from frisk_b200 import synth  # noqa: E402
from oracle import ref_exec  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (synth config, scale, params, optional host config)
CASES = {
    "edge_default": dict(genome=("edge", 1.0), params={}),
    "edge_scaffoldsAll": dict(genome=("edge", 1.0), params=dict(scaffoldsAll=True)),
    "edge_maskHost": dict(genome=("edge", 1.0), params=dict(maskHost=True)),
    "edge_k2_5_w1000_i250": dict(genome=("edge", 1.0), params=dict(kmin=2, kmax=5, w=1000, i=250)),
    "edge_k1_3_w3000_i1000": dict(genome=("edge", 1.0), params=dict(kmin=1, kmax=3, w=3000, i=1000, scaffoldsAll=True)),
    "edge_k4_8": dict(genome=("edge", 1.0), params=dict(kmin=4, kmax=8)),
    "edge_k1_1": dict(genome=("edge", 1.0), params=dict(kmin=1, kmax=1)),
    # beyond the default --maxWordSize 8: served by the general (global-memory) kernels
    "edge_k1_9": dict(genome=("edge", 1.0), params=dict(kmin=1, kmax=9)),
    "edge_k9_10_w2500_i2500": dict(genome=("edge", 1.0), params=dict(kmin=9, kmax=10, w=2500, i=2500, scaffoldsAll=True)),
    # 0.75 w < step < w: scaffolds shorter than the window are windowed through a negative slice start (F:231)
    "short_w1000_i800": dict(genome=("edge_short", 1.0), params=dict(w=1000, i=800)),
    "short_w1000_i800_all": dict(genome=("edge_short", 1.0), params=dict(w=1000, i=800, scaffoldsAll=True, kmin=2, kmax=6)),
    "c1_small": dict(genome=("C1", 0.04), params={}),
    "c2_small": dict(genome=("C2", 0.01), params={}),
    "c2_small_query_vs_c1_host": dict(genome=("C2", 0.005), host=("C1", 0.02), params={}),
    "c1_full": dict(genome=("C1", 1.0), params={}),
}
DEFAULT = [c for c in CASES if c != "c1_full"]


def tables_to_array(maps, kmin, kmax):
    parts = [np.fromiter(maps[k - kmin].values(), dtype=np.uint64) for k in range(kmin, kmax + 1)]
    return np.concatenate(parts)


def _score_chunk(job):
    """Worker: score a contiguous chunk of windows with the reference functions."""
    (args, genome, chunk) = job
    ref = ref_exec.load()
    blank = ref.blank_map(args.minWordSize, args.maxWordSize)
    do_rip = args.RIP and args.minWordSize <= 2 <= args.maxWordSize
    out = []
    for seq, name, start, stop in chunk:
        win = ref.computeKmers(args, genomepickle=None, window=[(name, seq)], genomeMode=False,
                               kmerMap=blank, getMeta=True)
        try:
            kld = ref.KLD(ref.IvomBuild(win, args, genome, True), ref.IvomBuild(win, args, genome, False), args)
        except ZeroDivisionError:
            kld = float("inf")          # sentinel: the reference raised ZeroDivisionError
        try:
            gc = ref.calcGC(seq)
        except ZeroDivisionError:
            gc = float("inf")
        pi, si, cri = ref.calcRIP(win, args) if do_rip else (np.nan, np.nan, np.nan)
        out.append((name, start, stop, float(kld), float(gc), float(pi), float(si), float(cri)))
    return out


def run_case(case, jobs):
    spec = CASES[case]
    cfg, scale = spec["genome"]
    scaffolds = synth.make(cfg, scale)
    p = dict(kmin=1, kmax=8, w=5000, i=2500, maskHost=False, scaffoldsAll=False, RIP=True)
    p.update(spec["params"])
    tmp = tempfile.mkdtemp(prefix="golden_")
    qpath = os.path.join(tmp, "query.fa")
    synth.write_fasta(scaffolds, qpath)
    hpath = qpath
    host_digest = ""
    if "host" in spec:
        host = synth.make(*spec["host"])
        hpath = os.path.join(tmp, "host.fa")
        synth.write_fasta(host, hpath)
        host_digest = synth.digest(host)
    args = ref_exec.make_args(hpath, querySeq=qpath, **p)
    ref = ref_exec.load()
    t0 = time.time()
    blank = ref.blank_map(p["kmin"], p["kmax"])
    genome = ref.computeKmers(args, genomepickle=os.path.join(tmp, "g.p"), window=None, genomeMode=True,
                              kmerMap=blank, getMeta=True)
    t_bg = time.time() - t0
    windows = list(ref.crawlGenome(args, qpath))
    t0 = time.time()
    if jobs > 1 and len(windows) > 4 * jobs:
        per = (len(windows) + jobs * 4 - 1) // (jobs * 4)
        chunks = [windows[a:a + per] for a in range(0, len(windows), per)]
        with mp.Pool(jobs) as pool:
            parts = pool.map(_score_chunk, [(args, genome, c) for c in chunks])
        rows = [r for part in parts for r in part]
    else:
        rows = _score_chunk((args, genome, windows))
    t_win = time.time() - t0
    # full window tables for up to 4 windows spread over the list
    pick = sorted(set(int(x) for x in np.linspace(0, len(windows) - 1, min(4, len(windows))))) if windows else []
    wtabs = []
    wmeta = []
    for idx in pick:
        seq, name, start, stop = windows[idx]
        win = ref.computeKmers(args, genomepickle=None, window=[(name, seq)], genomeMode=False,
                               kmerMap=blank, getMeta=True)
        wtabs.append(tables_to_array(win, p["kmin"], p["kmax"]).astype(np.uint32))
        kr = p["kmax"] - p["kmin"]
        wmeta.append([win[kr + 1]["totalLen"], win[kr + 2]["exMax"], win[kr + 3]["nnTotal"]])
    kr = p["kmax"] - p["kmin"]
    meta = dict(case=case, genome=[cfg, scale], host=list(spec.get("host", [])), params=p,
                genome_digest=synth.digest(scaffolds), host_digest=host_digest,
                ref_sha256=ref_exec.REF_SHA256, n_rows=len(rows), seconds_background=t_bg,
                seconds_windows=t_win, jobs=jobs,
                total_bases=synth.total_bases(scaffolds))
    np.savez_compressed(
        os.path.join(HERE, case + ".npz"),
        meta=np.array(json.dumps(meta)),
        genome_tables=tables_to_array(genome, p["kmin"], p["kmax"]),
        genome_meta=np.array([genome[kr + 1]["totalLen"], genome[kr + 2]["exMax"], genome[kr + 3]["nnTotal"]],
                             dtype=np.uint64),
        row_names=np.array([r[0] for r in rows]),
        row_coords=np.array([[r[1], r[2]] for r in rows], dtype=np.int64).reshape(-1, 2),
        row_vals=np.array([r[3:] for r in rows], dtype=np.float64).reshape(-1, 5),
        win_pick=np.array(pick, dtype=np.int64),
        win_tables=np.array(wtabs, dtype=np.uint32).reshape(len(pick), -1),
        win_meta=np.array(wmeta, dtype=np.uint64).reshape(len(pick), 3),
    )
    print("%-28s rows=%-5d bases=%-8d background %.1fs windows %.1fs" %
          (case, len(rows), meta["total_bases"], t_bg, t_win), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*")
    ap.add_argument("--full-c1", action="store_true")
    ap.add_argument("--jobs", type=int, default=8)
    a = ap.parse_args()
    cases = a.cases or list(DEFAULT)
    if a.full_c1 and "c1_full" not in cases:
        cases.append("c1_full")
    for c in cases:
        run_case(c, a.jobs)


if __name__ == "__main__":
    main()
