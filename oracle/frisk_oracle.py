"""CPU oracle (Python) for frisk's hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A from-scratch py3 restatement of the algorithm in /root/reference/frisk/__init__.py
("F:" below), written to be read side by side with it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module; the product package ``frisk_b200`` never does (it fails loudly without its
CUDA library instead).

PARITY PIN.  The reference has no tests or golden vectors of its own (SURVEY.md
section 4), so this oracle is pinned against the reference's *own source text executed
in the build container* (``oracle/ref_exec.py``): ``tests/golden/make_golden.py``
wrote the committed fixtures under ``tests/golden/`` from that execution, and
``tests/test_oracle.py`` checks this file against them (integer tables bit-exact,
floats to 1e-12).  Data structures deliberately match the reference (lists of
``dict[str,int]`` in A,T,G,C order, meta dicts appended) so that its timing is
representative of the reference's CPU cost.

Deliberate deviations from the reference, all behaviour-preserving:
  * py3 syntax (``range``, ``dict.values``); true division everywhere (F:22 already
    imports it).
  * ``compute_kmers`` takes the scaffold iterator as an argument instead of opening
    ``args.hostSeq`` itself and does not pickle (F:297, F:363 are I/O, not arithmetic).
"""
from __future__ import annotations

import math
from collections import Counter
from typing import Dict, Iterable, Iterator, List, Sequence, Tuple

LETTERS = ("A", "T", "G", "C")          # F:70 -- alphabet AND table order
_COMPLEMENT = {"A": "T", "T": "A", "G": "C", "C": "G"}   # F:277


# --------------------------------------------------------------------------- ingest
def iter_fasta(path: str) -> Iterator[Tuple[str, str]]:
    """F:139-164.  Name = first whitespace token after stripping '>' characters
    (F:156); lines are stripped, blank lines skipped (F:149-151).  Plain text only."""
    name, chunks = None, []
    with open(path) as handle:
        for raw in handle:
            line = raw.strip()
            if not line:
                continue
            if line.startswith(">"):
                if name:
                    yield name, "".join(chunks)
                name = line.strip(">").split()[0]
                chunks = []
            else:
                chunks.append(line)
    if name:
        yield name, "".join(chunks)


def count_n(sequence: str) -> Tuple[int, int]:
    """F:106-118: (#chars in upper-case ATGC, #everything else).  Case-sensitive."""
    tally = Counter(sequence)
    good = sum(tally[c] for c in LETTERS)
    return good, len(sequence) - good


def calc_gc(sequence: str) -> float:
    """F:120-137: (G+C)/(A+T+G+C) over upper-case bases only; ZeroDivisionError if none."""
    tally = Counter(sequence)
    gc = tally["G"] + tally["C"]
    at = tally["A"] + tally["T"]
    return float(gc) / (gc + at)


def crawl_genome(scaffolds: Iterable[Tuple[str, str]], w: int, step: int,
                 scaffolds_all: bool = False) -> Iterator[Tuple[str, str, int, int]]:
    """F:194-251: window enumeration.  Yields (window_seq, name, start, stop).

    * minimum size rule ``size <= w + (0.75 w - step)`` (F:211, F:222): skip, or with
      ``scaffolds_all`` emit the whole scaffold as window (1, size) unless >= 30 % of it
      is not upper-case ATGC (F:213).
    * windows at j = 0, step, 2*step, ... while j <= size - step (F:228); a window that
      would overshoot is replaced by the last w bases and, from then on (the flag is never
      cleared, F:232), coordinates are reported as (size - w, size) (F:243).
    * a window is dropped when its non-ATGC count >= 0.3 * len (F:238).
    """
    for name, seq in scaffolds:
        size = len(seq)
        jumped = False
        small = size <= w + ((w * 0.75) - step)
        if small and scaffolds_all:
            if count_n(seq)[1] >= 0.3 * size:
                continue
            yield seq, name, 1, size
        elif small:
            continue
        else:
            for j in range(0, size - step + 1, step):
                if j + w > size:
                    win = seq[size - w:size]
                    jumped = True
                else:
                    win = seq[j:j + w]
                if count_n(win)[1] >= 0.3 * len(win):
                    continue
                if jumped:
                    yield win, name, size - w, size
                else:
                    yield win, name, j + 1, j + w


# --------------------------------------------------------------------------- tables
def range_maps(kmin: int, kmax: int) -> List[Dict[str, int]]:
    """F:253-274: one zeroed dict per order, keys in A,T,G,C-lexicographic order."""
    maps = []
    for k in range(kmin, kmax + 1):
        words = [""]
        for _ in range(k):
            words = [wd + c for wd in words for c in LETTERS]
        maps.append(dict.fromkeys(words, 0))
    return maps


def rev_complement(kmer: str) -> str:
    """F:276-278."""
    return "".join(_COMPLEMENT[b] for b in reversed(kmer))


def compute_kmers(scaffolds: Iterable[Tuple[str, str]], kmin: int, kmax: int,
                  both_strands: bool, force_upper: bool = True) -> list:
    """F:280-367.  ``both_strands`` = ``genomeMode or sym`` (F:350); ``force_upper`` is
    False only for the genome pass under ``--maskHost`` (F:334-337).

    Returns [table_kmin, ..., table_kmax, {'totalLen'}, {'exMax'}, {'nnTotal'}]
    (F:356-359).  A word containing anything outside ATGC is skipped, and tallied in
    exMax when it is a kmax-word (F:341-346)."""
    maps = range_maps(kmin, kmax)
    total_len = ex_max = nn_total = 0
    for _name, seq in scaffolds:
        size = len(seq)
        total_len += size                      # F:323
        nn_total += count_n(seq)[1]            # F:324-325 (case-sensitive)
        text = seq.upper() if force_upper else seq
        for k in range(kmin, kmax + 1):        # F:327
            table = maps[k - kmin]
            for j in range(size - k + 1):      # F:329
                word = text[j:j + k]
                if word not in table:          # F:341-346
                    if k == kmax:
                        ex_max += 1
                    continue
                table[word] += 1               # F:348
                if both_strands:               # F:350-351 (palindromes get +2)
                    table[rev_complement(word)] += 1
    maps.append({"totalLen": total_len})
    maps.append({"exMax": ex_max})
    maps.append({"nnTotal": nn_total})
    return maps


# --------------------------------------------------------------------------- scoring
def ivom_build(window_kmers: list, genome_kmers: list, kmin: int, kmax: int,
               is_genome_ivom: bool) -> Dict[str, float]:
    """F:369-457.  For every kmax-mer present in the window (F:386-389):

      w_x = C_x(prefix_x) * 4**x                      (F:399-424)
      p_x = C_x(prefix_x) / ((S - (x-1)) * 2)
      a_x = w_x / sum_{y<=x} w_y                      (F:426-437)
      I_kmin = a*p ;  I_x = a_x p_x + (1-a_x) I_{x-1}  (F:439-446)

    with (C, S) = (window tables, window space) or (genome tables, genome space),
    S = totalLen - nnTotal (F:379-380); finally normalised to sum 1 (F:453-454)."""
    kr = kmax - kmin
    genome_space = genome_kmers[kr + 1]["totalLen"] - genome_kmers[kr + 3]["nnTotal"]
    window_space = window_kmers[kr + 1]["totalLen"] - window_kmers[kr + 3]["nnTotal"]
    tables, space = (genome_kmers, genome_space) if is_genome_ivom else (window_kmers, window_space)
    raw: Dict[str, float] = {}
    total = 0
    for word, count in window_kmers[kr].items():
        if count == 0:
            continue
        ivom = 0.0
        running = 0
        for x in range(kmin, kmax + 1):
            c = tables[x - kmin][word[:x]]
            weight = c * 4 ** x
            prob = float(c) / ((space - (x - 1)) * 2)
            running += weight
            alpha = float(weight) / running
            if x == kmin:
                ivom = alpha * prob
            else:
                ivom = alpha * prob + ((1 - alpha) * ivom)
        raw[word] = ivom
        total += ivom
    return {word: float(v) / total for word, v in raw.items()}


def kld(genome_ivom: Dict[str, float], window_ivom: Dict[str, float]) -> float:
    """F:459-472: sum_k w*log2(w/G), terms with G == 0 skipped.  Returns int 0 for an
    empty window (as the reference does)."""
    score = 0
    for word, wv in window_ivom.items():
        g = float(genome_ivom[word])
        wv = float(wv)
        if g != 0:
            score += wv * math.log(wv / g, 2)
    return score


def calc_rip(window_kmers: list, kmin: int, kmax: int) -> Tuple[float, float, float]:
    """F:474-495.  ValueError when 2 is outside [kmin, kmax] (F:478)."""
    di = window_kmers[list(range(kmin, kmax + 1)).index(2)]
    nan = float("nan")
    pi = di["TA"] / float(di["AT"]) if di["AT"] > 0 else nan
    sub = di["AC"] + di["GT"]
    si = (di["CA"] + di["TG"]) / float(sub) if sub > 0 else nan
    cri = pi - si if (pi and si) else nan      # F:491: 0.0 is falsy, NaN is truthy
    return pi, si, cri


# --------------------------------------------------------------------------- driver
def score_windows(scaffolds: Sequence[Tuple[str, str]], kmin: int = 1, kmax: int = 8, w: int = 5000,
                  step: int = 2500, mask_host: bool = False, scaffolds_all: bool = False,
                  rip: bool = True, host: Sequence[Tuple[str, str]] | None = None):
    """Stages 2+3 of the reference's main(): F:1442 then the loop F:1478-1494.

    Returns (genome_kmers, rows) with rows = (name, start, stop, KLD, GC, PI, SI, CRI);
    PI/SI/CRI are None when RIP is off.  A window on which the reference would raise
    ZeroDivisionError yields the string 'ZeroDivisionError' in place of the value."""
    genome = compute_kmers(host if host is not None else scaffolds, kmin, kmax,
                           both_strands=True, force_upper=not mask_host)
    rows = []
    do_rip = rip and kmin <= 2 <= kmax
    for seq, name, start, stop in crawl_genome(scaffolds, w, step, scaffolds_all):
        win = compute_kmers([(name, seq)], kmin, kmax, both_strands=False)
        try:
            score = kld(ivom_build(win, genome, kmin, kmax, True),
                        ivom_build(win, genome, kmin, kmax, False))
        except ZeroDivisionError:
            score = "ZeroDivisionError"
        try:
            gc = calc_gc(seq)
        except ZeroDivisionError:
            gc = "ZeroDivisionError"
        pi, si, cri = calc_rip(win, kmin, kmax) if do_rip else (None, None, None)
        rows.append((name, start, stop, score, gc, pi, si, cri))
    return genome, rows


# --------------------------------------------------------------------------- PCA features (F:797-831, F:1571-1591)
def scrub_mirrors(k_dicts: list) -> list:
    """F:797-811: keep, per order, the first of each {k-mer, reverse complement} pair in table order."""
    out = []
    for table in k_dicts:
        kept: Dict[str, int] = {}
        for key in table:
            if key in kept or rev_complement(key) in kept:
                continue
            kept[key] = table[key]
        out.append(kept)
    return out


def flatten_props(k_dicts: list) -> list:
    """F:813-831 with prop=True: every order's values divided by the order's sum, concatenated."""
    vec = []
    for table in k_dicts:
        total = sum(table.values())
        vec.extend(float(v) / total for v in table.values())
    return vec


def region_features(seq: str, kmin: int, kmax: int) -> list:
    """The reference's per-region feature vector (F:1576-1584): symmetric counts, mirrors scrubbed, proportions."""
    maps = compute_kmers([("r", seq)], kmin, kmax, both_strands=True)[:kmax - kmin + 1]
    return flatten_props(scrub_mirrors(maps))
