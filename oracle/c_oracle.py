"""ctypes wrapper around oracle/libfrisk_oracle.so (the C oracle) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ERR_KLD_ZERODIV, ERR_GC_ZERODIV, ERR_LOG_DOMAIN = 1, 2, 4


def build(force: bool = False) -> str:
    so = os.path.join(HERE, "libfrisk_oracle.so")
    src = os.path.join(HERE, "frisk_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "libfrisk_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p, u64p, u32p, i64p, f64p = (C.c_void_p,) * 5
        L.frisk_oracle_table_size.restype = C.c_size_t
        L.frisk_oracle_table_size.argtypes = [C.c_int, C.c_int]
        L.frisk_oracle_count.restype = None
        L.frisk_oracle_count.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, u64p, u64p]
        L.frisk_oracle_window.restype = C.c_uint
        L.frisk_oracle_window.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, u64p, u64p, C.c_int, f64p, u64p, u64p]
        L.frisk_oracle_crawl.restype = C.c_size_t
        L.frisk_oracle_crawl.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, u64p, u32p, i64p, i64p, C.c_size_t]
        L.frisk_oracle_background.restype = None
        L.frisk_oracle_background.argtypes = [u8p, u64p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, u64p, u64p]
        L.frisk_oracle_score.restype = None
        L.frisk_oracle_score.argtypes = [u8p, u64p, u32p, C.c_size_t, C.c_int, C.c_int, C.c_int, u64p, u64p, C.c_int,
                                         f64p, u32p]
        _LIB = L
    return _LIB


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def table_size(kmin: int, kmax: int) -> int:
    return int(lib().frisk_oracle_table_size(kmin, kmax))


def concat(scaffolds: Sequence[Tuple[str, np.ndarray]]):
    """-> (uint8 concatenation, uint64 offsets[n+1])."""
    lens = np.array([len(s) for _, s in scaffolds], dtype=np.uint64)
    off = np.zeros(len(scaffolds) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    seq = np.concatenate([np.ascontiguousarray(s, dtype=np.uint8) for _, s in scaffolds]) if scaffolds else np.zeros(0, np.uint8)
    return np.ascontiguousarray(seq), off


def background(seq: np.ndarray, off: np.ndarray, kmin=1, kmax=8, mask_host=False, threads=1):
    tabs = np.zeros(table_size(kmin, kmax), dtype=np.uint64)
    meta = np.zeros(3, dtype=np.uint64)
    lib().frisk_oracle_background(_ptr(seq), _ptr(off), len(off) - 1, kmin, kmax, int(mask_host), threads,
                                  _ptr(tabs), _ptr(meta))
    return tabs, meta


def crawl(seq: np.ndarray, off: np.ndarray, w=5000, step=2500, scaffolds_all=False):
    """Window list over all scaffolds: (scaffold index, absolute offset, length, start, stop)."""
    L = lib()
    sidx, woff, wlen, st, sp = [], [], [], [], []
    for s in range(len(off) - 1):
        a, b = int(off[s]), int(off[s + 1])
        size = b - a
        cap = size // max(step, 1) + 2
        o = np.zeros(cap, np.uint64); l = np.zeros(cap, np.uint32)
        x = np.zeros(cap, np.int64); y = np.zeros(cap, np.int64)
        sub = seq[a:b]
        n = int(L.frisk_oracle_crawl(_ptr(sub), size, w, step, int(scaffolds_all), _ptr(o), _ptr(l), _ptr(x), _ptr(y), cap))
        assert n <= cap
        sidx.append(np.full(n, s, np.int64)); woff.append(o[:n] + np.uint64(a)); wlen.append(l[:n])
        st.append(x[:n]); sp.append(y[:n])
    cat = lambda parts, dt: np.concatenate(parts).astype(dt) if parts else np.zeros(0, dt)
    return cat(sidx, np.int64), cat(woff, np.uint64), cat(wlen, np.uint32), cat(st, np.int64), cat(sp, np.int64)


def score(seq, woff, wlen, gtabs, gmeta, kmin=1, kmax=8, rip=True, threads=1):
    n = len(woff)
    rows = np.zeros((n, 5), dtype=np.float64)
    status = np.zeros(n, dtype=np.uint32)
    woff = np.ascontiguousarray(woff, np.uint64); wlen = np.ascontiguousarray(wlen, np.uint32)
    lib().frisk_oracle_score(_ptr(seq), _ptr(woff), _ptr(wlen), n, kmin, kmax, int(rip), _ptr(gtabs), _ptr(gmeta),
                             threads, _ptr(rows), _ptr(status))
    return rows, status


def window_tables(win: np.ndarray, gtabs, gmeta, kmin=1, kmax=8, rip=True):
    """One window: (row[5], status, window tables uint64, window meta[3])."""
    win = np.ascontiguousarray(win, np.uint8)
    out = np.zeros(5, np.float64)
    wt = np.zeros(table_size(kmin, kmax), np.uint64)
    wm = np.zeros(3, np.uint64)
    st = lib().frisk_oracle_window(_ptr(win), len(win), kmin, kmax, _ptr(gtabs), _ptr(gmeta), int(rip), _ptr(out),
                                   _ptr(wt), _ptr(wm))
    return out, int(st), wt, wm


def run(scaffolds, kmin=1, kmax=8, w=5000, step=2500, mask_host=False, scaffolds_all=False, rip=True,
        host=None, threads=1):
    """Whole hot path.  Returns dict(tables, meta, names, coords, rows, status)."""
    seq, off = concat(scaffolds)
    if host is not None:
        hseq, hoff = concat(host)
        tabs, meta = background(hseq, hoff, kmin, kmax, mask_host, threads)
    else:
        tabs, meta = background(seq, off, kmin, kmax, mask_host, threads)
    sidx, woff, wlen, st, sp = crawl(seq, off, w, step, scaffolds_all)
    rows, status = score(seq, woff, wlen, tabs, meta, kmin, kmax, rip, threads)
    if not (rip and kmin <= 2 <= kmax):
        rows[:, 2:] = np.nan
    names = [scaffolds[i][0] for i in sidx]
    return dict(tables=tabs, meta=meta, names=names, coords=np.stack([st, sp], 1) if len(st) else np.zeros((0, 2), np.int64),
                rows=rows, status=status, scaffold_index=sidx, win_off=woff, win_len=wlen)
