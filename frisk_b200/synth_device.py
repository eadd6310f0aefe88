"""Device-drawn stand-ins for the two largest BASELINE configs (C4: 3 Gbp in 24 chromosome-sized scaffolds with Mbp-scale N
runs; C5: 14 Gbp in 1,000,000 short scaffolds with many N gaps).  SURVEY.md section 8(d) specifies numpy generators for
them (frisk_b200/synth.py: config_c4 / config_c5, used at reduced scale in the CPU tests); at FULL size those take
minutes of host time and tens of GB of host memory per process, so the full-size tests, tools/c4_scaling.py and bench.py's
`strong` / `c5` blocks draw the genomes directly in packed form on the GPU instead: uniform random codes (optionally with
AT-rich blocks), the scaffold-length distributions and N-run placement of the spec, invalid padding between scaffolds.
Torch is only the random source and the buffer owner.  These are stand-ins with the spec's SIZES and STRUCTURE, not its
exact base streams; parity on them is checked through size-independent properties plus windows decoded from the planes
and scored by the C oracle (tests/test_scale_gpu.py)."""
from __future__ import annotations

import ctypes as C

import numpy as np

LETTERS = np.frombuffer(b"ATGC", dtype=np.uint8)      # code order of the reference's tables (F:70)


def _ranges_to_word_masks(starts, ends):
    """Bit ranges [s, e) of a 1-bit-per-base plane (bit 31 = first base of a word) -> the distinct
    partially covered words with their OR-ed masks, and the fully covered word runs [w0, w1)."""
    starts = np.asarray(starts, np.int64); ends = np.asarray(ends, np.int64)
    keep = ends > starts
    starts, ends = starts[keep], ends[keep]
    fw, lw = starts >> 5, (ends - 1) >> 5
    full = np.uint64(0xFFFFFFFF)
    head = (full >> (starts & 31).astype(np.uint64)).astype(np.uint64)
    tail = (full << (31 - ((ends - 1) & 31)).astype(np.uint64)).astype(np.uint64) & full
    same = fw == lw
    idx = np.concatenate([fw[same], fw[~same], lw[~same]])
    msk = np.concatenate([head[same] & tail[same], head[~same], tail[~same]])
    uniq, inverse = np.unique(idx, return_inverse=True)
    merged = np.zeros(len(uniq), np.uint64)
    np.bitwise_or.at(merged, inverse, msk)
    return uniq, merged.astype(np.uint32), fw[~same] + 1, lw[~same]


def build_device_genome(eng, scaf_len, n_runs, seed, at_rich_block=0, device="cuda:0"):
    """A random genome straight into device planes: uniform codes (optionally AT-rich blocks), N runs
    (scaffold[], offset[], length[]; non-overlapping), invalid padding between scaffolds."""
    import torch
    from frisk_b200 import _lib
    L = _lib.lib()
    dev = torch.device(device)
    scaf_len = np.ascontiguousarray(scaf_len, np.uint64)
    n = len(scaf_len)
    scaf_off = np.zeros(n, np.uint64)
    padded = C.c_uint64(0)
    _lib.check(L.frisk_b200_pack_layout(eng._ptr(scaf_len), n, eng._ptr(scaf_off), C.byref(padded)), "layout")
    P = int(padded.value)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    codes = torch.randint(-2 ** 63, 2 ** 63 - 1, (P // 32,), dtype=torch.int64, device=dev, generator=gen).view(torch.int32)
    if at_rich_block:
        # every other block of `at_rich_block` code words: P(G or C) = 1/4 (high bit of a code = G/C)
        r = torch.randint(-2 ** 63, 2 ** 63 - 1, (P // 32,), dtype=torch.int64, device=dev, generator=gen).view(torch.int32)
        blk = (torch.arange(P // 16, device=dev, dtype=torch.int64) // at_rich_block) & 1
        lowgc = codes & (r | 0x55555555)
        codes = torch.where(blk.bool(), lowgc, codes)
        del r, blk, lowgc
    inv = torch.zeros(P // 32, dtype=torch.int32, device=dev)
    so, sl = scaf_off.astype(np.int64), scaf_len.astype(np.int64)
    pad_s = so + sl
    pad_e = np.concatenate([so[1:], [P]])
    run_s, run_o, run_l = (np.asarray(x, np.int64) for x in n_runs)
    rs = so[run_s] + run_o
    re = rs + run_l
    uniq, masks, f0, f1 = _ranges_to_word_masks(np.concatenate([pad_s, rs]), np.concatenate([pad_e, re]))
    t_idx = torch.from_numpy(uniq).to(dev)
    inv[t_idx] = inv[t_idx] | torch.from_numpy(masks.view(np.int32)).to(dev)
    # fully covered words: +1 / -1 difference array over words, prefix sum > 0
    keep = f1 > f0
    if keep.any():
        delta = torch.zeros(P // 32 + 1, dtype=torch.int32, device=dev)
        ones = torch.ones(int(keep.sum()), dtype=torch.int32, device=dev)
        delta.index_add_(0, torch.from_numpy(f0[keep]).to(dev), ones)
        delta.index_add_(0, torch.from_numpy(f1[keep]).to(dev), -ones)
        inv = torch.where(torch.cumsum(delta[:-1], 0, dtype=torch.int32) > 0, torch.full_like(inv, -1), inv)
        del delta
    # invalid bases carry code 0 (plane convention): clear the codes under the mask, word-wise
    # (expand each mask bit to the two code bits of its base)
    m = inv.to(torch.int64) & 0xFFFFFFFF
    def spread16(x):                       # 16 mask bits -> 32 bits, each bit doubled
        x = (x | (x << 8)) & 0x00FF00FF
        x = (x | (x << 4)) & 0x0F0F0F0F
        x = (x | (x << 2)) & 0x33333333
        x = (x | (x << 1)) & 0x55555555
        return x | (x << 1)
    hi, lo = spread16(m >> 16), spread16(m & 0xFFFF)
    kill = torch.stack([hi, lo], 1).reshape(-1)
    kill = torch.where(kill >= 2 ** 31, kill - 2 ** 32, kill).to(torch.int32)
    codes = codes & ~kill
    del m, hi, lo, kill
    nn_total = int(run_l.sum())
    names = ["s%d" % i for i in range(n)]
    g = eng.PackedGenome(names, scaf_len, scaf_off, P, None, None, None, int(sl.sum()), nn_total, 0, False)
    return eng.DeviceGenome(g, dev, planes=(codes, inv, None))


def decode(dg, off, length):
    """ASCII bases [off, off+length) of the device planes ('N' where the invalid bit is set)."""
    w0, w1 = off >> 4, (off + length + 15) >> 4
    cw = dg.codes[w0:w1].cpu().numpy().view(np.uint32)
    m0, m1 = off >> 5, (off + length + 31) >> 5
    mw = dg.inv[m0:m1].cpu().numpy().view(np.uint32)
    pos = off + np.arange(length, dtype=np.int64)
    code = (cw[(pos >> 4) - w0] >> (30 - 2 * (pos & 15)).astype(np.uint32)) & 3
    bad = (mw[(pos >> 5) - m0] >> (31 - (pos & 31)).astype(np.uint32)) & 1
    out = LETTERS[code]
    out[bad.astype(bool)] = ord("N")
    return out




def c4_spec(scale: float = 1.0):
    """(scaffold lengths, N runs) of the C4 stand-in: 24 scaffolds of 50-250 Mbp summing to 3 Gbp x scale, plus two small
    ones; one 3 Mbp N run per long scaffold."""
    rng = np.random.Generator(np.random.PCG64(4004))
    lens = rng.uniform(50e6, 250e6, 24)
    lens = (lens * (3.0e9 * scale / lens.sum())).astype(np.int64)
    lens = np.concatenate([lens, [3_000_017, 1_234_567]])
    nrun = max(int(3_000_000 * scale), 5000)
    runs = (list(range(24)) + [24, 25], [int(lens[s] // 3) for s in range(24)] + [1_000_000, 5], [nrun] * 24 + [517, 2500])
    return lens, runs


def c5_spec(n: int = 1_000_000, total: float = 14.0e9, seed: int = 5005):
    """(scaffold lengths, N runs) of the C5 stand-in: n scaffolds, lognormal (median ~9 kbp at full size, min 500) summing to
    ~total bases; 30 % of them carry 1-3 N runs of 10-2,000 bp (one per third of the scaffold: no overlaps)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = np.exp(rng.normal(np.log(9000.0), 1.0, n))
    lens = np.maximum((lens * (total / lens.sum())).astype(np.int64), 500)
    has = np.nonzero(rng.random(n) < 0.30)[0]
    cnt = rng.integers(1, 4, len(has))
    run_s = np.repeat(has, cnt)
    j = np.arange(len(run_s)) - np.repeat(np.cumsum(cnt) - cnt, cnt)            # 0..cnt-1 inside each scaffold
    third = lens[run_s] // 3
    run_l = np.minimum(rng.integers(10, 2001, len(run_s)), np.maximum(third // 2, 1))
    run_o = j * third + (rng.random(len(run_s)) * np.maximum(third - run_l, 1)).astype(np.int64)
    return lens, (run_s, run_o, run_l)


def c4_device_genome(eng, scale: float = 1.0, device="cuda:0"):
    lens, runs = c4_spec(scale)
    return build_device_genome(eng, lens, runs, seed=44, at_rich_block=300_000 // 16, device=device)


def c5_device_genome(eng, n: int = 1_000_000, total: float = 14.0e9, seed: int = 5005, device="cuda:0"):
    lens, runs = c5_spec(n, total, seed)
    return build_device_genome(eng, lens, runs, seed=seed + 50, device=device)
