"""Multi-GPU: one process per GPU, ONE collective.  Two ways to cut the work:

* ``score_sharded``   -- scaffolds are dealt to ranks (LPT by bases); a rank only ever holds its own
  scaffolds.  Right for fragmented assemblies (C5) and when the host side is the bottleneck.
* ``score_fasta_sharded`` -- the FASTA TEXT is cut at record boundaries into ``world`` byte ranges of equal
  size; a rank uploads and tokenises only its range (device-side ingest), so host memory and PCIe traffic
  per GPU shrink with the number of GPUs (C5: 14 GB of text, 1 M scaffolds).  Rows come back in file order.
* ``score_balanced``  -- every rank ingests the whole FASTA on its own GPU (device-side ingest, one PCIe
  link per GPU; 14 Gbp packed is 5 GB of 180 GB), counts an equal slice of the BASE RANGE and scores an
  equal slice of the WINDOW LIST (by bases covered): exact balance even when 24 chromosomes meet 8 GPUs
  (C4), no halo logic because every word's owner is the rank whose slice holds its first base, rows come
  back already in reference order.

The path shards naturally: every rank counts the forward-strand k-mers of
its own scaffolds, the 87,380-counter tables (+ the genome-space scalar) are summed with a single
all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests), then every rank finalises the
tables redundantly and scores its own windows.  No other exchange exists on the data path.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np


def shard_scaffolds(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time assignment of scaffolds to ranks, balanced by bases; scaffold order
    inside a rank is preserved (row order within a scaffold matters downstream: the HMM sees each
    scaffold's rows as one sequence, F:769)."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(lengths[i])
    return [sorted(x) for x in out]


def global_genome_space(local_space: int, device=None, group=None) -> int:
    """Sum of totalLen - nnTotal over ranks (a property of the input, known at ingest time)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(local_space)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def make_allreduce(group=None):
    """Returns the hook for engine.Pipeline(allreduce=...): sums the forward counts over ranks in
    place -- the path's one collective.  The genome space passed in must already be the global one
    (global_genome_space) so the step needs no host round trip."""
    import torch.distributed as dist

    def hook(d_fwd, genome_space: int) -> int:
        dist.all_reduce(d_fwd, op=dist.ReduceOp.SUM, group=group)
        return genome_space

    return hook


class PeerExchange:
    """The all-reduce of the counters FUSED into the finalise kernels (frisk_b200_finalize_tables_peers):
    every rank's counter buffer lives in torch symmetric memory (CUDA IPC peer mappings over NVLink /
    NVSwitch -- torch is the plumbing: allocation, handle exchange, the cross-GPU stream barrier), and
    the kernels that marginalise the counters read and sum all ranks' buffers directly.  Two buffers
    alternate so that a rank may start counting the next genome while a slower peer still reads the
    previous one.  Falls back (``available`` False) when symmetric memory cannot be set up; callers
    then use ``make_allreduce``."""

    def __init__(self, kmax: int, device, group=None):
        import torch
        import torch.distributed as dist
        from . import _lib
        self.available = False
        self.reason = ""
        self.parity = 0
        self.world = dist.get_world_size(group)
        self.tsz = _lib.table_size(1, kmax)
        self.stride = (self.tsz + 15) // 16 * 16          # int64 elements per buffer (128-byte aligned)
        if kmax > 8 or self.world > 16:
            self.reason = "kmax > 8 or world > 16"
            return
        try:
            import torch.distributed._symmetric_memory as symm
            pg = group if group is not None else dist.group.WORLD
            self.rank = dist.get_rank(group)
            self.epoch = 0
            # [counters, even epochs | counters, odd epochs | one arrival flag per rank]
            self.buf = symm.empty(2 * self.stride + 16, dtype=torch.int64, device=device)
            self.hdl = symm.rendezvous(self.buf, pg)
            ptrs = [int(x) for x in self.hdl.buffer_ptrs]
            assert len(ptrs) == self.world and all(ptrs)
            self.ptr_arrays = []
            for par in (0, 1):
                arr = (C.c_void_p * self.world)(*[C.c_void_p(p + par * self.stride * 8) for p in ptrs])
                self.ptr_arrays.append(arr)
            self.flag_array = (C.c_void_p * self.world)(*[C.c_void_p(p + 2 * self.stride * 8) for p in ptrs])
            self.buf.zero_()
            torch.cuda.synchronize(device)
            dist.barrier(group)                         # every rank's flags are zero before anyone posts
            self.available = True
        except Exception as e:      # no P2P / no symmetric-memory backend in this build
            self.reason = "%s: %s" % (type(e).__name__, e)

    def local(self):
        """This rank's counter buffer for the current step (a view of the symmetric allocation)."""
        a = self.parity * self.stride
        return self.buf[a:a + self.tsz]

    def run_host(self, genome, wins, out, genome_space: int, kmin=1, kmax=8, mask_host=False, rip=True, stream_ptr=None, **_):
        """This rank's share end to end in ONE C call (frisk_b200_run_host_peers): chunked upload of the
        pinned host planes overlapped with the count, fused exchange, score, download into ``out``."""
        from . import _lib, engine
        self.epoch += 1
        hi, hv = genome.inv_sparse()
        n = len(wins)
        rc = _lib.lib().frisk_b200_run_host_peers(
            engine._ptr(genome.codes), engine._ptr(hi), engine._ptr(hv), len(hi), engine._ptr(genome.low), genome.padded_len,
            engine._ptr(wins.off), engine._ptr(wins.length), n, wins.max_len, kmin, kmax, int(mask_host), int(rip),
            int(genome_space), C.c_void_p(int(self.local().data_ptr())), self.ptr_arrays[self.parity], self.flag_array,
            self.rank, self.world, self.epoch, engine._ptr(out.rows), engine._ptr(out.status), engine._ptr(out.tables),
            engine._ptr(out.valid), stream_ptr)
        _lib.check(rc, "frisk_b200_run_host_peers")
        self.parity ^= 1
        return out

    def finalize_ivom(self, kmin: int, kmax: int, genome_space: int, d_tables, d_valid, d_ig, stream_ptr):
        """The fused barrier + sum + finalise + genome IVOM table, one cooperative launch.  Flips the buffer."""
        from . import _lib
        self.epoch += 1
        rc = _lib.lib().frisk_b200_finalize_ivom(None, self.ptr_arrays[self.parity], self.flag_array, self.rank, self.world,
                                                 self.epoch, kmin, kmax, int(genome_space), C.c_void_p(int(d_tables.data_ptr())),
                                                 C.c_void_p(int(d_valid.data_ptr())), C.c_void_p(int(d_ig.data_ptr())), stream_ptr)
        _lib.check(rc, "frisk_b200_finalize_ivom")
        self.parity ^= 1

    def finalize(self, kmax: int, d_tables, d_valid, stream_ptr):
        """The fused barrier + sum + finalise (the GPUs synchronise inside the first kernel).  Flips the buffer."""
        from . import _lib
        self.epoch += 1
        rc = _lib.lib().frisk_b200_finalize_tables_peers(self.ptr_arrays[self.parity], self.flag_array, self.rank, self.world,
                                                         self.epoch, kmax, 1, C.c_void_p(int(d_tables.data_ptr())),
                                                         C.c_void_p(int(d_valid.data_ptr())), stream_ptr)
        _lib.check(rc, "frisk_b200_finalize_tables_peers")
        self.parity ^= 1


def reduce_counts_cpu(tables: np.ndarray, genome_space: int, group=None) -> Tuple[np.ndarray, int]:
    """Same reduction on host arrays (gloo); used by the CPU world_size-2 tests of the sharding logic."""
    import torch
    import torch.distributed as dist
    buf = torch.from_numpy(np.concatenate([tables.astype(np.int64), np.array([genome_space], np.int64)]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    out = buf.numpy()
    return out[:-1].astype(np.uint64), int(out[-1])


def score_sharded(scaffolds, group=None, device=None, fused: bool = True, **params):
    """Multi-GPU hot path, one call per rank: this rank packs and scores its own scaffolds against
    the background of ALL ranks' scaffolds (one NCCL all-reduce inside the pipeline).

    Returns (local HotPathResult whose ``tables`` are the global ones and ``meta`` the global
    (totalLen, exMax, nnTotal), list of this rank's scaffold indices)."""
    import torch
    import torch.distributed as dist
    from . import engine
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    mine = shard_scaffolds([len(s) for _, s in scaffolds], world)[rank]
    genome = engine.PackedGenome.from_scaffolds([scaffolds[i] for i in mine], pinned=True)
    space = global_genome_space(genome.genome_space, device, group)
    peers = PeerExchange(params.get("kmax", 8), device, group) if fused else None
    pipe = engine.Pipeline(genome, device=device, allreduce=make_allreduce(group), genome_space=space, peers=peers, **params)
    pipe.enqueue()
    res = pipe.result()
    res.collective = "fused peer sum (NVLink)" if pipe.peers is not None else "NCCL all-reduce"
    # totalLen and nnTotal are sums over scaffolds; exMax = (all kmax-word start positions) - (valid
    # kmax-words), where the valid count finalised from the all-reduced counters is already global
    kmax = pipe.kmax
    possible = int(np.maximum(genome.scaf_len.astype(np.int64) - kmax + 1, 0).sum())
    valid_global = possible - res.meta[1]
    meta = torch.tensor([res.meta[0], possible, res.meta[2]], dtype=torch.int64, device=device)
    dist.all_reduce(meta, op=dist.ReduceOp.SUM, group=group)
    tot = [int(x) for x in meta.tolist()]
    res.meta = (tot[0], tot[1] - valid_global, tot[2])
    return res, mine


def gather_rows(res, mine, group=None, dst: int = 0):
    """Collect every rank's rows on rank `dst` in the reference's order (scaffold order of the
    input, windows in order inside a scaffold).  Returns (names, coords, rows, status) or None."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    original = np.asarray(mine, dtype=np.int64)[res.row_scaf] if len(res.row_scaf) else np.zeros(0, np.int64)
    payload = (original, res.names, res.coords, res.rows, res.status)
    out = [None] * world if rank == dst else None
    dist.gather_object(payload, out, dst=dst, group=group)
    if rank != dst:
        return None
    key = np.concatenate([o[0] for o in out])
    names = [n for o in out for n in o[1]]
    coords = np.concatenate([o[2].reshape(-1, 2) for o in out])
    rows = np.concatenate([o[3].reshape(-1, 5) for o in out])
    status = np.concatenate([o[4] for o in out])
    order = np.argsort(key, kind="stable")
    return [names[i] for i in order], coords[order], rows[order], status[order]


# ---------------------------------------------------------------------- balanced slices of one replicated genome
def split_base_range(padded_len: int, world: int) -> List[Tuple[int, int]]:
    """[first_base, last_base) per rank: contiguous, multiples of 32, together exactly the range the
    single-GPU background pass covers ([0, padded_len - 32): the last word is look-ahead padding)."""
    words = padded_len // 32 - 1
    cuts = [(words * r) // world for r in range(world + 1)]
    return [(32 * cuts[r], 32 * cuts[r + 1]) for r in range(world)]


def split_windows(lengths: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """[a, b) window-index ranges per rank, contiguous, balanced by the bases the windows cover
    (--scaffoldsAll windows vary in length)."""
    n = len(lengths)
    if n == 0:
        return [(0, 0)] * world
    csum = np.cumsum(np.asarray(lengths, dtype=np.int64))
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(csum, (total * r) // world, side="left")))
    cuts.append(n)
    cuts = [min(max(c, cuts[i - 1] if i else 0), n) for i, c in enumerate(cuts)]
    return [(cuts[r], max(cuts[r], cuts[r + 1])) for r in range(world)]


def score_balanced(fasta_text, group=None, device=None, host_text=None, fused: bool = True, **params):
    """Multi-GPU hot path on one replicated genome (see the module docstring).  Returns this rank's
    HotPathResult (global tables and meta, this rank's slice of the rows) and its window range."""
    import torch
    import torch.distributed as dist
    from . import engine
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    dq = engine.DeviceGenome.from_fasta_bytes(fasta_text, device)
    dh = engine.DeviceGenome.from_fasta_bytes(host_text, device) if host_text is not None else dq
    wins_all = dq.host.windows(params.get("w", 5000), params.get("step", 2500), params.get("scaffolds_all", False))
    a, b = split_windows(wins_all.length, world)[rank]
    peers = PeerExchange(params.get("kmax", 8), device, group) if fused else None
    pipe = engine.Pipeline(dq, dh if dh is not dq else None, device=device, allreduce=make_allreduce(group),
                           genome_space=dh.host.genome_space, wins=wins_all.slice(a, b),
                           bg_range=split_base_range(dh.host.padded_len, world)[rank], peers=peers, **params)
    pipe.enqueue()
    res = pipe.result()
    res.collective = "fused peer sum (NVLink)" if pipe.peers is not None else "NCCL all-reduce"
    return res, (a, b)


def gather_rows_in_order(res, group=None, dst: int = 0):
    """Rows of ``score_balanced``: the ranks hold consecutive slices of the window list, so rank order
    is reference order.  Returns (names, coords, rows, status) on ``dst``, None elsewhere."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    out = [None] * world if rank == dst else None
    dist.gather_object((res.names, res.coords, res.rows, res.status), out, dst=dst, group=group)
    if rank != dst:
        return None
    return ([n for o in out for n in o[0]], np.concatenate([o[1].reshape(-1, 2) for o in out]),
            np.concatenate([o[2].reshape(-1, 5) for o in out]), np.concatenate([o[3] for o in out]))


# ---------------------------------------------------------------------- the FASTA text cut at record boundaries
def split_fasta_text(text: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Byte ranges [a, b) of FASTA text per rank: consecutive, covering the whole text, every range
    starting at a header line (a '>' first on its line; the first range also keeps whatever precedes the
    first header, which the reference ignores), sizes as equal as record boundaries allow."""
    buf = np.ascontiguousarray(text, dtype=np.uint8)
    n = int(buf.shape[0])

    def next_header(pos: int) -> int:
        """First offset >= pos at which a header line starts (n if none)."""
        if pos <= 0:
            return 0
        if pos >= n:
            return n
        j = _find(buf, b"\n>", pos - 1, n)            # a header needs the newline before it
        return n if j < 0 else j + 1

    cuts = [0]
    for r in range(1, world):
        cuts.append(max(next_header((n * r) // world), cuts[-1]))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def _find(buf: np.ndarray, pat: bytes, lo: int, hi: int) -> int:
    """bytes.find over a numpy buffer without copying more than a window at a time."""
    step = 1 << 24
    while lo < hi:
        end = min(lo + step + len(pat), hi)
        k = buf[lo:end].tobytes().find(pat)
        if k >= 0:
            return lo + k
        lo += step
    return -1


def score_fasta_sharded(fasta_text, group=None, device=None, fused: bool = True, local_text=None, row_names: bool = True,
                        peers=None, **params):
    """Multi-GPU hot path from FASTA text with the INGEST sharded too: rank r uploads and tokenises only
    its byte range (split_fasta_text), counts and scores its own records against the background of all
    ranks (one exchange of the counters).  Returns this rank's HotPathResult with global tables/meta;
    ``gather_rows_in_order`` puts the rows back in file order.  ``local_text``: this rank's byte range when the
    caller has already cut the text (a rank then never sees the rest of the file); works without a process group
    (one GPU) too."""
    import torch
    import torch.distributed as dist
    from . import engine
    multi = dist.is_available() and dist.is_initialized()
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if multi else (0, 1)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    if local_text is not None:
        mine = np.ascontiguousarray(local_text, dtype=np.uint8)
        a, b = 0, int(mine.shape[0])
    else:
        buf = np.ascontiguousarray(fasta_text, dtype=np.uint8) if not isinstance(fasta_text, (bytes, bytearray)) else np.frombuffer(fasta_text, np.uint8)
        a, b = split_fasta_text(buf, world)[rank]
        mine = buf[a:b]
    dq = engine.DeviceGenome.from_fasta_bytes(mine, device)
    g = dq.host
    space = global_genome_space(g.genome_space, device, group) if world > 1 else g.genome_space
    if peers is None and fused and world > 1:          # (a caller that runs many genomes passes its PeerExchange in)
        peers = PeerExchange(params.get("kmax", 8), device, group)
    pipe = engine.Pipeline(dq, device=device, allreduce=make_allreduce(group) if world > 1 else None, genome_space=space, peers=peers, **params)
    pipe.enqueue()
    res = pipe.result(names=row_names)
    res.collective = "fused peer sum (NVLink)" if pipe.peers is not None else "NCCL all-reduce"
    kmax = pipe.kmax
    possible = int(np.maximum(g.scaf_len.astype(np.int64) - kmax + 1, 0).sum())
    valid_global = possible - res.meta[1]              # finalised from the summed counters: already global
    meta = torch.tensor([res.meta[0], possible, res.meta[2]], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(meta, op=dist.ReduceOp.SUM, group=group)
    tot = [int(x) for x in meta.tolist()]
    res.meta = (tot[0], tot[1] - valid_global, tot[2])
    return res, (a, b)
