"""CPU (gloo, world_size 2): the multi-GPU path's host logic -- scaffold sharding, the single
all-reduce of the counter tables, and re-assembly of rows in reference order.  The per-shard
arithmetic is supplied by the C oracle here (there is no GPU); on the GPU box the same plumbing is
exercised with the CUDA kernels by tests/test_dist_gpu.py."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from frisk_b200 import dist as fdist
from frisk_b200 import synth


def test_shard_scaffolds_balanced_and_complete():
    lens = [int(x) for x in np.random.default_rng(0).integers(500, 250_000, 57)]
    for world in (1, 2, 3, 8):
        parts = fdist.shard_scaffolds(lens, world)
        assert sorted(i for p in parts for i in p) == list(range(len(lens)))
        assert all(p == sorted(p) for p in parts)
        loads = [sum(lens[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(lens)


def test_balanced_slices_cover_everything_once():
    rng = np.random.default_rng(4)
    for world in (1, 2, 3, 8):
        for padded in (128, 4096, 128 * 77777, 128 * 1_000_003):
            parts = fdist.split_base_range(padded, world)
            assert parts[0][0] == 0 and parts[-1][1] == padded - 32
            assert all(a % 32 == 0 and b % 32 == 0 and a <= b for a, b in parts)
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 32
        for lengths in (np.full(1000, 5000), rng.integers(1, 6251, 100_000), np.array([5000]), np.zeros(0, np.int64)):
            parts = fdist.split_windows(lengths, world)
            assert parts[0][0] == 0 and parts[-1][1] == len(lengths)
            assert all(a <= b for a, b in parts) and all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            if len(lengths) >= 100 * world:
                loads = [int(lengths[a:b].sum()) for a, b in parts]
                assert max(loads) - min(loads) <= 2 * int(lengths.max())


def test_fasta_text_is_cut_at_record_boundaries():
    sc = synth.make("C5", 0.00005, seed=3) + synth.make("edge")
    text = np.frombuffer(synth.fasta_bytes(sc) + b">tail_without_newline\nACGT", dtype=np.uint8)
    whole = text.tobytes()
    for world in (1, 2, 3, 8, 64):
        parts = fdist.split_fasta_text(text, world)
        assert parts[0][0] == 0 and parts[-1][1] == len(text)
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        for a, b in parts:
            assert a == b or a == 0 or (whole[a:a + 1] == b">" and whole[a - 1:a] == b"\n")
        # every record lands in exactly one part, in order
        names = [n for a, b in parts for n in __import__("frisk_b200").engine.PackedGenome.from_fasta_bytes(whole[a:b]).names]
        assert names == [n for n, _ in sc] + ["tail_without_newline"]
        if world <= 8:
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) < 2 * max(len(s) for _, s in sc) + 4096
    # a '>' inside a sequence line is not a record boundary
    odd = np.frombuffer(b">a\nAC>GT\nACGT\n>b\nAC\n", dtype=np.uint8)
    assert fdist.split_fasta_text(odd, 2) == [(0, 14), (14, 20)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, queue):
    import torch.distributed as dist
    from oracle import c_oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scaffolds = synth.make("C2", 0.004, seed=5) + synth.make("edge")
    mine = fdist.shard_scaffolds([len(s) for _, s in scaffolds], world)[rank]
    shard = [scaffolds[i] for i in mine]
    seq, off = c_oracle.concat(shard)
    tabs, meta = c_oracle.background(seq, off, 1, 8, False, threads=1)     # this rank's counters
    space = int(meta[0]) - int(meta[2])
    g_tabs, g_space = fdist.reduce_counts_cpu(tabs, space)                  # THE collective
    sidx, woff, wlen, st, sp = c_oracle.crawl(seq, off)
    g_meta = np.array([0, 0, 0], np.uint64)
    g_meta[0] = g_space
    rows, status = c_oracle.score(seq, woff, wlen, g_tabs, g_meta, threads=1)

    class R:   # the fields gather_rows uses
        pass
    res = R()
    res.row_scaf = sidx
    res.names = [shard[i][0] for i in sidx]
    res.coords = np.stack([st, sp], 1)
    res.rows = rows
    res.status = status
    gathered = fdist.gather_rows(res, mine)
    if rank == 0:
        queue.put((g_tabs, g_space, gathered))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_rank_gloo_matches_single_process(world):
    from oracle import c_oracle
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, queue)) for r in range(world)]
    for p in procs:
        p.start()
    g_tabs, g_space, (names, coords, rows, status) = queue.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    scaffolds = synth.make("C2", 0.004, seed=5) + synth.make("edge")
    ref = c_oracle.run(scaffolds, threads=2)
    assert np.array_equal(g_tabs, ref["tables"])
    assert g_space == int(ref["meta"][0]) - int(ref["meta"][2])
    assert names == ref["names"]
    assert np.array_equal(coords, ref["coords"])
    assert np.array_equal(status, ref["status"])
    assert np.array_equal(rows, ref["rows"], equal_nan=True)     # same tables, same windows -> same bits
