"""torchrun entry: multi-GPU parity check (one rank per GPU, NCCL).  Rank 0 compares the gathered
rows and the all-reduced tables against the C oracle run on the whole genome."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frisk_b200 import dist as fdist, synth  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scaffolds = synth.make("C2", 0.02, seed=5) + synth.make("edge")
    params = dict(kmin=1, kmax=8, w=5000, step=2500, mask_host=False, scaffolds_all=True, rip=True)
    res, mine = fdist.score_sharded(scaffolds, fused=False, **params)          # NCCL all-reduce
    gathered = fdist.gather_rows(res, mine)
    res_f, mine_f = fdist.score_sharded(scaffolds, fused=True, **params)       # all-reduce fused into the finalise kernels
    gathered_f = fdist.gather_rows(res_f, mine_f)
    if dist.get_rank() == 0:
        print("collectives:", res.collective, "|", res_f.collective)
        assert np.array_equal(res_f.tables, res.tables), "fused peer sum != NCCL all-reduce"
        assert res_f.meta == res.meta
        assert np.array_equal(gathered_f[2], gathered[2], equal_nan=True)
        for _ in range(3):                                                     # buffer parity flips every pass
            again, _m = fdist.score_sharded(scaffolds, fused=True, **params)
            assert np.array_equal(again.tables, res.tables)
    else:
        for _ in range(3):
            fdist.score_sharded(scaffolds, fused=True, **params)
    ok = True
    if dist.get_rank() == 0:
        from oracle import c_oracle
        ref = c_oracle.run(scaffolds, threads=8, **params)
        names, coords, rows, status = gathered
        assert np.array_equal(res.tables, ref["tables"]), "all-reduced tables differ"
        assert list(res.meta) == [int(x) for x in ref["meta"]], (res.meta, ref["meta"])
        assert names == ref["names"] and np.array_equal(coords, ref["coords"])
        good = ref["status"] == 0
        err = np.abs(rows[good, 0] - ref["rows"][good, 0]) / np.maximum(np.abs(ref["rows"][good, 0]), 1e-300)
        assert err.max() < 1e-10, err.max()
        assert np.array_equal(rows[good, 1:], ref["rows"][good, 1:], equal_nan=True)
        print("sharded ok: world=%d rows=%d max KLD rel err %.2e" % (dist.get_world_size(), len(names), err.max()))
    # the same shard end to end through ONE C call per rank (frisk_b200_run_host_peers: upload overlapped with the
    # count, fused exchange, score, download)
    from frisk_b200 import engine
    dev = torch.device("cuda", local)
    shard = engine.PackedGenome.from_scaffolds([scaffolds[i] for i in mine_f], pinned=True)
    space = fdist.global_genome_space(shard.genome_space, dev)
    px = fdist.PeerExchange(8, dev)
    assert px.available, px.reason
    wins = shard.windows(params["w"], params["step"], params["scaffolds_all"])
    out = engine.HostOutputs(len(wins), 8)
    for _ in range(2):
        px.run_host(shard, wins, out, space, stream_ptr=engine._stream_ptr(dev), **params)
    one = engine.assemble(shard, shard, wins, out.tables, int(out.valid[0]), out.rows[:len(wins)], out.status[:len(wins)], 1, 8)
    assert np.array_equal(one.tables, res_f.tables), "run_host_peers: tables differ from the staged path"
    assert np.array_equal(one.rows, res_f.rows, equal_nan=True) and np.array_equal(one.status, res_f.status)
    # second scheme: one replicated genome (device-side ingest of the FASTA text on every rank), equal
    # slices of the base range and of the window list
    text = np.frombuffer(synth.fasta_bytes(scaffolds), dtype=np.uint8)
    res2, (a, b) = fdist.score_balanced(text, **params)
    gathered2 = fdist.gather_rows_in_order(res2)
    if dist.get_rank() == 0:
        names, coords, rows, status = gathered2
        assert np.array_equal(res2.tables, ref["tables"]), "balanced: all-reduced tables differ"
        assert list(res2.meta) == [int(x) for x in ref["meta"]], (res2.meta, ref["meta"])
        assert names == ref["names"] and np.array_equal(coords, ref["coords"])
        err2 = np.abs(rows[good, 0] - ref["rows"][good, 0]) / np.maximum(np.abs(ref["rows"][good, 0]), 1e-300)
        assert err2.max() < 1e-10, err2.max()
        assert np.array_equal(rows[good, 1:], ref["rows"][good, 1:], equal_nan=True)
        assert np.array_equal(rows, gathered[2], equal_nan=True), "both schemes run the same kernels on the same windows"
        print("balanced ok: rank 0 scored windows [%d, %d) of %d" % (a, b, len(names)))
    # third scheme: the text itself cut at record boundaries, every rank ingests only its byte range
    res3, (ta, tb) = fdist.score_fasta_sharded(text, **params)
    gathered3 = fdist.gather_rows_in_order(res3)
    if dist.get_rank() == 0:
        names3, coords3, rows3, status3 = gathered3
        assert np.array_equal(res3.tables, ref["tables"]), "fasta-sharded: tables differ"
        assert list(res3.meta) == [int(x) for x in ref["meta"]], (res3.meta, ref["meta"])
        assert names3 == ref["names"] and np.array_equal(coords3, ref["coords"])
        assert np.array_equal(rows3, gathered[2], equal_nan=True)
        print("fasta-sharded ok: rank 0 ingested bytes [%d, %d) of %d" % (ta, tb, len(text)))
        print("dist_check ok: world=%d rows=%d max KLD rel err %.2e" % (dist.get_world_size(), len(names), max(err.max(), err2.max())))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
