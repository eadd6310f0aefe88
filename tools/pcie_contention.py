"""Host-to-device copy rate of every GPU alone and of all GPUs at once (pinned buffers, 10 MB -- the planes of a C2 shard --
and 256 MB), one process per GPU under torchrun.  Explains bench.py's e2e.stages.upload_done at N > 1: is the slowdown the
path's or the host's?"""
import os, json, time
import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def rate(nbytes, reps):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    out = {}
    # alone: ranks take turns
    alone = 0.0
    for turn in range(world):
        barrier()
        if turn == rank:
            ts = []
            for _ in range(reps):
                e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize(dev)
                ts.append(e0.elapsed_time(e1))
            alone = nbytes / (sorted(ts)[len(ts) // 2] * 1e-3) / 1e9
        barrier()
    # together: every rank starts after the same barrier
    ts = []
    for _ in range(reps):
        barrier()
        e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize(dev)
        ts.append(e0.elapsed_time(e1))
    together = nbytes / (sorted(ts)[len(ts) // 2] * 1e-3) / 1e9
    return alone, together


res = {}
for nbytes, reps in ((10 << 20, 30), (256 << 20, 8)):
    a, t = rate(nbytes, reps)
    v = torch.tensor([a, t], dtype=torch.float64, device=dev)
    if world > 1:
        allv = [torch.zeros_like(v) for _ in range(world)]
        dist.all_gather(allv, v)
    else:
        allv = [v]
    if rank == 0:
        al = [float(x[0]) for x in allv]; tg = [float(x[1]) for x in allv]
        res["%d MiB" % (nbytes >> 20)] = {"alone_GBps_per_gpu": [round(x, 1) for x in al], "together_GBps_per_gpu": [round(x, 1) for x in tg],
                                          "together_aggregate_GBps": round(sum(tg), 1)}
if rank == 0:
    print(json.dumps({"gpus": world, "h2d": res}))
if world > 1:
    dist.destroy_process_group()
