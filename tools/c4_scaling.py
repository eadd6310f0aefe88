"""C4 (3 Gbp, 24 chromosome-sized scaffolds, 1.2 M windows) STRONG scaling: the same genome on N GPUs, every rank
holding the planes (drawn on its own device from the same seed), counting an equal slice of the base range and
scoring an equal slice of the window list (frisk_b200.dist.score_balanced's scheme); one NCCL all-reduce of the
87,380 counters per step.  Launch: python -m torch.distributed.run --nproc-per-node N tools/c4_scaling.py
Prints one JSON line on rank 0: step time (CUDA events, max over ranks), Gbp/s, and two size-independent checks that
must not depend on N (sum of all KLD scores, sum of the genome tables)."""
import json, os, sys
local = int(os.environ.get("LOCAL_RANK", "0"))
os.environ["CUDA_VISIBLE_DEVICES"] = str(local)          # before torch: each rank sees its GPU as cuda:0
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from frisk_b200 import engine as eng, dist as fdist
from frisk_b200.synth_device import c4_spec, build_device_genome

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
steps = int(os.environ.get("C4_STEPS", "5"))
dev = torch.device("cuda:0")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
lens, runs = c4_spec()
dg = build_device_genome(eng, lens, runs, seed=44, at_rich_block=300_000 // 16)
g = dg.host
wins_all = g.windows(5000, 2500, False)
a, b = fdist.split_windows(wins_all.length, world)[rank]
kw = {}
if world > 1:
    kw = dict(allreduce=fdist.make_allreduce(), bg_range=fdist.split_base_range(g.padded_len, world)[rank])
pipe = eng.Pipeline(dg, wins=wins_all.slice(a, b), **kw)

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

for _ in range(2):
    pipe.enqueue()
barrier()
ms = []
for _ in range(steps):
    barrier()
    marks = []
    pipe.enqueue(marks)
    torch.cuda.synchronize()
    ms.append([marks[i].elapsed_time(marks[i + 1]) for i in range(3)])
ms = np.array(ms)
step = torch.tensor([float(ms.sum(1).mean())] + [float(x) for x in ms.mean(0)], dtype=torch.float64, device=dev)
st = pipe.d_status.view(torch.int32)
ok = st == 0
kld = pipe.d_rows[:, 0]
chk = torch.stack([torch.where(ok, kld, torch.zeros_like(kld)).sum(), ok.sum().to(torch.float64),
                   ((st & 8) != 0).sum().to(torch.float64)])
if world > 1:
    dist.all_reduce(step, op=dist.ReduceOp.MAX)
    dist.all_reduce(chk, op=dist.ReduceOp.SUM)
if rank == 0:
    t = float(step[0])
    print(json.dumps({"workload": "C4: 3 Gbp, 26 scaffolds, w=5000 step=2500 k=1..8", "n_gpus": world, "bases": int(g.total_len),
                      "windows": len(wins_all), "ms_per_step": t, "gbp_per_s": g.total_len / t / 1e6,
                      "stage_ms_max": {"background": float(step[1]), "tables+ivom+allreduce": float(step[2]), "score": float(step[3])},
                      "scaling": "strong", "kld_sum": repr(float(chk[0])), "rows_ok": int(chk[1]), "rows_excluded": int(chk[2]),
                      "tables_sum": int(pipe.d_tables.sum().item()), "steps": steps}))
if world > 1:
    dist.destroy_process_group()
