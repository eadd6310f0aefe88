"""Where the time of the C5 ingest-inclusive path goes (one GPU, 125,000 scaffolds / 1.75 Gbp of FASTA text):
wall-clock per stage of dist.score_fasta_sharded's single-rank path."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from frisk_b200 import engine, _lib

dev = torch.device("cuda:0")
text, bases, _ = bench.c5_text_on_device(0, dev, 125_000, 14.0e9 / 8)
params = dict(bench.PARAMS, scaffolds_all=True)
out = {}
for rep in range(3):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    dq = engine.DeviceGenome.from_fasta_bytes(text, dev); torch.cuda.synchronize(); t.append(time.perf_counter())
    wins = dq.host.windows(params["w"], params["step"], True); t.append(time.perf_counter())
    pipe = engine.Pipeline(dq, device=dev, wins=wins, **params); torch.cuda.synchronize(); t.append(time.perf_counter())
    pipe.enqueue(); torch.cuda.synchronize(); t.append(time.perf_counter())
    res = pipe.result(names=False); t.append(time.perf_counter())
    out = dict(zip(["ingest(H2D+tokenise+pack+names)", "windows()", "Pipeline init (allocs, window upload)", "kernels", "result (D2H + assemble)"],
                   [round((b - a) * 1e3, 2) for a, b in zip(t, t[1:])]))
    out["total_ms"] = round((t[-1] - t[0]) * 1e3, 2)
    del pipe, dq
print(json.dumps(out))
