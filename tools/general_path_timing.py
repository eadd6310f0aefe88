"""Timing of the general path (kmax 9..12, long windows) on C2, planes resident."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frisk_b200 import engine, synth
g = engine.PackedGenome.from_scaffolds(synth.make("C2", 1.0))
for kw in (dict(kmax=8), dict(kmax=9), dict(kmax=10), dict(kmax=12), dict(kmax=8, w=100000, step=50000)):
    pipe = engine.Pipeline(g, **kw)
    pipe.enqueue(); torch.cuda.synchronize()
    marks = []
    pipe.enqueue(marks); torch.cuda.synchronize()
    st = [marks[i].elapsed_time(marks[i + 1]) for i in range(3)]
    print(kw, "windows", len(pipe.wins), "background %.3f ms, tables+ivom %.3f ms, score %.3f ms" % tuple(st))
    del pipe
    torch.cuda.empty_cache()
