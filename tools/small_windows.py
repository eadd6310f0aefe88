"""Throughput of short windows (-w 1000 -i 500) on C2, planes resident: the 2-round instantiation of the bucket kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frisk_b200 import engine, synth
g = engine.PackedGenome.from_scaffolds(synth.make("C2", 1.0))
for w, step in ((1000, 500), (2000, 1000), (5000, 2500)):
    pipe = engine.Pipeline(g, w=w, step=step)
    for _ in range(3):
        pipe.enqueue()
    torch.cuda.synchronize()
    ms = []
    for _ in range(10):
        marks = []
        pipe.enqueue(marks)
        torch.cuda.synchronize()
        ms.append(marks[2].elapsed_time(marks[3]))
    ms.sort()
    print("w=%d step=%d: %d windows, score kernel %.3f ms (median), %.1f M windows/s" % (w, step, len(pipe.wins), ms[5], len(pipe.wins) / ms[5] / 1e3))
