"""Score-stage time on C2 for windows longer than the 8,186 bases the nibble / bucketed kernels hold (dense-table kernel up
to 65,535 bases, general path beyond), kmax 8 and 7."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frisk_b200 import engine, synth
g = engine.PackedGenome.from_scaffolds(synth.make("C2", 1.0))
for kw in (dict(w=5000, step=2500), dict(w=8000, step=4000), dict(w=10000, step=5000), dict(w=14000, step=7000), dict(w=20000, step=10000), dict(w=50000, step=25000),
           dict(w=10000, step=5000, kmax=7)):
    pipe = engine.Pipeline(g, **kw)
    pipe.enqueue(); torch.cuda.synchronize()
    marks = []
    pipe.enqueue(marks); torch.cuda.synchronize()
    st = [marks[i].elapsed_time(marks[i + 1]) for i in range(3)]
    print(kw, "windows", len(pipe.wins), "score %.3f ms" % st[2], flush=True)
    del pipe
    torch.cuda.empty_cache()
