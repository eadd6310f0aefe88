"""--scaffoldsAll on a fragmented assembly (C5 at 1 %: 140 Mbp, 10 k scaffolds): windows of mixed length."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frisk_b200 import engine, synth, _lib
g = engine.PackedGenome.from_scaffolds(synth.make("C5", 0.01))
for sa, opt in ((False, None), (True, None), (False, b"force_direct_kernel"), (True, b"force_direct_kernel")):
    if opt:
        _lib.check(_lib.lib().frisk_b200_set_option(opt, 1), "option")
    pipe = engine.Pipeline(g, scaffolds_all=sa)
    for _ in range(3):
        pipe.enqueue()
    torch.cuda.synchronize()
    ms = []
    for _ in range(7):
        marks = []
        pipe.enqueue(marks)
        torch.cuda.synchronize()
        ms.append(marks[2].elapsed_time(marks[3]))
    ms.sort()
    if opt:
        _lib.lib().frisk_b200_set_option(opt, 0)
    print("%s scaffolds_all=%s: %d windows (max %d bases, %d longer than 5104), score %.3f ms, %.1f M windows/s" % (
        (opt or b"default").decode(), sa, len(pipe.wins), pipe.wins.max_len, int((pipe.wins.length > 5104).sum()), ms[3], len(pipe.wins) / ms[3] / 1e3))
