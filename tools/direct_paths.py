"""The direct window kernel (frisk_direct.cu) on a small genome through all of its paths: K = 7 and 8, dump on/off,
kmin > 1, short / long windows (2-, 5-, 8-round instantiations), the hand-over to the bucketed kernel (a K-mer seen
256+ times, many N boundaries).  What one puts under compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from frisk_b200 import _lib, engine, synth
rng = np.random.Generator(np.random.PCG64(8))
a = synth.iid_bases(rng, 30_000, 0.45)
a[2_000:2_300] = ord("A")
for p in range(10_000, 13_000, 40):
    a[p] = ord("N")
sc = synth.make("edge") + [("handover", a)]
g = engine.PackedGenome.from_scaffolds(sc)
L = _lib.lib()
L.frisk_b200_set_option(b"force_direct_kernel", 1)
for kw in (dict(), dict(kmax=7), dict(kmax=7, kmin=3, w=1500, step=700, scaffolds_all=True), dict(w=7000, step=3000, kmin=6),
           dict(dump=True), dict(kmax=7, dump=True)):
    res = engine.run(g, **kw)
    print(kw, len(res.rows), "rows", float(np.nansum(res.rows[:, 0])))
torch.cuda.synchronize()
print("done")
