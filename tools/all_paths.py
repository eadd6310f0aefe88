"""Every kernel path once on a small genome (a quick whole-library exercise; also what one would put under compute-sanitizer where that is allowed)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from frisk_b200 import _lib, engine, synth
sc = synth.make("edge") + synth.make("C2", 0.002, seed=3)
g = engine.PackedGenome.from_scaffolds(sc)
text = np.frombuffer(synth.fasta_bytes(sc), dtype=np.uint8)
L = _lib.lib()
runs = [("bucket K=8", dict()), ("bucket K=7", dict(kmax=7)), ("small K=5", dict(kmax=5, kmin=2)), ("small K=1", dict(kmax=1)),
        ("bucket short windows", dict(w=1000, step=500, scaffolds_all=True)), ("bucket 8 rounds", dict(w=8000, step=4000)),
        ("extension K=9", dict(kmax=9)), ("extension K=12 kmin 9", dict(kmax=12, kmin=9, w=3000, step=1000, scaffolds_all=True)),
        ("nibble K=7", dict(kmax=7)), ("nibble 32 positions/thread", dict(w=8000, step=4000, scaffolds_all=True)),
        ("long windows (dense)", dict(w=20000, step=10000))]
for name, kw in runs:
    res = engine.run(g, **kw)
    print(name, len(res.rows), "rows")
L.frisk_b200_set_option(b"force_dense_kernel", 1)
print("dense forced", len(engine.run(g, kmax=6).rows)); L.frisk_b200_set_option(b"force_dense_kernel", 0)
L.frisk_b200_set_option(b"force_general_kernel", 1)
print("general forced", len(engine.run(g).rows)); L.frisk_b200_set_option(b"force_general_kernel", 0)
print("run_host", len(engine.run_host(g, scaffolds_all=True).rows))
print("run_fasta", len(engine.run_fasta(text, scaffolds_all=True).rows))
dg = engine.DeviceGenome(g)
print("features", engine.region_features(dg, g.scaf_off, g.scaf_len.astype(np.uint32), 1, 6).shape)
print("dump", engine.run(g, dump=True).win_tables.shape)
print("sweep", sorted(engine.run_sweep(g)))
for opt in (b"force_bucket_kernel", b"force_direct_kernel"):
    L.frisk_b200_set_option(opt, 1); print(opt.decode(), len(engine.run(g).rows)); L.frisk_b200_set_option(opt, 0)
torch.cuda.synchronize()
print("done")
