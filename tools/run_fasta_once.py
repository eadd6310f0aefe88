"""One warm engine.run_fasta call on the C2 text (for an ncu launch list of the streamed FASTA path)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from frisk_b200 import engine, synth, _lib

raw = np.frombuffer(synth.fasta_bytes(synth.make("C2", 1.0)), dtype=np.uint8)
text = engine._alloc(raw.shape[0], np.uint8, True)
text[:] = raw
out = None
for opt in sys.argv[2:]:
    k, v = opt.split("=")
    _lib.check(_lib.lib().frisk_b200_set_option(k.encode(), int(v)), "set_option")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    out = engine.run_fasta(text, out=out, assemble_result=False, **bench.PARAMS)
torch.cuda.synchronize()
print("windows", out.n_win)
