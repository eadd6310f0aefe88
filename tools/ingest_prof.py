"""FASTA-text-in, rows-out on C2 (for an ncu launch list of the ingest + scoring kernels) and a
host-side wall-clock breakdown of the same call."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from frisk_b200 import engine, synth

sc = synth.make("C2", 1.0)
raw = np.frombuffer(synth.fasta_bytes(sc), dtype=np.uint8)
text = engine._alloc(raw.shape[0], np.uint8, True)
text[:] = raw
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
out = None
for i in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dq = engine.DeviceGenome.from_fasta_bytes(text)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    wins = dq.host.windows(5000, 2500, False)
    t2 = time.perf_counter()
    if out is None:
        out = engine.HostOutputs(len(wins), 8)
    engine.run_resident(dq, wins=wins, out=out, assemble_result=False)
    t3 = time.perf_counter()
    print("rep %d: ingest %.3f ms, windows %.3f ms, run_resident %.3f ms" % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
