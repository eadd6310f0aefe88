"""Stage marks (frisk_b200_last_run_timing) of the one-call FASTA path on the C2 text: ms since the start of the call at which
the text was uploaded, tokenised + packed + counted, tables finalised, windows scored, rows on the host."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
import bench
from frisk_b200 import engine, _lib, synth

dev = torch.device("cuda:0")
raw = np.frombuffer(synth.fasta_bytes(synth.make("C2", 1.0)), dtype=np.uint8)
text = engine._alloc(raw.shape[0], np.uint8, True)
text[:] = raw
L = _lib.lib()
out = None
rows = []
for opt in [None] + [a for a in sys.argv[1:]]:
    if opt:
        k, v = opt.split("=")
        _lib.check(L.frisk_b200_set_option(k.encode(), int(v)), "set_option")
    reps = []
    for rep in range(10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = engine.run_fasta(text, out=out, assemble_result=False, **bench.PARAMS)
        wall = (time.perf_counter() - t0) * 1e3
        ms = (C.c_float * 8)(); nn = C.c_int(0)
        L.frisk_b200_last_run_timing(ms, 8, C.byref(nn))
        reps.append([round(float(x), 3) for x in ms[:5]] + [round(wall, 3), round(float(ms[5]), 3)])
    med = [float(np.median([r[i] for r in reps[3:]])) for i in range(7)]
    rows.append({"option": opt, "uploaded/counted/finalised/scored/end/wall/score_start_ms": med})
    if opt:
        L.frisk_b200_set_option(k.encode(), 0)
print(json.dumps({"text_bytes": int(text.shape[0]), "n_win": out.n_win, "open_stats": None, "rows": rows}, indent=1))
