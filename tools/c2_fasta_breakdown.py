"""Where the time of the C2 FASTA-text path (bench.py's e2e_fasta) goes on one GPU: wall clock per stage with a device
synchronisation after each (so stages do not overlap here; the sum is an upper bound of the pipelined call)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
import bench
from frisk_b200 import engine, _lib, synth

dev = torch.device("cuda:0")
scaffolds = synth.make("C2", 1.0)
raw = np.frombuffer(synth.fasta_bytes(scaffolds), dtype=np.uint8)
text = engine._alloc(raw.shape[0], np.uint8, True)
text[:] = raw
params = dict(bench.PARAMS)
L = _lib.lib()
rows = []
for rep in range(8):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    with torch.cuda.device(dev):
        st = engine._stream_ptr(dev)
        h = C.c_void_p(); nrec, padded = C.c_uint64(0), C.c_uint64(0); stats = np.zeros(3, np.uint64)
        _lib.check(L.frisk_b200_fasta_open(engine._ptr(text), text.shape[0], st, C.byref(h), C.byref(nrec), C.byref(padded),
                                           engine._ptr(stats)), "open")
        t.append(time.perf_counter())
        R, P = int(nrec.value), int(padded.value)
        name_off = np.zeros(R, np.uint64); name_len = np.zeros(R, np.uint32)
        seq_len = np.zeros(R, np.uint64); scaf_off = np.zeros(R, np.uint64)
        L.frisk_b200_fasta_records(h, engine._ptr(name_off), engine._ptr(name_len), engine._ptr(seq_len), engine._ptr(scaf_off))
        codes = torch.empty(P // 16, dtype=torch.int32, device=dev)
        inv = torch.empty(P // 32, dtype=torch.int32, device=dev)
        low = torch.empty(P // 32, dtype=torch.int32, device=dev) if int(stats[2]) else None
        t.append(time.perf_counter())
        L.frisk_b200_fasta_pack(h, engine._ptr(codes), engine._ptr(inv), engine._ptr(low), st)
        torch.cuda.synchronize(); t.append(time.perf_counter())
        L.frisk_b200_fasta_close(h, st)
        g = engine.PackedGenome(engine.LazyNames(text, name_off, name_len), seq_len, scaf_off, P, None, None, None,
                                int(stats[0]), int(stats[1]), int(stats[2]), False)
        dq = engine.DeviceGenome(g, dev, planes=(codes, inv, low))
        t.append(time.perf_counter())
        wins = g.windows(params["w"], params["step"], False); t.append(time.perf_counter())
        out = engine.HostOutputs(len(wins), params["kmax"]); t.append(time.perf_counter())
        engine.run_resident(dq, None, wins=wins, out=out, assemble_result=False, **params)
        torch.cuda.synchronize(); t.append(time.perf_counter())
    ms = (C.c_float * 8)(); nn = C.c_int(0)
    L.frisk_b200_last_run_timing(ms, 8, C.byref(nn))
    d = dict(zip(["open (H2D + tokenise + record table)", "records + plane alloc (host)", "pack (device, synced)",
                  "close + genome objects (host)", "windows() (host)", "HostOutputs alloc (pinned; reused in the bench)",
                  "run_resident (one C call)"],
                 [round((b - a) * 1e3, 3) for a, b in zip(t, t[1:])]))
    d["run_resident marks ms"] = [round(float(x), 3) for x in ms[:nn.value]]
    d["total_ms"] = round((t[-1] - t[0]) * 1e3, 3)
    rows.append(d)
print(json.dumps({"text_bytes": int(text.shape[0]), "records": R, "reps": rows[3:]}, indent=1))
