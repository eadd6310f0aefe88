"""Window-kernel A/B on C2 (40 Mbp, 15,806 windows): nibble (frisk_nibble.cu) vs bucketed vs direct kernel, kmax 8 and 7,
-w 5000 / 2000 / 8000.  Score-kernel time per run (CUDA events, L2 flushed, median of 7) and the largest relative KLD
difference between the kernels (they must agree to ~1e-12: same integers, different summation order)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from frisk_b200 import engine, synth, _lib
g = engine.PackedGenome.from_scaffolds(synth.make("C2", 1.0, seed=2002))
dq = engine.DeviceGenome(g)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
d_tables, _ = engine.finalize(engine.background(dq, 8), 8)
KERNELS = (("nibble", b"force_nibble_kernel"), ("bucket", b"force_bucket_kernel"), ("direct", b"force_direct_kernel"))
only = os.environ.get("FRISK_AB_ONLY", "").split(",") if os.environ.get("FRISK_AB_ONLY") else None
shapes = ((5000, 2500), (2000, 1000), (8000, 4000)) if not os.environ.get("FRISK_AB_QUICK") else ((5000, 2500),)
if os.environ.get("FRISK_AB_SHAPES"):                      # e.g. "1000:500,500:250"
    shapes = tuple(tuple(int(x) for x in item.split(":")) for item in os.environ["FRISK_AB_SHAPES"].split(","))
out = []
for (w, step) in shapes:
    wins = g.windows(w, step, False)
    n = len(wins)
    d_off = torch.from_numpy(wins.off.view(np.int64)).cuda(); d_len = torch.from_numpy(wins.length.view(np.int32)).cuda()
    for k in (8, 7):
        d_ig = engine.genome_ivom(d_tables[:_lib.table_size(1, k)], 1, k, g.genome_space)
        ref_rows = None
        for name, opt in KERNELS:
            if only and name not in only:
                continue
            _lib.check(_lib.lib().frisk_b200_set_option(opt, 1), "opt")
            ts = []
            d_rows = torch.empty((n, 5), dtype=torch.float64, device="cuda"); d_status = torch.empty(n, dtype=torch.int32, device="cuda")
            P = engine._ptr
            for rep in range(10):
                flush.fill_(1)
                a = ev()
                _lib.check(_lib.lib().frisk_b200_score(P(dq.codes), P(dq.inv), P(dq.low), P(d_off), P(d_len), n, wins.max_len, P(d_ig), 1, k, 1,
                                                       P(d_rows), P(d_status), None, engine._stream_ptr(dq.device)), "score")
                b = ev()
                torch.cuda.synchronize()
                if rep >= 3: ts.append(a.elapsed_time(b))
            _lib.lib().frisk_b200_set_option(opt, 0)
            rows = d_rows.cpu().numpy(); st = d_status.cpu().numpy()
            diff = None
            if ref_rows is None:
                ref_rows, ref_st = rows, st
            else:
                assert np.array_equal(st, ref_st), "status differs between kernels"
                ok = st == 0
                diff = float(np.max(np.abs(rows[ok, 0] - ref_rows[ok, 0]) / np.maximum(np.abs(ref_rows[ok, 0]), 1e-300)))
                assert np.array_equal(rows[ok, 1:], ref_rows[ok, 1:], equal_nan=True)
            rec = {"w": w, "kmax": k, "kernel": name, "ms": float(np.median(ts)), "windows": n, "kld_max_rel_vs_first": diff}
            out.append(rec)
            print(json.dumps(rec), flush=True)
