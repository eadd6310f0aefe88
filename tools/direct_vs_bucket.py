"""Window-kernel A/B on C2 (40 Mbp, 15,806 windows): direct (frisk_direct.cu) vs bucketed kernel, kmax 7 and 8,
and -w 2000 / -w 8000 windows.  Prints the score-kernel time per run (CUDA events, L2 flushed, median of 7)."""
import sys, os
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
for (w, step) in ((5000, 2500), (2000, 1000), (8000, 4000)):
    wins = g.windows(w, step, False)
    for k in (8, 7):
        d_ig = engine.genome_ivom(d_tables[:_lib.table_size(1, k)], 1, k, g.genome_space)
        only = os.environ.get("FRISK_AB_ONLY")
        for name, opts in (("direct", {b"force_direct_kernel": 1}), ("bucket", {b"force_bucket_kernel": 1}), ("default", {})):
            if only and name != only:
                continue
            for o, v in opts.items():
                _lib.check(_lib.lib().frisk_b200_set_option(o, v), "opt")
            ts = []
            n = len(wins)
            d_off = torch.from_numpy(wins.off.view(np.int64)).cuda(); d_len = torch.from_numpy(wins.length.view(np.int32)).cuda()
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
            _lib.lib().frisk_b200_set_option(b"force_bucket_kernel", 0)
            _lib.lib().frisk_b200_set_option(b"force_direct_kernel", 0)
            print("w=%d kmax=%d %-10s %.4f ms (%d windows)" % (w, k, name, float(np.median(ts)), len(wins)), flush=True)
