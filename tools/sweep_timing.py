"""C3 (120 Mbp) k sweep 1..8: time of the shared background pass and of each k' (genome IVOM + window kernel)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frisk_b200 import engine, synth, _lib
g = engine.PackedGenome.from_scaffolds(synth.make("C3", 1.0))
dq = engine.DeviceGenome(g)
wins = g.windows(5000, 2500, False)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for rep in range(2):
    e0 = ev()
    d_tables, _ = engine.finalize(engine.background(dq, 8), 8)
    e1 = ev()
    times = []
    for k in range(1, 9):
        a = ev()
        d_ig = engine.genome_ivom(d_tables[:_lib.table_size(1, k)], 1, k, g.genome_space)
        engine.score(dq, wins, d_ig, 1, k, True)
        times.append((a, ev()))
    torch.cuda.synchronize()
    if rep:
        print("background+finalise %.3f ms" % e0.elapsed_time(e1))
        tot = 0
        for k, (a, b) in enumerate(times, start=1):
            t = a.elapsed_time(b); tot += t
            print("k'=%d: %.3f ms" % (k, t))
        print("sweep total %.2f ms for %d windows x 8 = %.1f Gbp/s per k-run" % (tot + e0.elapsed_time(e1), len(wins), 8 * g.total_len / (tot + e0.elapsed_time(e1)) / 1e6))
