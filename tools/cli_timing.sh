#!/bin/bash
# wall-clock of the reference's CLI surface on C2 (40 Mbp FASTA on local disk)
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from frisk_b200 import synth
synth.write_fasta(synth.make("C2", 1.0), "/tmp/c2.fa")
PY
for i in 1 2; do
  python - <<PY
import subprocess, time, sys
t = time.time()
r = subprocess.run([sys.executable, "-m", "frisk", "-H", "/tmp/c2.fa", "-t", "/tmp/frisk_tmp$i", "--quiet", "--RIP"], capture_output=True, text=True)
print("python -m frisk (40 Mbp, 15,806 rows, --RIP): %.2f s, exit %d" % (time.time() - t, r.returncode))
PY
done
python - <<'PY'
import time, sys, os, cProfile, pstats, io
sys.path.insert(0, os.getcwd())
t = time.time()
import frisk
print("import frisk %.2f s" % (time.time() - t))
pr = cProfile.Profile(); pr.enable()
t = time.time()
frisk.main(["-H", "/tmp/c2.fa", "-t", "/tmp/frisk_tmp3", "--quiet", "--RIP"])
print("main() %.2f s; torch loaded: %s" % (time.time() - t, "torch" in sys.modules))
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3000])
PY
