"""Golden values of the reference's threshold functions (FDBins, otsu, setKLDThresh; F:508-543, F:664-690),
produced by EXECUTING THE REFERENCE'S OWN SOURCE TEXT (oracle/ref_exec.load_thresholds) on the window KLDs
of the committed golden fixtures.  Container-only; writes tests/golden/thresholds.json.

    python tests/golden/make_thresholds.py
"""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_exec  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = ref_exec.load_thresholds()
    out = {}
    for case in ("c1_full", "c1_small", "c2_small", "edge_default"):
        kld = np.load(os.path.join(HERE, case + ".npz"))["row_vals"][:, 0]
        kld = kld[np.isfinite(kld) & (kld > 0)]
        log = np.log10(kld[:, np.newaxis])                       # (n, 1): what main() passes (F:1530-1532)
        rec = {"n": int(len(kld)), "FDBins": int(ref.FDBins(log))}
        for kind, extra in (("otsu", {}), ("percentile", {"percentileKLD": 99.0}), ("percentile", {"percentileKLD": 90.0})):
            args = types.SimpleNamespace(threshTypeKLD=kind, forceThresholdKLD=None, percentileKLD=extra.get("percentileKLD", 99.0))
            thr, bins = ref.setKLDThresh(args, log)
            rec["%s%s" % (kind, extra.get("percentileKLD", ""))] = [float(np.ravel(thr)[0]), int(bins)]
        args = types.SimpleNamespace(threshTypeKLD=None, forceThresholdKLD="0.35", percentileKLD=99.0)
        thr, bins = ref.setKLDThresh(args, log)
        rec["force0.35"] = [float(np.ravel(thr)[0]), int(bins)]
        rec["otsu_40bins"] = float(np.ravel(ref.otsu(log, 40))[0])
        out[case] = rec
    with open(os.path.join(HERE, "thresholds.json"), "w") as fh:
        json.dump({"ref_sha256": ref_exec.REF_SHA256, "cases": out}, fh, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
