"""Measure the shared-memory atomic rate of the GPU (roofline denominator for the counting step)."""
import ctypes as C
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frisk_b200 import _lib

L = _lib.lib()
_lib.require_device()
out = {}
for mode, name in [(0, "conflict_free"), (1, "random"), (2, "single_address")]:
    best = None
    for _ in range(3):
        ms = C.c_float(0)
        blocks, iters = 148 * 2, 4096 if mode < 2 else 512
        _lib.check(L.frisk_b200_bench_smem_atomics(blocks, iters, mode, C.byref(ms), None), "bench")
        rate = blocks * 1024 * iters / (ms.value * 1e-3)
        best = rate if best is None else max(best, rate)
    out[name] = best
print(json.dumps({"smem_atomic_updates_per_s": out}))
