"""frisk_b200: B200-native implementation of the hot path of Adamtaranto/frisk (sliding-window
k-mer counting + per-window IVOM/KLD composition scoring), behind frisk's own function surface.

    import frisk_b200 as frisk          # or: import frisk   (shim package at the repo root)
    frisk.main()                        # same CLI: python -m frisk_b200 -H genome.fa --exitAfter WindowKLD

`frisk_b200.engine` is the batch API (PackedGenome, Pipeline, run, run_host); `frisk_b200.api`
mirrors the reference's functions.  All arithmetic of the path runs in libfrisk_b200.so (CUDA,
sm_100a); there is no CPU fallback.
"""
from .api import (FRISK_VERSION, IvomBuild, KLD, LETTERS, calcGC, calcRIP, computeKmers, countN, crawlGenome,  # noqa: F401
                  iterFasta, main, mainArgs, makePicklePath, prepareMaps, rangeMaps, revComplement, score_genome,
                  tempPathCheck)

__version__ = FRISK_VERSION
