"""Drop-in shim: `import frisk` / `python -m frisk` resolve to the B200 implementation of the hot
path (see frisk_b200).  Mirrors the reference's module-level names used by its main() (F:1400-1507)."""
from frisk_b200.api import *  # noqa: F401,F403
from frisk_b200.api import FRISK_VERSION, main  # noqa: F401

__version__ = FRISK_VERSION
