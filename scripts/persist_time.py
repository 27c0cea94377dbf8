"""Development timing (GPU): the persistent selection kernel's phase table on one configuration.
GPIS_LIB=<path to an experimental build of the library> selects another .so (kernel experiments)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpis_b200 import _lib
if os.environ.get("GPIS_LIB"):
    _lib.LIB_PATH = os.environ["GPIS_LIB"]
from gpis_b200 import GpisEngine
from gpis_b200.synth import box_tsdf, first_index

g = int(sys.argv[1]) if len(sys.argv) > 1 else 128
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
pts, tsdf = box_tsdf(g)
N = g ** 3
eng = GpisEngine(3, N, K, ell=4.0, noise=0.05, level=0.0, eps=1e-2, beta=10.0, spin_timeout_ms=10000)
eng.set_grid_candidates((g, g, g), (0, 0, 0), (1, 1, 1), tsdf)
for rep in range(2):
    eng.select(first_index(N), K)
t = eng.timing()
ids, nm = eng.active()
bal = eng.persist_balance()
if len(bal):
    import numpy as np
    o = np.argsort(bal[:, 0])
    print("phase-D ms per block: min %.1f (block %d) median %.1f max %.1f (block %d); wait-for-data ms: min %.1f median %.1f max %.1f" % (
        bal[o[0], 0], o[0], np.median(bal[:, 0]), bal[o[-1], 0], o[-1], bal[:, 1].min(), np.median(bal[:, 1]), bal[:, 1].max()), file=sys.stderr)
    print("slowest 8 blocks:", [(int(b), round(bal[b, 0], 1), round(bal[b, 1], 1)) for b in o[-8:]], file=sys.stderr)
    print("fastest 8 blocks:", [(int(b), round(bal[b, 0], 1), round(bal[b, 1], 1)) for b in o[:8]], file=sys.stderr)
print(json.dumps({"lib": os.path.basename(_lib.LIB_PATH), "grid": g, "K": K, "n": int(len(ids)), "select_ms": t["select_ms"],
                  "phase_ms": {k: round(v, 2) for k, v in eng.persist_timing().items()},
                  "ids_crc32": int(__import__("zlib").crc32(ids.tobytes()))}))
