"""Batch of independent GPIS fits (BASELINE.json config 5: a Dex-Net-style shape batch).

Objects are independent, so there is no collective and nothing to shard inside an object:
each of `n_workers` host threads owns one engine on its own CUDA stream and works through its
share of the objects (the C ABI releases the GIL for the duration of a call, the greedy loop of
each object is a captured CUDA graph).  With N GPUs every rank takes objects rank::N ("replicas
only" -- DESIGN.md section 6)."""
import os
import threading

import numpy as np

from .engine import GpisEngine

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def superellipsoid_tsdf(g, seed, trunc=3.0, jitter=1e-3):
    """Object k of config 5: superellipsoid with exponents / axes drawn from default_rng(k)."""
    rng = np.random.default_rng(seed)
    ax = rng.uniform(0.25, 0.42, 3) * g
    ex = rng.uniform(0.6, 2.0, 2)
    c = (g - 1) / 2.0 + rng.uniform(-1.5, 1.5, 3)
    idx = np.indices((g, g, g)).reshape(3, -1)[::-1].T.astype(np.float64)        # x fastest
    p = np.abs(idx - c) / ax
    f = ((p[:, 0] ** (2 / ex[1]) + p[:, 1] ** (2 / ex[1])) ** (ex[1] / ex[0]) + p[:, 2] ** (2 / ex[0])) ** (ex[0] / 2)
    sd = (f - 1.0) * ax.min()
    return np.clip(sd / trunc, -1, 1) + jitter * rng.standard_normal(idx.shape[0])


def select_batch(tsdfs, dims, K, first_index, n_workers=8, rank=0, world=1, keep_posterior=True, **cfg):
    """tsdfs: sequence of per-object TSDF arrays on the same lattice `dims` (x fastest).
    Returns a list (one entry per object handled by this rank) of
    dict(object, ids, n_in_model[, mean, var])."""
    D = len(dims)
    n = int(np.prod(dims))
    mine = list(range(rank, len(tsdfs), world))
    results = [None] * len(mine)
    errors = []
    n_workers = max(1, min(n_workers, len(mine)))
    # Engines of a batch run concurrently on one device.  The single-engine fast path fuses two
    # kernels behind an in-kernel grid barrier, which needs ALL its CTAs resident at once -- not
    # guaranteed when several engines' kernels share the SMs -- so the batch uses the unfused
    # kernels (read by gpis_create; see include/gpis_b200.h).
    prev = os.environ.get("GPIS_SOLVE_REGS")
    os.environ["GPIS_SOLVE_REGS"] = "1"

    def work(w):
        try:
            stream = torch.cuda.Stream() if torch is not None and torch.cuda.is_available() else None
            ctx = torch.cuda.stream(stream) if stream is not None else _Null()
            with ctx:
                eng = GpisEngine(D, n, K, **cfg)
                for slot in range(w, len(mine), n_workers):
                    k = mine[slot]
                    eng.set_grid_candidates(dims, (0.0,) * D, (1.0,) * D, tsdfs[k])
                    eng.select(first_index, K)
                    ids, nm = eng.active()
                    out = {"object": k, "ids": ids, "n_in_model": nm}
                    if keep_posterior:
                        out["mean"], out["var"] = eng.candidate_posterior()
                    results[slot] = out
                eng.close()
        except Exception as e:  # surfaced to the caller below
            errors.append(e)

    threads = [threading.Thread(target=work, args=(w,)) for w in range(n_workers)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if prev is None:
        os.environ.pop("GPIS_SOLVE_REGS", None)
    else:
        os.environ["GPIS_SOLVE_REGS"] = prev
    if errors:
        raise errors[0]
    return results


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
