"""Batch of independent GPIS fits (BASELINE.json config 5: a Dex-Net-style shape batch).

Objects are independent, so there is no collective and nothing to shard inside an object:
each of `n_workers` host threads owns one engine on its own CUDA stream and works through its
share of the objects (the C ABI releases the GIL for the duration of a call, the greedy loop of
each object is a captured CUDA graph).  With N GPUs every rank takes objects rank::N ("replicas
only" -- DESIGN.md section 6).

`select_batch_fused` is the object-batched form for small lattices (<= 32 per axis, <= 256 points,
function values, fixed beta): ONE launch in which every thread block runs the whole selection of an
object (gpis_batch_select, gpis_b200/csrc/gpis_batch.cuh)."""
import ctypes as C
import os
import threading

import numpy as np

from . import _lib
from .engine import GpisEngine, _ptr, _stream

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
    # Engines of a batch run concurrently on one device.  The single-engine fast paths (the persistent
    # selection kernel, the fused append) hold every SM behind in-kernel grid barriers, which is not
    # possible when several engines' kernels share the SMs -- so the batch asks for the plain per-step
    # kernels in each engine's config (include/gpis_b200.h: loop_mode, append_mode).
    cfg = dict(cfg, loop_mode=_lib.LOOP_STEPS, append_mode=_lib.APPEND_TWO_KERNELS)

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
    if errors:
        raise errors[0]
    return results


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def select_batch_fused(tsdfs, dims, K, first_index, rank=0, world=1, keep_posterior=True, *, ell, sf2=1.0,
                       mean_w=None, mean_c=0.0, noise=0.01, level=0.0, eps=1e-2, beta=10.0,
                       variance_mode=_lib.VAR_LATENT, device=-1, kernel=_lib.BATCH_AUTO):
    """Same inputs and result structure as `select_batch` (list of dict(object, ids, n_in_model[, mean,
    var, flags])), through gpis_batch_select: one kernel launch for this rank's share of the objects.
    `tsdfs` may be a sequence of per-object arrays or one (n_objects, N) array (NumPy or torch CUDA).
    The result's "device_ms" (on every entry) is the CUDA-event time of that launch."""
    lib = _lib.load()
    D = len(dims)
    n = int(np.prod(dims))
    mine = list(range(rank, len(tsdfs), world))
    if not mine:
        return []
    on_dev = torch is not None and isinstance(tsdfs, torch.Tensor) and tsdfs.is_cuda
    if on_dev:
        t = tsdfs[mine].to(torch.float64).contiguous().reshape(len(mine), n)
    else:
        t = np.ascontiguousarray(np.stack([np.asarray(tsdfs[k], np.float64).reshape(-1) for k in mine]))
    B = len(mine)
    cfg = _lib.GpisConfig()
    cfg.struct_size = C.sizeof(_lib.GpisConfig)
    cfg.dim, cfg.num_candidates, cfg.max_active = D, n, int(K)
    cfg.precision, cfg.beta_mode, cfg.variance_mode = _lib.PREC_FP64, _lib.BETA_FIXED, int(variance_mode)
    cfg.world_size, cfg.device = 1, int(device)
    cfg.ell, cfg.sf2 = float(ell), float(sf2)
    mw = np.zeros(3) if mean_w is None else np.asarray(mean_w, np.float64).reshape(-1)
    for d in range(D):
        cfg.mean_w[d] = float(mw[d])
    cfg.mean_c, cfg.noise = float(mean_c), float(noise)
    cfg.level, cfg.eps, cfg.beta = float(level), float(eps), float(beta)
    cfg.batch_kernel = int(kernel)
    dims_a = np.ascontiguousarray(np.asarray(dims, np.int32))
    ids = np.empty((B, int(K)), np.int64)
    n_sel, n_model = np.empty(B, np.int32), np.empty(B, np.int32)
    mean = var = flags = None
    if keep_posterior:
        mean, var, flags = np.empty((B, n)), np.empty((B, n)), np.empty((B, n), np.uint8)
    ms = C.c_double(0.0)
    st = lib.gpis_batch_select(C.byref(cfg), _ptr(dims_a), None, None, B, _ptr(t), int(first_index), int(K),
                               _ptr(ids), _ptr(n_sel), _ptr(n_model), _ptr(mean), _ptr(var), _ptr(flags),
                               C.byref(ms), _stream())
    _lib.check(lib, None, st)
    out = []
    for i, k in enumerate(mine):
        r = {"object": k, "ids": ids[i, :n_sel[i]].copy(), "n_in_model": int(n_model[i]), "device_ms": ms.value}
        if keep_posterior:
            r["mean"], r["var"], r["flags"] = mean[i], var[i], flags[i]
        out.append(r)
    return out
