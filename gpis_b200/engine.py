"""Host-side handle over the C-ABI engine (include/gpis_b200.h).

Arrays may be NumPy (host) or torch CUDA tensors (device); the C ABI takes either
kind of pointer.  Points are given row-major (N, D) as in the reference's MATLAB
structs and converted to the reference's SoA layout pts[i + d*N]
(src/gpu_active_set_selector.cpp:96-101) here.
"""
import ctypes as C

import numpy as np

from . import _lib

try:  # torch is plumbing only: device memory + streams
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def _soa(x, n_cols=None):
    """(N, D) -> contiguous (D, N) float64, NumPy or torch (kept on its device)."""
    if x is None:
        return None
    if _is_torch(x):
        x = x.to(torch.float64)
        if x.dim() == 1:
            x = x[:, None]
        return x.t().contiguous()
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    return np.ascontiguousarray(x.T)


def _vec(x, dtype=np.float64):
    if x is None:
        return None
    if _is_torch(x):
        tdt = {np.float64: torch.float64, np.int64: torch.int64}[dtype]
        return x.to(tdt).contiguous().reshape(-1)
    return np.ascontiguousarray(np.asarray(x, dtype=dtype).reshape(-1))


def _ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(x.ctypes.data)


def _stream():
    if torch is not None and torch.cuda.is_available():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
    return C.c_void_p(0)


class GpisEngine:
    """One engine = one candidate shard + one (replicated) active-set factor."""

    def __init__(self, dim, num_candidates, max_active, *, ell, sf2=1.0, mean_w=None, mean_c=0.0,
                 noise=0.01, level=0.0, eps=1e-2, beta=10.0, use_gradients=False,
                 beta_mode=_lib.BETA_FIXED, beta_delta=0.01, variance_mode=_lib.VAR_LATENT,
                 precision=_lib.PREC_FP64, rank=0, world_size=1, total_candidates=0, id_base=0, device=-1,
                 loop_mode=_lib.LOOP_AUTO, append_mode=_lib.APPEND_AUTO, spin_timeout_ms=0):
        self.lib = _lib.load()
        self.dim, self.N, self.max_active = int(dim), int(num_candidates), int(max_active)
        self.use_gradients = bool(use_gradients)
        self.nb = self.dim + 1 if self.use_gradients else 1
        cfg = _lib.GpisConfig()
        cfg.struct_size = C.sizeof(_lib.GpisConfig)
        cfg.dim, cfg.num_candidates, cfg.max_active = self.dim, self.N, self.max_active
        cfg.total_candidates, cfg.id_base = int(total_candidates), int(id_base)
        cfg.use_gradients, cfg.precision = int(self.use_gradients), int(precision)
        cfg.beta_mode, cfg.variance_mode = int(beta_mode), int(variance_mode)
        cfg.rank, cfg.world_size, cfg.device = int(rank), int(world_size), int(device)
        cfg.ell, cfg.sf2 = float(ell), float(sf2)
        mw = np.zeros(3) if mean_w is None else np.asarray(mean_w, np.float64).reshape(-1)
        for d in range(self.dim):
            cfg.mean_w[d] = float(mw[d])
        cfg.mean_c, cfg.noise = float(mean_c), float(noise)
        cfg.level, cfg.eps, cfg.beta, cfg.beta_delta = float(level), float(eps), float(beta), float(beta_delta)
        cfg.loop_mode, cfg.append_mode, cfg.spin_timeout_ms = int(loop_mode), int(append_mode), int(spin_timeout_ms)
        self.cfg = cfg
        self.peer_mode = False
        self.h = C.c_void_p()
        st = self.lib.gpis_create(C.byref(self.h), C.byref(cfg))
        if st != _lib.GPIS_OK:
            self.h = C.c_void_p()
            _lib.check(self.lib, None, st)

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.gpis_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, st):
        _lib.check(self.lib, self.h, st)

    # ------------------------------------------------------------------ inputs
    def set_candidates(self, points, tsdf, normals=None, noise=None, ids=None, soa=False):
        """soa=True: points / normals are already (D, N) contiguous float64 (no host transpose)."""
        p, t = (points if soa else _soa(points)), _vec(tsdf)
        n, z, i = (normals if soa else _soa(normals)), _vec(noise), _vec(ids, np.int64)
        assert p.shape == (self.dim, self.N), (p.shape, self.dim, self.N)
        self._keep = (p, t, n, z, i)
        self._ck(self.lib.gpis_set_candidates(self.h, _ptr(p), _ptr(t), _ptr(n), _ptr(z), _ptr(i), _stream()))

    def set_grid_candidates(self, dims, origin, spacing, tsdf, normals=None, noise=None, z0=0, nz_total=None,
                            soa=False):
        """Candidates = every point of a regular lattice (x fastest); selects the grid backend."""
        D = self.dim
        dims_a = np.ascontiguousarray(np.asarray(dims, np.int32).reshape(-1)[:D])
        org = np.ascontiguousarray(np.asarray(origin, np.float64).reshape(-1)[:D])
        sp = np.ascontiguousarray(np.asarray(spacing, np.float64).reshape(-1)[:D])
        t, z = _vec(tsdf), _vec(noise)
        n = normals if soa else _soa(normals)
        nzt = int(nz_total) if nz_total is not None else (int(dims_a[2]) + int(z0) if D == 3 else 1)
        self._keep = (dims_a, org, sp, t, n, z)
        self._ck(self.lib.gpis_set_grid_candidates(self.h, _ptr(dims_a), _ptr(org), _ptr(sp), int(z0), nzt,
                                                   _ptr(t), _ptr(n), _ptr(z), _stream()))

    # ------------------------------------------------------------------ selection
    def select(self, first_local, K):
        self._ck(self.lib.gpis_select(self.h, int(first_local), int(K), _stream()))

    def select_begin(self, first_local):
        self._ck(self.lib.gpis_select_begin(self.h, int(first_local), _stream()))

    def step_append(self):
        self._ck(self.lib.gpis_step_append(self.h, _stream()))

    def step_update(self):
        self._ck(self.lib.gpis_step_update(self.h, _stream()))

    def select_end(self):
        self._ck(self.lib.gpis_select_end(self.h, _stream()))

    def exchange_buffers(self):
        s, r, b = C.c_void_p(), C.c_void_p(), C.c_int64()
        self._ck(self.lib.gpis_exchange_buffers(self.h, C.byref(s), C.byref(r), C.byref(b)))
        return s.value, r.value, b.value

    def peer_export(self):
        buf = (C.c_ubyte * 64)()
        self._ck(self.lib.gpis_peer_export(self.h, C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def peer_import(self, handles):
        """handles: list of world_size 64-byte IPC handles, rank-major."""
        blob = b"".join(handles)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._ck(self.lib.gpis_peer_import(self.h, C.cast(buf, C.c_void_p), 0))
        self.peer_mode = True

    def active(self):
        ids = np.zeros(self.max_active + 1, np.int64)
        ns, nm = C.c_int32(), C.c_int32()
        self._ck(self.lib.gpis_get_active(self.h, _ptr(ids), C.byref(ns), C.byref(nm), _stream()))
        return ids[:ns.value].copy(), nm.value

    def factor(self):
        rows = C.c_int32()
        self._ck(self.lib.gpis_get_factor(self.h, None, 0, None, C.byref(rows), _stream()))
        R = rows.value
        L, a = np.zeros((R, R)), np.zeros(R)
        if R:
            self._ck(self.lib.gpis_get_factor(self.h, _ptr(L), R, _ptr(a), C.byref(rows), _stream()))
        return L, a

    def fit(self, points, tsdf, normals=None, noise=None):
        p, t, n, z = _soa(points), _vec(tsdf), _soa(normals), _vec(noise)
        m = t.shape[0]
        self._ck(self.lib.gpis_fit(self.h, _ptr(p), _ptr(t), _ptr(n), _ptr(z), m, _stream()))

    # ------------------------------------------------------------------ outputs
    def predict_into(self, query_soa, mean, var, normals_soa=None):
        """Raw form: (D, nq) contiguous query, preallocated outputs (host or device)."""
        nq = query_soa.shape[1]
        self._ck(self.lib.gpis_predict(self.h, _ptr(query_soa), nq, _ptr(mean), _ptr(var), _ptr(normals_soa),
                                       _stream()))

    def predict(self, query, want_normals=False, out_torch=None):
        q = _soa(query)
        nq = q.shape[1]
        on_dev = _is_torch(q) and q.is_cuda if out_torch is None else out_torch
        if on_dev:
            dev = q.device if _is_torch(q) else torch.device("cuda")
            mean = torch.empty(nq, dtype=torch.float64, device=dev)
            var = torch.empty(nq, dtype=torch.float64, device=dev)
            nrm = torch.empty((self.dim, nq), dtype=torch.float64, device=dev) if want_normals else None
        else:
            mean, var = np.empty(nq), np.empty(nq)
            nrm = np.empty((self.dim, nq)) if want_normals else None
        self._ck(self.lib.gpis_predict(self.h, _ptr(q), nq, _ptr(mean), _ptr(var), _ptr(nrm), _stream()))
        if want_normals:
            nrm = nrm.t() if on_dev else nrm.T
        return mean, var, nrm

    def posterior_cov(self, query, use_derivative=False, want_cov=True):
        """gp_mean.m + the FULL gp_cov.m at the query points: (MEAN [nv], COV [nv, nv]); values are
        dimension-major [f(all q); d1 f(all q); ..] like the reference's joint vector."""
        q = _soa(query)
        nq = q.shape[1]
        nv = nq * (self.dim + 1 if use_derivative else 1)
        on_dev = _is_torch(q) and q.is_cuda
        if on_dev:
            mean = torch.empty(nv, dtype=torch.float64, device=q.device)
            cov = torch.empty((nv, nv), dtype=torch.float64, device=q.device) if want_cov else None
        else:
            mean, cov = np.empty(nv), (np.empty((nv, nv)) if want_cov else None)
        self._ck(self.lib.gpis_posterior_cov(self.h, _ptr(q), nq, int(bool(use_derivative)), _ptr(mean), _ptr(cov),
                                             nv, _stream()))
        return mean, cov

    def sample_shapes(self, query, z, use_derivative=False, jitter=1e-12, want_chol=False):
        """sample_shapes.m:14-31 at the query points with the standard-normal draws z [S, nv] made
        explicit: (samples [S, nv], logpdf [S][, T upper with T'T = COV + jitter I])."""
        q = _soa(query)
        nq = q.shape[1]
        nv = nq * (self.dim + 1 if use_derivative else 1)
        on_dev = _is_torch(z) and z.is_cuda
        if on_dev:
            z = z.to(torch.float64).contiguous().reshape(-1, nv)
            ns = z.shape[0]
            samples = torch.empty((ns, nv), dtype=torch.float64, device=z.device)
            logpdf = torch.empty(ns, dtype=torch.float64, device=z.device)
            chol = torch.empty((nv, nv), dtype=torch.float64, device=z.device) if want_chol else None
        else:
            z = np.ascontiguousarray(np.asarray(z, np.float64).reshape(-1, nv))
            ns = z.shape[0]
            samples, logpdf = np.empty((ns, nv)), np.empty(ns)
            chol = np.empty((nv, nv)) if want_chol else None
        self._ck(self.lib.gpis_sample_shapes(self.h, _ptr(q), nq, int(bool(use_derivative)), ns, _ptr(z),
                                             float(jitter), _ptr(samples), _ptr(logpdf), _ptr(chol), nv, _stream()))
        return (samples, logpdf, chol) if want_chol else (samples, logpdf)

    def sampling_timing(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._ck(self.lib.gpis_get_sampling_timing(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"cov_ms": a.value, "chol_ms": b.value, "draw_ms": c.value}

    def levelset(self):
        amb = np.empty(self.N)
        hi, lo, ac = (np.empty(self.N, np.uint8) for _ in range(3))
        self._ck(self.lib.gpis_get_levelset(self.h, _ptr(amb), _ptr(hi), _ptr(lo), _ptr(ac), _stream()))
        return amb, hi.astype(bool), lo.astype(bool), ac.astype(bool)

    def candidate_posterior(self):
        mu, var = np.empty(self.N), np.empty(self.N)
        self._ck(self.lib.gpis_get_candidate_posterior(self.h, _ptr(mu), _ptr(var), _stream()))
        return mu, var

    def candidate_posterior_into(self, mean, var):
        """Raw form: preallocated host or device outputs."""
        self._ck(self.lib.gpis_get_candidate_posterior(self.h, _ptr(mean), _ptr(var), _stream()))

    def stats(self):
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self.lib.gpis_get_stats(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"kernel_launches": a.value, "live_tiles": b.value, "iterations": c.value, "panel_bytes": d.value}

    def timing(self):
        a, b, c, n = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        self._ck(self.lib.gpis_get_timing(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(n)))
        p, q, r = C.c_double(), C.c_double(), C.c_double()
        self._ck(self.lib.gpis_get_phase_timing(self.h, C.byref(p), C.byref(q), C.byref(r)))
        return {"select_ms": a.value, "predict_ms": b.value, "update_ms": c.value, "update_launches": n.value,
                "append_ms": p.value, "main_ms": q.value, "epilogue_ms": r.value}

    PERSIST_SLOTS = ("record_wait", "w_setup", "w_pass", "barrier1", "new_row", "barrier2", "prefix", "gemm", "epilogue",
                     "barrier3", "winner", "steps", "gemm_precompute", "gemm_partial_tiles", "gemm_epilogue_regs", "units")

    def persist_timing(self):
        """Lap times (ms, block 0's clock) of the persistent selection kernel in the last select()."""
        a = np.zeros(16)
        self._ck(self.lib.gpis_get_persist_timing(self.h, _ptr(a)))
        return dict(zip(self.PERSIST_SLOTS, a.tolist()))

    def persist_balance(self):
        """Per-block (time in the update phase, of that waiting for operand data), ms, of the last persistent select()."""
        n = C.c_int32()
        self._ck(self.lib.gpis_get_persist_balance(self.h, None, 0, C.byref(n)))
        a = np.zeros(2 * max(n.value, 1))
        if n.value:
            self._ck(self.lib.gpis_get_persist_balance(self.h, _ptr(a), a.size, C.byref(n)))
        return a.reshape(-1, 2)[:n.value]

    def set_profiling(self, on=True):
        self._ck(self.lib.gpis_set_profiling(self.h, int(bool(on))))

    def history(self):
        cap = self.max_active + 1
        rows, tiles, scored = (np.zeros(cap, np.int32) for _ in range(3))
        n = C.c_int32()
        self._ck(self.lib.gpis_get_history(self.h, _ptr(rows), _ptr(tiles), _ptr(scored), C.byref(n)))
        k = n.value
        return {"rows": rows[:k].copy(), "live_tiles": tiles[:k].copy(), "scored": scored[:k].copy()}
