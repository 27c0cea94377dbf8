"""Multi-GPU host logic: candidates (and query points) are sharded across ranks in contiguous
index ranges, the active set and its factor are replicated, and each selection step exchanges
ONE fixed-size record per rank (best local candidate: score, id, point, its column of L^-1 K)
with an all-gather.  Every rank then appends the same winner (max score, lowest id) to its
replica of the factor, bit-identically -- no other collective is on the path.

The reference has no multi-GPU code at all; this is new design.
torch.distributed is plumbing only: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def shard_range(n, rank, world):
    """Contiguous split of n candidates in whole 64-candidate tiles (the panel granularity),
    remainder to the last ranks' tiles; returns [lo, hi)."""
    tiles = (n + 63) // 64
    per, extra = divmod(tiles, world)
    t0 = rank * per + min(rank, extra)
    t1 = t0 + per + (1 if rank < extra else 0)
    return min(t0 * 64, n), min(t1 * 64, n)


def slab_range(nz, rank, world):
    """z-slabs of a lattice for the grid backend: [z0, z1)."""
    per, extra = divmod(nz, world)
    z0 = rank * per + min(rank, extra)
    return z0, z0 + per + (1 if rank < extra else 0)


def slab_range_balanced(nz, rank, world, ell, spacing=1.0, fixed=0.35, align=1):
    """z-slabs of a lattice with about equal WORK per rank for the persistent selection kernel: a slice only
    sums the model rows within 9.85 ell of it (gpis_grid_persist.cuh), so with rows spread over the volume the
    slices near the two ends of the z axis cost about half of a middle slice.  Weight of slice z = the part of
    [z - c, z + c] inside the lattice (c = 9.85 ell / spacing slices) + `fixed` x 2c for the per-slice work
    that does not depend on the rows (epilogue, state traffic); slabs are contiguous and cut where the running
    weight crosses k / world of the total.  Every rank gets at least one slice.  `align` = 16 moves the cuts to
    multiples of 16 slices when every rank can have 16: the persistent kernel works on 16 x 16 x 16 cubes of a 3-D
    slab, and a slab of 19 slices would pay for 32.  Returns [z0, z1)."""
    c = 9.85 * float(ell) / float(spacing)
    z = np.arange(nz, dtype=np.float64)
    w = np.minimum(z + c, nz - 1) - np.maximum(z - c, 0.0) + fixed * 2.0 * c
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for k in range(1, world):
        zc = int(np.searchsorted(cum, cum[-1] * k / world))
        zc = min(max(zc, cuts[-1] + 1), nz - (world - k))
        cuts.append(zc)
    cuts.append(nz)
    if align > 1 and nz >= align * world:
        al = [0]
        for k in range(1, world):
            zc = int(round(cuts[k] / align)) * align
            al.append(min(max(zc, al[-1] + align), nz - align * (world - k)))
        cuts = al + [nz]
    return cuts[rank], cuts[rank + 1]


def owner_of(index, n, world):
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        if lo <= index < hi:
            return r, index - lo
    raise IndexError(index)


class _CudaView:
    """Expose an engine-owned device buffer to torch through the CUDA array interface."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def engine_exchange_tensors(eng, world):
    send_p, recv_p, nbytes = eng.exchange_buffers()
    dev = torch.device("cuda", torch.cuda.current_device())
    send = torch.as_tensor(_CudaView(send_p, nbytes), device=dev)
    recv = send if world == 1 else torch.as_tensor(_CudaView(recv_p, nbytes * world), device=dev)
    return send, recv


class ShardedSelector:
    """Drives the greedy loop.  `eng` needs select_begin / step_append / step_update / select_end
    (+ select for world == 1); `exchange` returns (send, recv) tensors for the all-gather."""

    def __init__(self, eng, rank, world, lo, hi, exchange=None, group=None):
        self.eng, self.rank, self.world, self.lo, self.hi = eng, rank, world, lo, hi
        self.group = group
        self.last_select_ms = 0.0
        self.last_update_ms = 0.0
        self.profile = False
        self._ex = exchange
        self._bufs = None

    def _exchange(self):
        if self._bufs is None:
            self._bufs = self._ex(self.eng, self.world) if self._ex else engine_exchange_tensors(self.eng, self.world)
        return self._bufs

    def enable_peer_exchange(self):
        """Replace the per-step all-gather by P2P stores over NVLink from inside the kernels:
        exchange CUDA IPC handles of the engines' exchange regions once, then the whole loop
        runs in gpis_select (captured CUDA graph, no host call per step)."""
        mine = self.eng.peer_export()
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=self.group)
        self.eng.peer_import(handles)
        dist.barrier(group=self.group)

    def _gather(self):
        send, recv = self._exchange()
        dist.all_gather_into_tensor(recv, send, group=self.group)

    def select(self, first_index, K):
        """first_index is GLOBAL.  Returns nothing; read results from the engine."""
        if self.world == 1 or getattr(self.eng, "peer_mode", False):
            mine = self.lo <= first_index < self.hi
            self.eng.select(first_index - self.lo if mine else -1, K)
            t = getattr(self.eng, "timing", None)
            if t:
                tt = t()
                self.last_select_ms, self.last_update_ms = tt["select_ms"], tt["update_ms"]
            return
        cuda = torch is not None and torch.cuda.is_available() and self._exchange()[0].is_cuda
        if cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        mine = self.lo <= first_index < self.hi
        self.eng.select_begin(first_index - self.lo if mine else -1)
        self._gather()
        prof = cuda and self.profile
        evs = []
        for _ in range(K - 1):
            self.eng.step_append()
            if prof:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
            self.eng.step_update()
            if prof:
                b.record()
                evs.append((a, b))
            self._gather()
        self.eng.select_end()
        if cuda:
            e1.record()
            e1.synchronize()
            self.last_select_ms = e0.elapsed_time(e1)
            self.last_update_ms = float(sum(a.elapsed_time(b) for a, b in evs))
