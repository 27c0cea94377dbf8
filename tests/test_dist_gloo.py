"""N > 1 host logic on CPU: world_size 2 over gloo.  The sharding / exchange / termination logic
of gpis_b200.dist.ShardedSelector is driven with a stand-in engine whose arithmetic is the oracle
(tests may use it; the product path never does).  Each rank owns a candidate shard, the factor is
replicated, one fixed-size record per rank is all-gathered per step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gpis_oracle as O
from gpis_b200 import dist as gdist

HDR = 8


class OracleShardEngine:
    """Same call protocol as GpisEngine's step API, arithmetic from the oracle."""

    def __init__(self, pts, tsdf, lo, hi, hyp, train, rank, world):
        self.pts, self.tsdf, self.lo, self.hi = pts, tsdf, lo, hi
        self.rank, self.world = rank, world
        self.h, self.s, self.eps = float(train["levelSet"]), np.sqrt(float(train["beta"])), float(train["eps"])
        self.noise = float(np.exp(2.0 * hyp["lik"]))
        self.gp = O.IncrementalGP(pts[lo:hi], hyp, int(train["activeSetSize"]), False)
        n = hi - lo
        self.active = np.zeros(n, bool); self.high = np.zeros(n, bool); self.low = np.zeros(n, bool)
        self.send = torch.zeros(HDR, dtype=torch.float64)
        self.recv = torch.zeros(HDR * world, dtype=torch.float64)
        self.order, self.done = [], False

    def _record(self, score, local):
        r = self.send
        r.zero_()
        r[0], r[2] = score, float(local)
        if local >= 0:
            g = self.lo + local
            r[1], r[3] = float(g), self.tsdf[g]
            r[4:7] = torch.from_numpy(self.pts[g])

    def _winner(self):
        rec = self.recv.view(self.world, HDR).numpy()
        best = None
        for w in range(self.world):
            if rec[w, 2] < 0:
                continue
            if best is None or rec[w, 0] > rec[best, 0] or (rec[w, 0] == rec[best, 0] and rec[w, 1] < rec[best, 1]):
                best = w
        return best, rec

    def select_begin(self, first_local):
        self._record(np.inf if first_local >= 0 else -np.inf, first_local)

    def step_append(self):
        if self.done:
            return
        w, rec = self._winner()
        if w is None:
            self.done = True
            return
        if w == self.rank:
            self.active[int(rec[w, 2])] = True
        self.order.append(int(rec[w, 1]))
        self.gp.append(rec[w, 4:7].copy(), rec[w, 3], None, self.noise)

    def step_update(self):
        if self.done:
            return
        unc = ~self.active & ~self.high & ~self.low
        if not unc.any():
            self._record(-np.inf, -1)
            return
        idx = np.flatnonzero(unc)
        mu, var = self.gp.mu[idx], self.gp.var[idx]
        hi, lo = mu + self.s * var, mu - self.s * var
        amb = np.minimum(hi - self.h, self.h - lo)
        b = int(np.flatnonzero(amb == amb.max())[0])
        self._record(amb[b], int(idx[b]))
        self.high[idx[lo + self.eps > self.h]] = True
        self.low[idx[hi - self.eps < self.h]] = True

    def select_end(self):
        if self.done:
            return
        w, rec = self._winner()
        if w is not None:
            if w == self.rank:
                self.active[int(rec[w, 2])] = True
            self.order.append(int(rec[w, 1]))


def _worker(rank, world, port, K, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gpis_b200.synth import sphere_tsdf
        pts, tsdf = sphere_tsdf(12, trunc=3.0)
        hyp = {"cov": np.array([np.log(2.5), 0.0]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.05)}
        tr = {"activeSetSize": K, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
        lo, hi = gdist.shard_range(pts.shape[0], rank, world)
        eng = OracleShardEngine(pts, tsdf, lo, hi, hyp, tr, rank, world)
        sel = gdist.ShardedSelector(eng, rank, world, lo, hi, exchange=lambda e, w: (e.send, e.recv))
        sel.select(777, K)
        q.put((rank, eng.order, eng.gp.m, eng.high.copy(), eng.low.copy(), lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("K", [40, 300])
def test_two_rank_sharded_selection_matches_single_process(K):
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from gpis_b200.synth import sphere_tsdf
    pts, tsdf = sphere_tsdf(12, trunc=3.0)
    hyp = {"cov": np.array([np.log(2.5), 0.0]), "mean": np.zeros(4), "lik": 0.5 * np.log(0.05)}
    tr = {"activeSetSize": K, "beta": 10.0, "eps": 1e-2, "levelSet": 0.0}
    ref = O.select_straddle_incremental({"points": pts, "tsdf": tsdf, "normals": None, "noise": None}, tr, 777,
                                        hyp=hyp, prune=False)
    assert res[0][1] == res[1][1] == list(ref["order"])          # every rank agrees with the oracle
    assert res[0][2] == res[1][2] == ref["n_in_model"]           # K limit (40) and early stop (300)
    high = np.concatenate([r[3] for r in res]); low = np.concatenate([r[4] for r in res])
    assert np.array_equal(high, ref["high"]) and np.array_equal(low, ref["low"])


def test_shard_ranges_cover_and_align():
    for n in [1, 63, 64, 65, 1000, 2097152]:
        for w in [1, 2, 3, 8]:
            rs = [gdist.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            assert all(lo % 64 == 0 or lo == n for lo, _ in rs)
    for nz in [1, 9, 16, 128]:
        for w in [1, 2, 3, 8]:
            rs = [gdist.slab_range(nz, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == nz and all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
    assert gdist.owner_of(560310, 2097152, 8) == (2, 36022)


def test_balanced_slabs_partition_the_lattice():
    """slab_range_balanced: contiguous, complete, non-empty slabs; end slabs are thicker than middle ones when the
    reach of a model row (9.85 ell) is a fraction of the axis; equal split when it covers the whole axis."""
    for nz, world, ell in [(128, 8, 4.0), (128, 4, 4.0), (128, 2, 4.0), (10, 8, 4.0), (64, 3, 1.0), (8, 8, 2.0), (256, 8, 6.0)]:
        rs = [gdist.slab_range_balanced(nz, r, world, ell) for r in range(world)]
        assert rs[0][0] == 0 and rs[-1][1] == nz
        assert all(a[1] == b[0] for a, b in zip(rs, rs[1:])) and all(z1 > z0 for z0, z1 in rs)
    rs = [gdist.slab_range_balanced(128, r, 8, 4.0) for r in range(8)]
    sizes = [z1 - z0 for z0, z1 in rs]
    assert sizes[0] > sizes[3] and sizes[-1] > sizes[4] and max(sizes) - min(sizes) <= 8
    assert [gdist.slab_range_balanced(32, r, 4, 100.0) for r in range(4)] == [gdist.slab_range(32, r, 4) for r in range(4)]
    # cube-aligned cuts (the persistent kernel's 16 x 16 x 16 tiles): multiples of 16 when every rank can have 16 slices
    for nz, world in ((128, 2), (128, 4), (128, 8), (100, 4), (96, 3)):
        rs = [gdist.slab_range_balanced(nz, r, world, 4.0, align=16) for r in range(world)]
        assert rs[0][0] == 0 and rs[-1][1] == nz and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert all(z1 - z0 >= 16 for z0, z1 in rs) and all(z0 % 16 == 0 for z0, _ in rs)
    assert [gdist.slab_range_balanced(128, r, 8, 4.0, align=16) for r in range(8)] == [gdist.slab_range(128, r, 8) for r in range(8)]
    # too few slices for that: the unaligned cuts stay
    assert [gdist.slab_range_balanced(40, r, 4, 4.0, align=16) for r in range(4)] == [gdist.slab_range_balanced(40, r, 4, 4.0) for r in range(4)]
