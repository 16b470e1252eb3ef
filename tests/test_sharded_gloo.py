"""CPU, world_size 2 (and 3) over gloo: the host logic of the multi-GPU path (pdwt_b200/sharded.py) -- partition,
scatter, per-block transform, gather, norm all-gather.  The transform engine is injected: here the CPU oracle stands in
for the CUDA `Wavelets` (which cannot run without a device); the sharding code is the product code."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from pdwt_b200.sharded import partition


def test_partition_blocks():
    assert partition(512, 8) == [(64 * r, 64 * (r + 1)) for r in range(8)]          # BASELINE configs[4]
    assert partition(5, 2) == [(0, 3), (3, 5)]
    assert partition(2, 3) == [(0, 1), (1, 2), (2, 2)]                               # a rank may own nothing
    for n, w in ((7, 3), (1, 4), (100, 7)):
        b = partition(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))


class OracleBatch:
    """batched engine with the Wavelets method set, one oracle object per image"""

    def __init__(self, block, wname, levels, **kw):
        import oracle
        self.objs = [oracle.Wavelets(x, wname, levels, **kw) for x in np.asarray(block)]

    def forward(self): [o.forward() for o in self.objs]
    def inverse(self): [o.inverse() for o in self.objs]
    def soft_threshold(self, *a): [o.soft_threshold(*a) for o in self.objs]
    def hard_threshold(self, *a): [o.hard_threshold(*a) for o in self.objs]
    def norm1(self): return np.array([o.norm1() for o in self.objs], np.float32)
    def norm2sq(self): return np.array([o.norm2sq() for o in self.objs], np.float32)
    def get_image(self): return np.stack([o.get_image() for o in self.objs])
    def get_coeff(self, n): return np.stack([o.get_coeff(n) for o in self.objs])

    @property
    def ncoeffs(self): return self.objs[0].ncoeffs


def _worker(rank, world, port, B, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pdwt_b200.sharded import ShardedWavelets
        full = np.stack([(np.random.default_rng(i).standard_normal((48, 64)) * 50 + 128).astype(np.float32)
                         for i in range(B)])
        S = ShardedWavelets(full if rank == 0 else None, "db3", 2, engine=OracleBatch)
        assert (S.lo, S.hi) == partition(B, world)[rank]
        S.forward()
        n1 = S.norm1()
        c1 = S.gather_coeff(1)
        S.soft_threshold(10.0)
        n1t = S.norm1()
        S.inverse()
        rec = S.gather_image()
        nco = S.ncoeffs          # collective: every rank calls it
        if rank == 0:
            q.put((n1, c1, n1t, rec, nco))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,B", [(2, 5), (3, 2)])
def test_sharded_batch_equals_one_by_one(world, B):
    import oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    [p.start() for p in procs]
    n1, c1, n1t, rec, ncoeffs = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert ncoeffs == 7 and n1.shape == (B,) and c1.shape[0] == B and rec.shape == (B, 48, 64)
    for i in range(B):
        x = (np.random.default_rng(i).standard_normal((48, 64)) * 50 + 128).astype(np.float32)
        O = oracle.Wavelets(x, "db3", 2)
        O.forward()
        assert n1[i] == np.float32(O.norm1())
        assert np.array_equal(c1[i], O.get_coeff(1))
        O.soft_threshold(10.0)
        assert n1t[i] == np.float32(O.norm1())
        O.inverse()
        assert np.array_equal(rec[i], O.get_image())
