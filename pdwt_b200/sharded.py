"""Batch sharding over the ranks of a torch.distributed job.

The path shards by independent units: image i of a (B, Nr, Nc) batch is transformed on its own, so rank r of G owns
the contiguous block [r*B/G, (r+1)*B/G) as ONE batched `Wavelets` object and the transform kernels never communicate.
The only collectives move inputs and outputs: `scatter_batch` (root -> owners), `gather_*` (owners -> root) and one
all-gather of per-image scalars for the norms.  One process per GPU, NCCL over NVLink on GPUs; the same host logic runs
under gloo in the CPU tests, with the transform engine injected (`engine=`), because the product engine
(`pdwt_b200.Wavelets`) exists only on a CUDA device.
"""
from __future__ import annotations

import numpy as np

__all__ = ["partition", "ShardedWavelets"]


def partition(n_items: int, world: int):
    """contiguous blocks, the first `n_items % world` ranks get one more: [(start, stop)] * world"""
    base, extra = divmod(n_items, world)
    out, s = [], 0
    for r in range(world):
        e = s + base + (1 if r < extra else 0)
        out.append((s, e))
        s = e
    return out


class ShardedWavelets:
    """A batch of B images spread over the ranks of `group`; every rank holds a `Wavelets`-like engine over its block.

    img     : (B, Nr, Nc) float32 array on `root` (ignored elsewhere) or, with `local=True`, this rank's own block.
    engine  : callable(local_block, wname, levels, **kw) -> object with forward/inverse/soft_threshold/
              hard_threshold/norm1/norm2sq/get_image/get_coeff/ncoeffs; default = pdwt_b200.Wavelets (CUDA).
    """

    def __init__(self, img, wname, levels, batch=None, shape=None, root=0, local=False, group=None, engine=None, **kw):
        import torch.distributed as dist
        self.dist, self.group, self.root = dist, group, root
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if engine is None:
            from . import Wavelets as engine  # the CUDA product path; raises without a device (no CPU fallback)
        meta = [None]
        if self.rank == root or local:
            a = np.asarray(img, dtype=np.float32)
            if a.ndim != 3:
                raise ValueError("ShardedWavelets wants a (B, Nr, Nc) batch")
            meta = [(a.shape[0] if not local else None, a.shape[1], a.shape[2])]
        if not local:
            dist.broadcast_object_list(meta, src=root, group=group)
            self.B, self.Nr, self.Nc = meta[0]
            self.blocks = partition(self.B, self.world)
            mine = self.scatter_batch(a if self.rank == root else None)
        else:
            mine = a
            counts = [None] * self.world
            dist.all_gather_object(counts, mine.shape[0], group=group)
            self.Nr, self.Nc = mine.shape[1], mine.shape[2]
            self.B = sum(counts)
            self.blocks, s = [], 0
            for c in counts:
                self.blocks.append((s, s + c))
                s += c
        self.lo, self.hi = self.blocks[self.rank]
        self.W = engine(mine, wname, levels, **kw) if self.hi > self.lo else None

    # ---- collectives (inputs / outputs only) --------------------------------------------------------------
    def _device(self):
        import torch
        return torch.device("cuda", torch.cuda.current_device()) if self.dist.get_backend(self.group) == "nccl" \
            else torch.device("cpu")

    def scatter_batch(self, full):
        """root's (B, Nr, Nc) array -> this rank's block, point to point (root's NVLink egress is the only cost)"""
        import torch
        dev = self._device()
        lo, hi = self.blocks[self.rank]
        mine = torch.empty((hi - lo, self.Nr, self.Nc), dtype=torch.float32, device=dev)
        if self.rank == self.root:
            t = torch.from_numpy(np.ascontiguousarray(full, np.float32))
            reqs = []
            for r, (s, e) in enumerate(self.blocks):
                if e == s:
                    continue
                if r == self.root:
                    mine.copy_(t[s:e])
                else:
                    reqs.append(self.dist.isend(t[s:e].contiguous().to(dev), dst=r, group=self.group))
            for q in reqs:
                q.wait()
        elif hi > lo:
            self.dist.recv(mine, src=self.root, group=self.group)
        return mine if dev.type == "cuda" else mine.numpy()

    def _gather(self, local, plane_shape):
        """per-rank (n_r, *plane_shape) arrays -> (B, *plane_shape) on root, None elsewhere"""
        import torch
        dev = self._device()
        mine = torch.from_numpy(np.ascontiguousarray(local, np.float32)).to(dev) if local is not None else None
        if self.rank != self.root:
            if mine is not None and mine.shape[0]:
                self.dist.send(mine, dst=self.root, group=self.group)
            return None
        out = np.empty((self.B, *plane_shape), dtype=np.float32)
        for r, (s, e) in enumerate(self.blocks):
            if e == s:
                continue
            if r == self.root:
                out[s:e] = mine.cpu().numpy()
            else:
                buf = torch.empty((e - s, *plane_shape), dtype=torch.float32, device=dev)
                self.dist.recv(buf, src=r, group=self.group)
                out[s:e] = buf.cpu().numpy()
        return out

    # ---- the Wavelets methods, applied to every block ----------------------------------------------------
    def forward(self):
        if self.W is not None:
            self.W.forward()

    def inverse(self):
        if self.W is not None:
            self.W.inverse()

    def soft_threshold(self, beta, do_thresh_appcoeffs=0, normalize=0):
        if self.W is not None:
            self.W.soft_threshold(beta, do_thresh_appcoeffs, normalize)

    def hard_threshold(self, beta, do_thresh_appcoeffs=0, normalize=0):
        if self.W is not None:
            self.W.hard_threshold(beta, do_thresh_appcoeffs, normalize)

    def _scalars(self, name):
        """one value per image, on every rank: all-gather of the local norms"""
        import torch
        local = np.atleast_1d(np.asarray(getattr(self.W, name)(), np.float32)) if self.W is not None \
            else np.zeros(0, np.float32)
        outs = [None] * self.world
        self.dist.all_gather_object(outs, local.tolist(), group=self.group)
        return np.array([v for o in outs for v in o], dtype=np.float32)

    def norm1(self):
        return self._scalars("norm1")

    def norm2sq(self):
        return self._scalars("norm2sq")

    def _stack_local(self, a):
        a = np.asarray(a, np.float32)
        return a[None] if a.ndim == 2 else a

    def gather_image(self):
        local = self._stack_local(self.W.get_image()) if self.W is not None else None
        return self._gather(local, (self.Nr, self.Nc))

    def gather_coeff(self, num):
        if self.W is not None:
            local = self._stack_local(self.W.get_coeff(num))
            shp = list(local.shape[1:])
        else:
            local, shp = None, None
        shapes = [None] * self.world
        self.dist.all_gather_object(shapes, shp, group=self.group)
        shp = next(s for s in shapes if s is not None)
        return self._gather(local, tuple(shp))

    @property
    def ncoeffs(self):
        n = [None] * self.world
        self.dist.all_gather_object(n, self.W.ncoeffs if self.W is not None else None, group=self.group)
        return next(v for v in n if v is not None)
