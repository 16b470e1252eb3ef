"""Batch sharding over the ranks of a torch.distributed job.

The path shards by independent units: image i of a (B, Nr, Nc) batch is transformed on its own, so rank r of G owns
the contiguous block [r*B/G, (r+1)*B/G) as ONE batched `Wavelets` object and the transform kernels never communicate.
The only collectives move inputs and outputs: `scatter_batch` (root -> owners), `gather_*` (owners -> root) and one
all-gather of per-image scalars for the norms.

On GPUs (the product path: no `engine` argument) this class is a BINDING over Layer C of the C ABI
(`pdwt_shard_*`, csrc/pdwt_sharded.cu): the library owns an NCCL communicator, inputs and outputs move DEVICE TO
DEVICE (grouped ncclSend/ncclRecv between the root's device buffer and the owners' `d_image` / sub-band buffers,
ncclAllGather for the norms) -- nothing bounces through the host; torch.distributed is only used to hand the 128-byte
NCCL id to the other ranks.  With an injected `engine` (the CPU tests run the oracle under gloo, because the product
engine exists only on a CUDA device) the same partitioning runs with torch.distributed point-to-point calls.
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
        self.native = engine is None
        if self.native:
            self._init_native(img, wname, levels, root, local, kw)
            return
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

    # ---- product path: Layer C of the C ABI over NCCL, device to device ---------------------------------------
    def _init_native(self, img, wname, levels, root, local, kw):
        import ctypes as C
        import torch
        from . import PdwtError, Wavelets, lib
        dist, group = self.dist, self.group
        L = self._L = lib()
        vp, ll, sz, ci = C.c_void_p, C.c_longlong, C.c_size_t, C.c_int
        L.pdwt_shard_last_error.restype = C.c_char_p
        L.pdwt_shard_unique_id.argtypes = [C.c_char_p]
        L.pdwt_shard_create.argtypes = [C.POINTER(vp), C.c_char_p, ci, ci]
        L.pdwt_shard_destroy.argtypes = [vp]
        L.pdwt_shard_destroy.restype = None
        L.pdwt_shard_scatter_image.argtypes = [vp, vp, vp, ll, sz, ci]
        L.pdwt_shard_gather_image.argtypes = [vp, vp, vp, ll, sz, ci]
        L.pdwt_shard_gather_coeff.argtypes = [vp, vp, ci, vp, ll, sz, ci]
        L.pdwt_shard_norms.argtypes = [vp, vp, ci, C.POINTER(C.c_float), ll]
        self._C = C
        if L.pdwt_device_count() < 1:
            raise PdwtError("ShardedWavelets: no CUDA device -- pdwt_b200 has no CPU fallback")
        # the NCCL id travels out of band (any channel would do); everything else is device to device
        ident = [None]
        if self.rank == root:
            buf = C.create_string_buffer(128)
            if L.pdwt_shard_unique_id(buf) != 0:
                raise PdwtError("pdwt_shard_unique_id: " + L.pdwt_shard_last_error().decode())
            ident = [buf.raw]
        dist.broadcast_object_list(ident, src=root, group=group)
        h = vp()
        if L.pdwt_shard_create(C.byref(h), ident[0], self.world, self.rank) != 0:
            raise PdwtError("pdwt_shard_create: " + L.pdwt_shard_last_error().decode())
        self._shard = h
        # geometry: the batch lives on root (numpy, or a CUDA tensor = no host copy at all), or block by block (`local`)
        a = None
        if self.rank == root or local:
            a = img if isinstance(img, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(img, np.float32))
            if a.ndim != 3:
                raise ValueError("ShardedWavelets wants a (B, Nr, Nc) batch")
            a = a.to(device="cuda", dtype=torch.float32).contiguous()
        if local:
            counts = [None] * self.world
            dist.all_gather_object(counts, int(a.shape[0]), group=group)
            self.B, self.Nr, self.Nc = sum(counts), int(a.shape[1]), int(a.shape[2])
            if [e - s_ for s_, e in partition(self.B, self.world)] != counts:
                raise ValueError("local=True wants the blocks of partition(B, world)")
        else:
            meta = [tuple(a.shape) if self.rank == root else None]
            dist.broadcast_object_list(meta, src=root, group=group)
            self.B, self.Nr, self.Nc = meta[0]
        self.blocks = partition(self.B, self.world)
        self.lo, self.hi = self.blocks[self.rank]
        n = self.hi - self.lo
        if local:
            self.W = Wavelets(a, wname, levels, **kw) if n else None
        else:
            self.W = Wavelets(None, wname, levels, shape=(n, self.Nr, self.Nc), **kw) if n else None
            self.scatter_batch(a)

    def _h(self):
        return self.W._h if self.W is not None else None

    def _rc(self, rc, what):
        from . import PdwtError
        if rc != 0:
            raise PdwtError(f"{what} failed ({rc}): {self._L.pdwt_shard_last_error().decode()} / "
                            f"{self._L.pdwt_last_cuda_error_string().decode()}")

    def close(self):
        if getattr(self, "native", False) and getattr(self, "_shard", None):
            self._L.pdwt_shard_destroy(self._shard)
            self._shard = None

    def _native_gather(self, num, plane_shape, device, out=None):
        """sub-band `num` (None: the image) of every plane -> (B, *plane_shape) on root: a CUDA tensor if `device`
        (`out`: a contiguous float32 CUDA tensor of that shape to gather into)"""
        import torch
        C = self._C
        plane = int(np.prod(plane_shape))
        full = None
        if self.rank == self.root:
            full = out if out is not None else torch.empty((self.B, *plane_shape), dtype=torch.float32, device="cuda")
            if tuple(full.shape) != (self.B, *plane_shape) or not full.is_contiguous() or not full.is_cuda:
                raise ValueError("gather: `out` must be a contiguous float32 CUDA tensor of shape (B, *plane)")
        ptr = C.c_void_p(full.data_ptr()) if full is not None else None
        if num is None:
            rc = self._L.pdwt_shard_gather_image(self._shard, self._h(), ptr, self.B, plane, self.root)
        else:
            rc = self._L.pdwt_shard_gather_coeff(self._shard, self._h(), num, ptr, self.B, plane, self.root)
        self._rc(rc, "pdwt_shard_gather")
        if self.W is not None:
            self.W.sync()
        if full is None:
            return None
        torch.cuda.current_stream().synchronize()
        return full if device else full.cpu().numpy()

    # ---- collectives (inputs / outputs only) --------------------------------------------------------------
    def _device(self):
        import torch
        return torch.device("cuda", torch.cuda.current_device()) if self.dist.get_backend(self.group) == "nccl" \
            else torch.device("cpu")

    def scatter_batch(self, full):
        """root's (B, Nr, Nc) array -> this rank's block, point to point (root's NVLink egress is the only cost)"""
        import torch
        if self.native:   # device to device, into the owners' d_image (pdwt_shard_scatter_image)
            if self.rank == self.root:
                full = full if isinstance(full, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(full, np.float32))
                full = full.to(device="cuda", dtype=torch.float32).contiguous()
                torch.cuda.current_stream().synchronize()
            ptr = self._C.c_void_p(full.data_ptr()) if self.rank == self.root else None
            self._rc(self._L.pdwt_shard_scatter_image(self._shard, self._h(), ptr, self.B, self.Nr * self.Nc, self.root),
                     "pdwt_shard_scatter_image")
            if self.W is not None:
                self.W.sync()     # `full` may be released by the caller
            return None
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
        if self.native:
            out = np.empty(self.B, np.float32)
            self._rc(self._L.pdwt_shard_norms(self._shard, self._h(), 1 if name == "norm1" else 2,
                                              out.ctypes.data_as(self._C.POINTER(self._C.c_float)), self.B),
                     "pdwt_shard_norms")
            return out
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

    def gather_image(self, device=False, out=None):
        """(B, Nr, Nc) on root (numpy; `device=True`: the CUDA tensor it was gathered into), None elsewhere"""
        if self.native:
            return self._native_gather(None, (self.Nr, self.Nc), device or out is not None, out)
        local = self._stack_local(self.W.get_image()) if self.W is not None else None
        return self._gather(local, (self.Nr, self.Nc))

    def gather_coeff(self, num, device=False):
        if self.native:
            shapes = [None] * self.world   # a rank that owns nothing has no object to ask
            self.dist.all_gather_object(shapes, self.W.coeff_shape(num) if self.W is not None else None, group=self.group)
            shp = next(s_ for s_ in shapes if s_ is not None)
            return self._native_gather(num, tuple(int(v) for v in shp), device)
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
