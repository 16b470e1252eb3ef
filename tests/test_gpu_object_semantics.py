"""GPU (-m gpu): object-level behaviour around streams, caches and refusals (reference wt.cu:84-233, 370-418)."""
import numpy as np
import pytest
from conftest import bitexact

import oracle
import pdwt_b200
from pdwt_b200 import Wavelets

pytestmark = pytest.mark.gpu


def rnd(shape, seed=0):
    return (np.random.default_rng(seed).standard_normal(shape) * 50 + 128).astype(np.float32)


def test_construct_then_forward_at_once_on_a_nonblocking_stream():
    """the constructor's fills run on the legacy stream, which does not order against cudaStreamNonBlocking streams: they
    must have landed before the first transform on such a stream (large batch = long fills)"""
    import torch
    x = torch.randn((48, 1024, 1024), device="cuda") * 50 + 128
    ref = None
    for _ in range(3):
        s = torch.cuda.Stream()   # torch streams are non-blocking
        W = Wavelets(x, "db7", 3)
        W.set_stream(s)
        W.forward()
        c = [W.get_coeff(i) for i in range(W.ncoeffs)]
        if ref is None:
            O = oracle.Wavelets(x[5].cpu().numpy(), "db7", 3)
            O.forward()
            ref = [O.get_coeff(i) for i in range(W.ncoeffs)]
        for i in range(W.ncoeffs):
            assert bitexact(c[i][5], ref[i]), i
        W2 = W.copy()             # deep copy while the source's stream may still be busy
        W.inverse()
        W2.inverse()
        assert bitexact(W.get_image(), W2.get_image())


def test_norm_cache_survives_a_stream_switch_and_can_be_invalidated():
    import torch
    x = rnd((512, 768), 3)
    W, O = Wavelets(x, "db7", 3), oracle.Wavelets(x, "db7", 3)
    W.set_stream(torch.cuda.Stream())
    W.forward(); O.forward()
    W.soft_threshold(9.0); O.soft_threshold(9.0)
    W.set_stream(torch.cuda.Stream())          # the sums were published in the old stream's order
    assert abs(W.norm1() - O.norm1()) <= 1e-5 * O.norm1()
    # a write behind the object's back is invisible to the cache until it is told
    W.soft_threshold(1.0); O.soft_threshold(1.0)
    n = W.norm1()
    import ctypes as C
    L = pdwt_b200.lib()
    ptrs = (C.c_void_p * W.ncoeffs)(*[W.coeff_int_ptr(i) for i in range(W.ncoeffs)])
    torch.cuda.synchronize()
    assert L.pdwt_call_soft_thresh(ptrs, 1e9, W.info, 0, 0, 1, None) == 0   # Layer A on the object's buffers: details -> 0
    torch.cuda.synchronize()
    assert W.norm1() == n                      # stale by design
    W.invalidate_norm_cache()
    O.soft_threshold(1e9)
    assert abs(W.norm1() - O.norm1()) <= 1e-5 * O.norm1()


def test_cycle_spinning_is_refused_in_1d_like_the_reference():
    """wt.cu:175-179: W_CREATION_ERROR, forward() is then a no-op"""
    W = Wavelets(rnd((4, 512), 1), "db4", 2, do_cycle_spinning=1, ndim=1)
    assert W.state == pdwt_b200.W_CREATION_ERROR
    W.forward()
    assert W.state == pdwt_b200.W_CREATION_ERROR
