"""CPU: the C-ABI library loads, exports every symbol include/pdwt_b200.h declares, and its host-only entry
points (filters, size helpers) agree with the oracle.  No compute call needs a device here; on a box without one
the compute entry points must FAIL LOUDLY (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
from conftest import ROOT

import oracle
import pdwt_b200


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "pdwt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(pdwt_[a-z0-9_]+)\s*\(", hdr))
    return sorted(n for n in names if n not in ("pdwt_status",))


def test_library_exports_every_declared_symbol():
    L = pdwt_b200.lib()
    syms = declared_symbols()
    assert len(syms) > 60
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_filter_tables_match_oracle():
    names = pdwt_b200.wavelet_names()
    assert len(names) == 72 and names[0] == "db2" and names[-1] == "haar"
    for n in names:
        for swt in (0, 1):
            a = pdwt_b200.filters(n, swt)
            b = oracle.filters(n, swt)
            assert a[0] == b[0]
            for x, y in zip(a[1:], b[1:]):
                assert np.array_equal(x, y)
    assert pdwt_b200.filters("DB7")[0] == 14                      # strcasecmp, separable.cu:33
    for alias in ("haar", "db1", "bior1.1", "rbior1.1"):           # separable.cu:24-28
        assert pdwt_b200.filters(alias, 0)[0] == 2
    with pytest.raises(KeyError):
        pdwt_b200.filters("nosuch")
    with pytest.raises(KeyError):
        pdwt_b200.filters("db1", 1)                                # only the Haar kernels know that alias


def test_size_helpers_follow_reference_rules():
    L = pdwt_b200.lib()
    assert [L.pdwt_div2(n) for n in (1, 2, 3, 4096, 4097)] == [1, 1, 2, 2048, 2049]   # utils.cu:24-27
    assert L.pdwt_max_level(4096, 4096, 2, 14) == 8                # ilog2(4096/13)
    assert L.pdwt_max_level(40, 40, 2, 14) == 1
    assert L.pdwt_max_level(1, 4096, 1, 2) == 12
    w = pdwt_b200.WInfo(2, 57, 64, 3, 0, 4)
    assert L.pdwt_num_coeffs(w) == 10
    nr, nc = C.c_int(), C.c_int()
    L.pdwt_coeff_dims(w, 0, C.byref(nr), C.byref(nc))
    assert (nr.value, nc.value) == (8, 8)                          # 57 -> 29 -> 15 -> 8
    L.pdwt_coeff_dims(w, 4, C.byref(nr), C.byref(nc))
    assert (nr.value, nc.value) == (15, 16)
    assert L.pdwt_coeff_alloc_elems(w, 0) == 29 * 32               # level-1 sized, common.cu:402-406
    w1 = pdwt_b200.WInfo(1, 3, 1001, 4, 0, 2)
    assert L.pdwt_num_coeffs(w1) == 5
    assert L.pdwt_coeff_alloc_elems(w1, 0) == 3 * 501
    ws = pdwt_b200.WInfo(2, 33, 47, 2, 1, 6)
    assert L.pdwt_coeff_alloc_elems(ws, 0) == 33 * 47 and L.pdwt_coeff_alloc_elems(ws, 5) == 33 * 47
    assert L.pdwt_coeff_dims(w, 10, C.byref(nr), C.byref(nc)) < 0


def test_custom_filter_validation():
    L = pdwt_b200.lib()
    h = C.c_void_p()
    z = (C.c_float * 64)()
    assert L.pdwt_filters_create_custom(C.byref(h), 41, z, z, z, z) == pdwt_b200.PDWT_ERR_FILTER_LEN
    assert L.pdwt_filters_create_custom(C.byref(h), 9, z, z, z, z) == 9
    assert L.pdwt_filters_hlen(h) == 9
    L.pdwt_filters_destroy(h)


def test_no_cpu_fallback_without_device():
    L = pdwt_b200.lib()
    if L.pdwt_device_count() > 0:
        pytest.skip("a device is present")
    with pytest.raises(pdwt_b200.PdwtError):
        pdwt_b200.Wavelets(np.zeros((16, 16), np.float32), "db2", 1)
    # Layer A directly: a compute entry point reports a CUDA failure instead of computing anything on the host
    h = C.c_void_p()
    assert L.pdwt_filters_create(C.byref(h), b"db2", 0) == 4
    w = pdwt_b200.WInfo(2, 16, 16, 1, 0, 4)
    fake = C.c_void_p(0x1000)
    ptrs = (C.c_void_p * 4)(0x1000, 0x1000, 0x1000, 0x1000)
    rc = L.pdwt_forward_separable(h, fake, ptrs, fake, w, 1, None)
    assert rc == pdwt_b200.PDWT_ERR_CUDA
    L.pdwt_filters_destroy(h)


def test_filter_table_regenerates_from_the_reference_sources(tmp_path):
    """the committed table is exactly what tools/gen_filter_bank.py makes of the reference's filters.cpp (build container
    only: the GPU box has no /root/reference; there the live reference library checks the values, test_gpu_reference_fullsize)"""
    if not os.path.exists("/root/reference/src/filters.cpp"):
        pytest.skip("/root/reference is not available here")
    import subprocess
    import sys
    out = tmp_path / "filter_bank.inc"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_filter_bank.py"), str(out)], cwd=ROOT,
                          stdout=subprocess.DEVNULL)
    assert out.read_bytes() == open(os.path.join(ROOT, "pdwt_b200", "csrc", "filter_bank.inc"), "rb").read()


def test_shard_block_arithmetic_matches_the_python_partition():
    """Layer C: contiguous blocks, the first n % world ranks own one plane more (pure arithmetic, no NCCL, no device)"""
    from pdwt_b200.sharded import partition
    L = pdwt_b200.lib()
    L.pdwt_shard_block.argtypes = [C.c_longlong, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    L.pdwt_shard_block.restype = None
    for n, world in ((512, 8), (5, 2), (2, 3), (7, 3), (1, 4), (100, 7), (0, 2)):
        blocks = partition(n, world)
        for r in range(world):
            f, c = C.c_longlong(), C.c_longlong()
            L.pdwt_shard_block(n, world, r, C.byref(f), C.byref(c))
            assert (f.value, f.value + c.value) == blocks[r], (n, world, r)
