"""GPU (-m gpu): the CUDA path against the LIVE reference library at BASELINE.json's own sizes.

`oracle/_ref/libpdwt_ref.so` is the unmodified reference (pierrepaleo/PDWT, src/*.cu) compiled for sm_100 by
`make -C oracle ref` in the build container; it travels to the GPU box with the snapshot.  It is the CHECKER here,
never the thing measured or shipped.  Every sub-band and every reconstruction of the four GPU configurations of
BASELINE.json is compared bit for bit (north_star's bar is 1e-5 relative; the kernels keep the reference's FMA order,
so the buffers are identical), all 72 wavelet names are compared with the reference's own tables
(filters.cpp:5919-6002), and Layer A of the C ABI is driven with caller-owned, reference-style buffers
(3L+1 separate allocations, common.cu:400-445)."""
import ctypes as C
import os

import numpy as np
import pytest
from conftest import ROOT, bitexact, nerr

import pdwt_b200
from pdwt_b200 import Wavelets

pytestmark = pytest.mark.gpu

_REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpdwt_ref.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(_REF_SO):
        pytest.skip("oracle/_ref/libpdwt_ref.so not built (needs /root/reference in the build container)")
    from make_golden import load_ref
    return load_ref()


def rnd(shape, seed=0):
    return (np.random.default_rng(seed).standard_normal(shape) * 50 + 128).astype(np.float32)


def ref_obj(L, x, wname, levels, sep=1, swt=0, ndim=2):
    from make_golden import Ref
    return Ref(L, x, wname, levels, sep, swt, ndim)


def assert_same_coeffs(W, R, what):
    assert W.info.nlevels == R.nlevels and W.info.hlen == R.hlen, what
    for i in range(W.ncoeffs):
        a, b = W.get_coeff(i), R.coeff(i)
        assert nerr(a, b) <= 1e-5, f"{what}: sub-band {i} off by {nerr(a, b)}"
        assert bitexact(a, b), f"{what}: sub-band {i} not bit-exact"


# ---------------------------------------------------------------------------------------- BASELINE.json configs
def test_c2_4096_db7_l3_every_subband(ref):
    """configs[1]: 2-D separable DWT db7, 3 levels, 4096 x 4096 (wt.cu:236-307 -> separable.cu:179-209, 332-364)"""
    x = rnd((4096, 4096), 2)
    W, R = Wavelets(x, "db7", 3), ref_obj(ref, x, "db7", 3)
    W.forward(); ref.ref_forward(R.h)
    assert_same_coeffs(W, R, "C2")
    assert abs(W.norm1() - ref.ref_norm1(R.h)) <= 1e-5 * ref.ref_norm1(R.h)
    W.inverse(); ref.ref_inverse(R.h)
    assert bitexact(W.get_image(), R.image())
    R.close()


@pytest.mark.parametrize("multi", ["0", "1"])
def test_c2_cross_level_launch_matches(ref, multi, monkeypatch):
    """the opt-in cross-level launch (PDWT_MULTI=1: all levels from one work queue) gives the same bits"""
    monkeypatch.setenv("PDWT_MULTI", multi)
    x = rnd((2048, 4096), 3)
    W, R = Wavelets(x, "db7", 4), ref_obj(ref, x, "db7", 4)
    for _ in range(3):   # the queue's counters are cumulative over launches
        W.set_image(x)
        W.forward()
    ref.ref_forward(R.h)
    assert_same_coeffs(W, R, f"multi={multi}")
    W.inverse(); ref.ref_inverse(R.h)
    assert bitexact(W.get_image(), R.image())
    R.close()


def test_c3_2048_sym8_swt_l4(ref):
    """configs[2]: 2-D SWT sym8, 4 levels, 2048 x 2048 (separable.cu:496-515, 629-649)"""
    x = rnd((2048, 2048), 4)
    W, R = Wavelets(x, "sym8", 4, do_swt=1), ref_obj(ref, x, "sym8", 4, swt=1)
    W.forward(); ref.ref_forward(R.h)
    assert_same_coeffs(W, R, "C3")
    W.inverse(); ref.ref_inverse(R.h)
    assert bitexact(W.get_image(), R.image())
    R.close()


def test_c4_4096_nonseparable_soft_norm_roundtrip(ref):
    """configs[3]: non-separable DWT, 2 levels, 4096 x 4096: forward -> norm1 -> soft_threshold -> norm1 -> inverse
    (README.md:90-103; nonseparable.cu:233-291, common.cu:219-249, wt.cu:398-418)"""
    x = rnd((4096, 4096), 5)
    W, R = Wavelets(x, "db7", 2, do_separable=0), ref_obj(ref, x, "db7", 2, sep=0)
    W.forward(); ref.ref_forward(R.h)
    assert_same_coeffs(W, R, "C4 forward")
    n, nr_ = W.norm1(), ref.ref_norm1(R.h)
    assert abs(n - nr_) <= 1e-5 * nr_
    W.soft_threshold(10.0, 0, 0); ref.ref_soft_threshold(R.h, 10.0, 0, 0)
    assert_same_coeffs(W, R, "C4 thresholded")
    n, nr_ = W.norm1(), ref.ref_norm1(R.h)
    assert abs(n - nr_) <= 1e-5 * nr_
    W.inverse(); ref.ref_inverse(R.h)
    assert bitexact(W.get_image(), R.image())
    R.close()


def test_c5_batch_planes_match_single_reference_objects(ref):
    """configs[4] per GPU: 64 images of 2048 x 2048, db7, 3 levels, as ONE batched object; planes 0, 31 and 63 against
    one reference object each (the reference has no batch: it would loop 64 objects, TODO.txt:15)"""
    import torch
    B = 64
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn((B, 2048, 2048), device="cuda", generator=g) * 50 + 128
    W = Wavelets(x, "db7", 3)
    W.forward()
    planes = (0, 31, 63)
    refs = {}
    for p in planes:
        R = ref_obj(ref, x[p].cpu().numpy(), "db7", 3)
        ref.ref_forward(R.h)
        refs[p] = R
    for i in range(W.ncoeffs):
        c = W.get_coeff(i)
        for p in planes:
            assert bitexact(c[p], refs[p].coeff(i)), f"plane {p} sub-band {i}"
    W.inverse()
    img = W.get_image()
    for p in planes:
        ref.ref_inverse(refs[p].h)
        assert bitexact(img[p], refs[p].image()), f"plane {p} reconstruction"
        refs[p].close()


# ---------------------------------------------------------------------------------------------- all 72 banks
def test_all_72_wavelets_against_the_reference_tables(ref):
    """every name of filters.cpp:5919-6002 through the reference's own lookup and constant-memory upload
    (separable.cu:19-54), not through a table shared with the oracle"""
    x = rnd((96, 160), 5)
    names = pdwt_b200.wavelet_names()
    assert len(names) == 72
    for wname in names:
        W, R = Wavelets(x, wname, 2), ref_obj(ref, x, wname, 2)
        W.forward(); ref.ref_forward(R.h)
        assert_same_coeffs(W, R, wname)
        W.inverse(); ref.ref_inverse(R.h)
        assert bitexact(W.get_image(), R.image()), wname
        R.close()


# ------------------------------------------------------------------------ Layer A with reference-style buffers
def _dev(n):
    import torch
    return torch.zeros(int(n), dtype=torch.float32, device="cuda")


@pytest.mark.parametrize("wname,levels,swt,shape", [("db7", 3, 0, (512, 768)), ("sym8", 2, 1, (256, 320)),
                                                    ("haar", 3, 0, (300, 200))])
def test_layer_a_with_caller_owned_buffers(ref, wname, levels, swt, shape):
    """INTEGRATION.md route 2: the 16 drivers + thresholds + norms called the way wt.cu would call them, on buffers the
    CALLER allocated one by one (w_create_coeffs_buffer, common.cu:400-445)"""
    import torch
    L = pdwt_b200.lib()
    x = rnd(shape, 9)
    f = C.c_void_p()
    hlen = L.pdwt_filters_create(C.byref(f), wname.encode(), swt)
    assert hlen > 0
    w = pdwt_b200.WInfo(2, shape[0], shape[1], levels, swt, hlen)
    nco = L.pdwt_num_coeffs(w)
    d_image = torch.from_numpy(x).cuda()
    bufs = [_dev(L.pdwt_coeff_alloc_elems(w, i)) for i in range(nco)]   # 3L+1 separate allocations
    d_tmp = _dev(2 * x.size)
    ptrs = (C.c_void_p * nco)(*[b.data_ptr() for b in bufs])
    stream = torch.cuda.Stream()
    sp = C.c_void_p(stream.cuda_stream)
    torch.cuda.synchronize()
    args = (f, C.c_void_p(d_image.data_ptr()), ptrs, C.c_void_p(d_tmp.data_ptr()), w, 1, sp)
    assert L.pdwt_forward(*args, 1) == 0
    R = ref_obj(ref, x, wname, levels, swt=swt)
    ref.ref_forward(R.h)
    stream.synchronize()

    def sub(i):
        nr, nc = C.c_int(), C.c_int()
        assert L.pdwt_coeff_dims(w, i, C.byref(nr), C.byref(nc)) == 0
        return bufs[i][: nr.value * nc.value].cpu().numpy().reshape(nr.value, nc.value)

    for i in range(nco):
        assert bitexact(sub(i), R.coeff(i)), f"{wname} sub-band {i}"
    out = (C.c_float * 1)()
    assert L.pdwt_norm1(ptrs, w, 1, out, sp) == 0
    assert abs(out[0] - ref.ref_norm1(R.h)) <= 1e-5 * ref.ref_norm1(R.h)
    assert L.pdwt_norm2sq(ptrs, w, 1, out, sp) == 0
    assert abs(out[0] - ref.ref_norm2sq(R.h)) <= 1e-5 * ref.ref_norm2sq(R.h)
    assert L.pdwt_call_soft_thresh(ptrs, 7.5, w, 1, 1, 1, sp) == 0
    ref.ref_soft_threshold(R.h, 7.5, 1, 1)
    stream.synchronize()
    for i in range(nco):
        assert bitexact(sub(i), R.coeff(i)), f"{wname} thresholded sub-band {i}"
    assert L.pdwt_norm1(ptrs, w, 1, out, sp) == 0
    assert abs(out[0] - ref.ref_norm1(R.h)) <= 1e-5 * ref.ref_norm1(R.h)
    assert L.pdwt_inverse(*args, 1) == 0
    ref.ref_inverse(R.h)
    stream.synchronize()
    assert bitexact(d_image.cpu().numpy(), R.image())
    R.close()
    L.pdwt_filters_destroy(f)


# --------------------------------------------------------------------- custom 2-D quadruples (non-separable mode)
def _custom_cases():
    import glob
    return sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ns*2_custom_*.npz")))


@pytest.mark.parametrize("path_", _custom_cases() or [None], ids=lambda p: os.path.basename(p)[:-4] if p else "none")
def test_custom_2d_filter_quadruple_golden(path_):
    """Wavelets::set_filters_forward/_inverse with four len x len filters (wt.cu:560-602, nonseparable.cu:86-106) against
    vectors dumped from the reference's CUDA build (tests/golden/make_golden_custom2d.py)"""
    if path_ is None:
        pytest.skip("no custom-2-D golden vectors committed")
    g = np.load(path_)
    x, (levels, n) = g["x"], g["meta"]
    swt = 1 if os.path.basename(path_).startswith("nsswt2") else 0
    F, I = [g[f"F{j}"] for j in range(4)], [g[f"I{j}"] for j in range(4)]
    for kpath in ("stream", "generic"):   # tiled kernels, then the generic ones
        os.environ["PDWT_PATH"] = kpath
        try:
            W = Wavelets(x, "db%d" % (n // 2), int(levels), do_separable=0, do_swt=swt)
            assert W.set_filters_forward("custom", *F) == 0
            assert W.set_filters_inverse(*I) == 0      # both up front: each direction keeps its own quadruple here
            W.forward()
            for i in range(W.ncoeffs):
                assert bitexact(W.get_coeff(i), g[f"c{i}"]), (kpath, i)
            W.inverse()
            assert bitexact(W.get_image(), g["recon"]), kpath
        finally:
            os.environ.pop("PDWT_PATH", None)


def test_custom_2d_filter_quadruple_live_reference(ref):
    from make_golden_custom2d import declare, quadruple, run_ref
    declare(ref)
    x = rnd((200, 264), 21)
    n, levels = 10, 2
    F, I = quadruple(n, 3), quadruple(n, 4)
    r = run_ref(ref, x, n, levels, 0, F, I)
    W = Wavelets(x, "db5", levels, do_separable=0)
    assert W.set_filters_forward("mine", *F) == 0 and W.set_filters_inverse(*I) == 0
    assert W.set_filters_forward("bad", F[0], F[1]) == -2      # the reference's code for missing 2-D filters
    W.forward()
    for i in range(W.ncoeffs):
        assert bitexact(W.get_coeff(i), r[f"c{i}"]), i
    W.inverse()
    assert bitexact(W.get_image(), r["recon"])
