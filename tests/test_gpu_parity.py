"""GPU (-m gpu): the CUDA path, called through the C ABI, against
  (1) the golden vectors of the reference's own CUDA build (tests/golden/*.npz),
  (2) the CPU oracle on seeded inputs (sizes the oracle finishes in seconds),
  (3) size-independent properties at BASELINE.json's full sizes.
Bar (north_star): <= 1e-5 relative (max|d| / max|ref|) for the float transforms -- in practice the kernels keep the
reference's FMA order and the assertions below demand BIT-EXACT buffers -- bit-exact Haar and thresholds, norms
1e-5 relative."""
import os

import numpy as np
import pytest
from cases import CASES, THRESH, THRESH_CASES, make_input
from conftest import bitexact, nerr

import oracle
import pdwt_b200
from pdwt_b200 import Wavelets

pytestmark = pytest.mark.gpu

TOL = 1e-5  # north_star tolerance for float32 transforms


def rnd(shape, seed=0):
    return (np.random.default_rng(seed).standard_normal(shape) * 50 + 128).astype(np.float32)


@pytest.fixture(params=["stream", "fused", "generic"])
def path(request, monkeypatch):
    """the three kernel families of the separable 2-D DWT (pdwt_capi.cu path_cap); the others ignore the switch"""
    monkeypatch.delenv("PDWT_FORCE_GENERIC", raising=False)
    monkeypatch.setenv("PDWT_PATH", request.param)
    return request.param


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_golden_vectors(case, golden_dir, path):
    name, shape, wname, levels, sep, swt, ndim = case
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    x = make_input(name, shape)
    W = Wavelets(x, wname, levels, do_separable=sep, do_swt=swt, ndim=ndim)
    w = W.info
    assert [w.nlevels, w.hlen, w.ndims] == list(g["meta"])
    W.forward()
    assert W.state == pdwt_b200.W_FORWARD
    for i in range(W.ncoeffs):
        c = W.get_coeff(i)
        assert nerr(c, g[f"c{i}"]) <= TOL, f"sub-band {i}"
        assert bitexact(c, g[f"c{i}"]), f"sub-band {i} not bit-exact"
    assert abs(W.norm1() - g["norm1"]) <= 1e-5 * abs(g["norm1"])
    if ndim == 2 and w.ndims == 2:   # 1-D norm2sq: the reference sums asum() (wt.cu:389); we return the true value
        assert abs(W.norm2sq() - g["norm2sq"]) <= 1e-5 * abs(g["norm2sq"])
    W.inverse()
    assert bitexact(W.get_image(), g["recon"])
    if name not in THRESH_CASES:
        return
    for tag, kind, beta, app, nrm in THRESH:
        W = Wavelets(x, wname, levels, do_separable=sep, do_swt=swt, ndim=ndim)
        W.forward()
        getattr(W, f"{kind}_threshold")(beta, app, nrm)
        for i in range(W.ncoeffs):
            assert bitexact(W.get_coeff(i), g[f"{tag}_c{i}"]), f"{tag} sub-band {i}"
        assert abs(W.norm1() - g[f"{tag}_norm1"]) <= 1e-5 * abs(g[f"{tag}_norm1"])
        if tag == "soft":
            W.inverse()
            assert bitexact(W.get_image(), g["soft_recon"])


ORACLE_CASES = [
    # shape, wname, levels, sep, swt, ndim
    ((1024, 1024), "db7", 3, 1, 0, 2), ((512, 1280), "sym8", 2, 1, 0, 2), ((768, 1028), "db2", 2, 1, 0, 2),
    ((640, 600), "db3", 2, 1, 0, 2), ((520, 776), "db4", 2, 1, 0, 2), ((512, 520), "db5", 2, 1, 0, 2),
    ((384, 644), "db6", 2, 1, 0, 2), ((400, 900), "db9", 2, 1, 0, 2), ((448, 704), "coif3", 2, 1, 0, 2),
    ((330, 1100), "sym7", 2, 1, 0, 2),
    ((512, 512), "db7", 3, 1, 0, 2), ((1024, 768), "db7", 3, 1, 0, 2), ((300, 520), "sym8", 3, 1, 0, 2),
    ((257, 255), "db4", 3, 1, 0, 2), ((130, 66), "db10", 2, 1, 0, 2), ((512, 384), "coif1", 4, 1, 0, 2),
    ((640, 648), "bior4.4", 3, 1, 0, 2), ((200, 200), "db12", 2, 1, 0, 2), ((96, 1000), "db2", 4, 1, 0, 2),
    ((511, 512), "haar", 5, 1, 0, 2), ((16, 8191), "db4", 5, 1, 0, 1), ((7, 4096), "haar", 6, 1, 0, 1),
    ((256, 256), "sym8", 4, 1, 1, 2), ((100, 90), "db3", 3, 1, 1, 2), ((8, 1024), "db5", 4, 1, 1, 1),
    ((256, 256), "db7", 2, 0, 0, 2), ((97, 131), "sym4", 3, 0, 0, 2), ((64, 80), "db2", 2, 0, 1, 2),
    # batched 1-D transforms: all levels from one launch (a row in shared memory); long rows fall back to per-level kernels
    ((64, 4096), "db7", 3, 1, 0, 1), ((33, 1001), "sym8", 4, 1, 0, 1), ((5, 12000), "db2", 6, 1, 0, 1),
    ((3, 20000), "db3", 3, 1, 0, 1), ((40, 514), "db10", 2, 1, 0, 1),
    # odd / unaligned widths large enough for interior tiles of the tile-fused kernels (4-byte asynchronous staging)
    ((513, 1031), "db7", 3, 1, 0, 2), ((771, 1290), "sym8", 2, 1, 0, 2), ((640, 1026), "db2", 3, 1, 0, 2),
    # non-separable SWT through the tiled kernels (dilations 1, 2, 4, 8; odd sizes; a long filter)
    ((200, 264), "sym4", 3, 0, 1, 2), ((130, 96), "db7", 2, 0, 1, 2), ((97, 131), "db2", 4, 0, 1, 2),
    ((160, 192), "haar", 3, 0, 1, 2),
    # separable SWT, streaming inverse: several column tiles and row chunks, sizes that are no multiple of the dilation,
    # a dilation whose halo is wider than a short image allows (falls back), short filters
    ((300, 1100), "sym8", 4, 1, 1, 2), ((1000, 700), "db3", 3, 1, 1, 2), ((523, 1301), "db5", 4, 1, 1, 2),
    ((96, 2050), "haar", 4, 1, 1, 2), ((2049, 130), "coif2", 3, 1, 1, 2),
    # batched 1-D SWT: all levels of a row in one launch (extended row buffers in shared memory); odd lengths, a long
    # filter, a row too short for the coarsest dilation (falls back to the per-level kernels)
    ((33, 1001), "sym8", 3, 1, 1, 1), ((5, 4096), "db4", 3, 1, 1, 1), ((4, 600), "db7", 4, 1, 1, 1),
    ((3, 130), "db10", 3, 1, 1, 1), ((2, 12000), "db2", 5, 1, 1, 1),
]


@pytest.mark.parametrize("case", ORACLE_CASES, ids=[f"{c[1]}-{c[0][0]}x{c[0][1]}-s{c[3]}w{c[4]}d{c[5]}" for c in ORACLE_CASES])
def test_against_oracle(case, path):
    shape, wname, levels, sep, swt, ndim = case
    x = rnd(shape, 11)
    W = Wavelets(x, wname, levels, do_separable=sep, do_swt=swt, ndim=ndim)
    O = oracle.Wavelets(x, wname, levels, do_separable=sep, do_swt=swt, ndim=ndim)
    assert W.info.nlevels == O.info.nlevels and W.info.hlen == O.info.hlen
    W.forward()
    O.forward()
    for i in range(W.ncoeffs):
        assert bitexact(W.get_coeff(i), O.get_coeff(i)), f"sub-band {i}: {nerr(W.get_coeff(i), O.get_coeff(i))}"
    assert abs(W.norm1() - O.norm1()) <= 1e-5 * O.norm1()
    assert abs(W.norm2sq() - O.norm2sq()) <= 1e-5 * O.norm2sq()
    W.soft_threshold(12.5, 1, 1)
    O.soft_threshold(12.5, 1, 1)
    for i in range(W.ncoeffs):
        assert bitexact(W.get_coeff(i), O.get_coeff(i))
    W.inverse()
    O.inverse()
    assert bitexact(W.get_image(), O.get_image())


@pytest.mark.parametrize("shape,wname,levels,batch", [((512, 1024), "db7", 2, 1), ((256, 768), "sym8", 2, 3),
                                                      ((384, 512), "db6", 1, 2), ((200, 520), "db9", 2, 1),
                                                      ((1024, 1024), "db7", 3, 1)])
def test_inverse_tma_variant_against_oracle(shape, wname, levels, batch, monkeypatch):
    """k_inv2d_tma (producer warp + tensor-map ring; picked automatically for large batched launches) forced onto small
    cases -- edge super-slots in rows and columns, strips cut by the plane, several chunks -- bit-exact against the
    oracle and against the default level kernel (separable.cu:246-328)."""
    x = rnd((batch,) + shape if batch > 1 else shape, 21)
    rec = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("PDWT_INV_TMA", mode)
        W = Wavelets(x, wname, levels)
        W.forward()
        W.soft_threshold(3.0, 0, 1)
        L = pdwt_b200.lib()
        L.pdwt_profile_begin()
        W.inverse()
        ents = (pdwt_b200.ProfileEntry * 64)()
        names = [ents[k].name.decode() for k in range(L.pdwt_profile_end(ents, 64))]
        assert any(n.startswith("k_inv2d_tma") for n in names) == (mode == "1"), names   # the variant under test did run
        rec[mode] = W.get_image()
    assert bitexact(rec["1"], rec["0"])
    for p in range(batch):
        O = oracle.Wavelets(x[p] if batch > 1 else x, wname, levels)
        O.forward()
        O.soft_threshold(3.0, 0, 1)
        O.inverse()
        assert bitexact(rec["1"][p] if batch > 1 else rec["1"], O.get_image())


def test_every_wavelet_roundtrips_and_matches_oracle():
    x = rnd((96, 160), 5)
    for wname in pdwt_b200.wavelet_names():
        W = Wavelets(x, wname, 2)
        O = oracle.Wavelets(x, wname, 2)
        W.forward()
        O.forward()
        for i in range(W.ncoeffs):
            assert bitexact(W.get_coeff(i), O.get_coeff(i)), (wname, i)
        W.inverse()
        assert nerr(W.get_image(), x) < 2e-5, wname


def test_c2_full_size_properties():
    """BASELINE configs[1]: 4096^2 db7 L3 -- perfect reconstruction, Parseval, linearity, fused == generic."""
    x = rnd((4096, 4096), 0)
    W = Wavelets(x, "db7", 3)
    W.forward()
    e = float((x.astype(np.float64) ** 2).sum())
    assert abs(W.norm2sq() - e) <= 2e-5 * e
    coeffs = [W.get_coeff(i) for i in range(W.ncoeffs)]
    W.inverse()
    assert nerr(W.get_image(), x) < 1e-5
    for other in ("fused", "generic"):
        os.environ["PDWT_PATH"] = other
        try:
            G = Wavelets(x, "db7", 3)
            G.forward()
            for i in range(G.ncoeffs):
                assert bitexact(G.get_coeff(i), coeffs[i]), (other, i)
            G.inverse()
            assert bitexact(G.get_image(), W.get_image()), other
        finally:
            del os.environ["PDWT_PATH"]
    # linearity: W(2x) == 2 W(x) exactly (power-of-two scaling commutes with every rounding)
    W2 = Wavelets(2 * x, "db7", 3)
    W2.forward()
    for i in (0, 1, 5, 9):
        assert bitexact(W2.get_coeff(i), 2 * coeffs[i])
    # a 1024x1024 corner of the oracle's answer at full size would take long; check one level-1 strip instead
    O = oracle.Wavelets(x[:, :], "db7", 1)
    O.forward()
    W1 = Wavelets(x, "db7", 1)
    W1.forward()
    for i in range(4):
        assert bitexact(W1.get_coeff(i), O.get_coeff(i))


def test_c3_swt_and_c4_nonseparable_properties():
    x = rnd((2048, 2048), 1)
    W = Wavelets(x, "sym8", 4, do_swt=1)           # BASELINE configs[2]
    assert W.info.nlevels == 4
    W.forward()
    W.inverse()
    assert nerr(W.get_image(), x) < 1e-5
    y = rnd((1024, 1024), 2)                        # configs[3] sequence at a size the oracle checks quickly
    W = Wavelets(y, "db7", 2, do_separable=0)
    O = oracle.Wavelets(y, "db7", 2, do_separable=0)
    W.forward(); O.forward()
    n0 = W.norm1()
    assert abs(n0 - O.norm1()) <= 1e-5 * n0
    W.soft_threshold(10.0); O.soft_threshold(10.0)
    n1 = W.norm1()
    assert n1 < n0 and abs(n1 - O.norm1()) <= 1e-5 * n1
    W.inverse(); O.inverse()
    assert bitexact(W.get_image(), O.get_image())
    W.forward()                                     # forward after inverse uses the FORWARD filters (SURVEY B6 fixed)
    O2 = oracle.Wavelets(W.get_image(), "db7", 2, do_separable=0)
    O2.forward()
    assert bitexact(W.get_coeff(3), O2.get_coeff(3))


def test_batch_planes_are_independent():
    """configs[4] in miniature: a batch equals the same images transformed one by one"""
    xs = np.stack([rnd((256, 320), 100 + i) for i in range(5)])
    for wname, sep, swt in [("db7", 1, 0), ("haar", 1, 0), ("sym4", 1, 1), ("db2", 0, 0)]:
        B = Wavelets(xs, wname, 3, do_separable=sep, do_swt=swt)
        assert B.batch == 5
        B.forward()
        n1 = B.norm1()
        for p in range(5):
            S = Wavelets(xs[p], wname, 3, do_separable=sep, do_swt=swt)
            S.forward()
            for i in range(S.ncoeffs):
                assert bitexact(B.get_coeff(i)[p], S.get_coeff(i)), (wname, p, i)
            assert abs(n1[p] - S.norm1()) <= 1e-6 * n1[p]
        B.hard_threshold(20.0, 1, 0)
        B.inverse()
        rec = B.get_image()
        S = Wavelets(xs[3], wname, 3, do_separable=sep, do_swt=swt)
        S.forward(); S.hard_threshold(20.0, 1, 0); S.inverse()
        assert bitexact(rec[3], S.get_image())


def test_state_machine_matches_reference():
    x = rnd((64, 64), 7)
    W = Wavelets(x, "db7", 9)
    assert W.info.nlevels == 2 and W.state == pdwt_b200.W_INIT      # clamp ilog2(64/13), wt.cu:156-165
    W.forward()
    c1 = W.get_coeff(1)
    W.inverse()
    assert W.state == pdwt_b200.W_INVERSE
    img = W.get_image()
    W.inverse()                                   # refused: wt.cu:274-277
    W.soft_threshold(5.0)                         # refused: wt.cu:311-314
    assert W.get_coeff(1) is None                 # refused: wt.cu:476-479
    assert bitexact(img, W.get_image())
    W.set_image(x)                                # resets the state, wt.cu:433
    assert W.state == pdwt_b200.W_INIT
    W.forward()
    assert bitexact(W.get_coeff(1), c1)
    bad = Wavelets(x, "nosuchwavelet", 2)         # error state instead of an endless loop
    assert bad.state == pdwt_b200.W_CREATION_ERROR
    bad.forward()
    assert bad.state == pdwt_b200.W_CREATION_ERROR
    one_d = Wavelets(rnd((1, 256), 3), "db2", 2, do_separable=0)    # wt.cu:133-142
    assert one_d.info.ndims == 1 and one_d.do_separable == 1
    cp = W.copy()                                 # deep copy, wt.cu:191-222
    W.soft_threshold(1e9)
    assert bitexact(cp.get_coeff(1), c1) and not W.get_coeff(1).any()


def test_set_coeff_and_device_image():
    import torch
    x = rnd((128, 128), 8)
    W = Wavelets(torch.from_numpy(x).cuda(), "sym4", 2)      # memisonhost = 0, wt.cu:121-124
    W.forward()
    z = np.zeros(W.coeff_shape(0), np.float32)
    W.set_coeff(z, 0)
    assert not W.get_coeff(0).any()
    O = oracle.Wavelets(x, "sym4", 2)
    O.forward(); O.set_coeff(z, 0); O.inverse()
    W.inverse()
    assert bitexact(W.get_image(), O.get_image())


def test_custom_filters_cdf97_like():
    """set_filters_forward/inverse (wt.cu:560-602): feeding a built-in bank back as 'custom' gives identical results"""
    x = rnd((128, 96), 9)
    hlen, L, H, IL, IH = pdwt_b200.filters("bior4.4")
    W = Wavelets(x, "db5", 2)                      # same length (10) as bior4.4
    assert W.set_filters_forward("mybank", L, H) == 0
    assert W.set_filters_inverse(IL, IH) == 0
    R = Wavelets(x, "bior4.4", 2)
    W.forward(); R.forward()
    for i in range(W.ncoeffs):
        assert bitexact(W.get_coeff(i), R.get_coeff(i))
    W.inverse(); R.inverse()
    assert bitexact(W.get_image(), R.get_image())
    assert W.set_filters_forward("toolong", np.zeros(41, np.float32), np.zeros(41, np.float32)) == -1


@pytest.mark.parametrize("env", [{"PDWT_SMALL_PX": "100000000"}, {"PDWT_SMALL_PX": "70000"},
                                 {"PDWT_PATH": "fused", "PDWT_FUSED_TILE": "0"}, {"PDWT_PATH": "fused", "PDWT_FUSED_TILE": "1"},
                                 {"PDWT_PDL": "1"}, {"PDWT_PDL": "2"}, {"PDWT_PDL": "2", "PDWT_SMALL_PX": "70000"}],
                         ids=lambda e: ",".join(f"{k[5:]}={v}" for k, v in e.items()))
def test_dispatch_switches_do_not_change_a_bit(env, monkeypatch):
    """per-level family choice (stream for large levels, tile kernels for small ones), both tile shapes and the
    programmatic-dependent-launch modes are scheduling decisions: results stay bit-identical to the oracle"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for shape, wname, levels in (((1024, 768), "db7", 3), ((520, 776), "db4", 3), ((512, 1280), "sym8", 2)):
        x = rnd(shape, 21)
        W = Wavelets(x, wname, levels)
        O = oracle.Wavelets(x, wname, levels)
        for _ in range(2):                      # twice: back-to-back launches exercise the dependent-launch chain
            W.set_image(x)
            W.forward()
        O.forward()
        for i in range(W.ncoeffs):
            assert bitexact(W.get_coeff(i), O.get_coeff(i)), (env, shape, i)
        W.inverse(); O.inverse()
        assert bitexact(W.get_image(), O.get_image()), (env, shape)


def test_async_copies_pipeline_over_streams():
    """set_async(True): host copies are only enqueued on the object's stream; several objects on their own streams
    overlap H2D / kernels / D2H and still deliver exactly the blocking result"""
    import torch
    xs = [rnd((512, 640), 40 + i) for i in range(3)]
    ref = []
    for x in xs:
        W = Wavelets(x, "db7", 3)
        W.forward(); W.soft_threshold(5.0); W.inverse()
        ref.append(W.get_image())
    h_in = [torch.from_numpy(x).pin_memory() for x in xs]
    h_out = [torch.empty((512, 640), dtype=torch.float32).pin_memory() for _ in xs]
    Ws = [Wavelets(None, "db7", 3, shape=(512, 640)) for _ in xs]
    streams = [torch.cuda.Stream() for _ in xs]
    for W, s in zip(Ws, streams):
        W.set_stream(s)
        W.set_async(True)
    for rep in range(3):
        for W, a, b in zip(Ws, h_in, h_out):
            W.set_image(a.numpy())
            W.forward(); W.soft_threshold(5.0); W.inverse()
            W.get_image(b.numpy())
    for W in Ws:
        W.sync()
    for b, r in zip(h_out, ref):
        assert bitexact(b.numpy(), r)


def test_norms_after_threshold_come_from_the_threshold_launch(monkeypatch):
    """soft/hard_threshold leave the L1 / L2 sums of their result behind, so the norm that follows reads no
    coefficients -- same values as a fresh reduction (1e-6), invalidated by everything that changes coefficients"""
    L = pdwt_b200.lib()
    for shape, wname, levels, kw in (((512, 768), "db7", 3, {}), ((96, 1000), "db4", 3, {"ndim": 1}),
                                     ((128, 160), "sym4", 2, {"do_swt": 1}), ((3, 200, 264), "db3", 2, {})):
        x = rnd(shape, 31)
        for kind, app in (("soft", 0), ("soft", 1), ("hard", 0), ("hard", 1)):
            W = Wavelets(x, wname, levels, **kw)
            W.forward()
            getattr(W, f"{kind}_threshold")(15.0, app, 1)
            before = L.pdwt_launch_count()
            n1, n2 = np.atleast_1d(W.norm1()), np.atleast_1d(W.norm2sq())
            assert L.pdwt_launch_count() == before, "norms after a threshold must not launch a kernel"
            monkeypatch.setenv("PDWT_NORM_CACHE", "0")
            F = Wavelets(x, wname, levels, **kw)
            F.forward()
            getattr(F, f"{kind}_threshold")(15.0, app, 1)
            f1, f2 = np.atleast_1d(F.norm1()), np.atleast_1d(F.norm2sq())
            monkeypatch.delenv("PDWT_NORM_CACHE")
            assert np.all(np.abs(n1 - f1) <= 1e-6 * np.abs(f1)) and np.all(np.abs(n2 - f2) <= 1e-6 * np.abs(f2)), (shape, kind, app)
            for i in range(W.ncoeffs):
                assert bitexact(W.get_coeff(i), F.get_coeff(i))
            W.shrink(0.5)                                  # changes the coefficients: the cache must not survive
            before = L.pdwt_launch_count()
            s1 = np.atleast_1d(W.norm1())
            assert L.pdwt_launch_count() > before
            assert np.all(np.abs(s1 - f1 / np.float32(1.5)) <= 1e-5 * np.abs(f1))


def test_norm_publication_repeats_and_unaligned_planes():
    """the norms travel to the host as tagged 8-byte words written by the last block of the launch (HostPublish,
    pdwt_common.cuh): many launches in a row on one object (tickets and scratch re-armed by the kernel itself), thresholds
    back to back, and odd-sized batched planes (plane strides that are not multiples of 4 floats: scalar heads and tails
    of the flat tile walk) -- every value against the oracle at 1e-5, thresholded coefficients bit-exact"""
    for shape, wname, levels, kw in (((5, 3, 37), "db2", 3, {"ndim": 1}), ((3, 33, 47), "db3", 2, {}),
                                     ((2, 57, 64), "sym4", 2, {"do_swt": 1}), ((4096, 512), "db7", 3, {})):
        x = rnd(shape, 77)
        W = Wavelets(x, wname, levels, **kw)
        planes = x if x.ndim == 3 else x[None]
        Os = [oracle.Wavelets(p, wname, levels, **kw) for p in planes]
        W.forward()
        for O in Os:
            O.forward()
        ref1 = np.array([O.norm1() for O in Os], dtype=np.float64)
        ref2 = np.array([O.norm2sq() for O in Os], dtype=np.float64)
        for _ in range(25):   # uncached reductions, alternating modes
            n1, n2 = np.atleast_1d(W.norm1()), np.atleast_1d(W.norm2sq())
            assert np.all(np.abs(n1 - ref1) <= 1e-5 * np.abs(ref1)) and np.all(np.abs(n2 - ref2) <= 1e-5 * np.abs(ref2)), shape
        for it in range(6):   # thresholds back to back: only the last one's sums may be read
            W.soft_threshold(3.0, it & 1, 1)
            W.hard_threshold(2.0, 0, 0)
            for O in Os:
                O.soft_threshold(3.0, it & 1, 1)
                O.hard_threshold(2.0, 0, 0)
            ref1 = np.array([O.norm1() for O in Os], dtype=np.float64)
            ref2 = np.array([O.norm2sq() for O in Os], dtype=np.float64)
            n1, n2 = np.atleast_1d(W.norm1()), np.atleast_1d(W.norm2sq())
            assert np.all(np.abs(n1 - ref1) <= 1e-5 * np.abs(ref1) + 1e-30), (shape, it)
            assert np.all(np.abs(n2 - ref2) <= 1e-5 * np.abs(ref2) + 1e-30), (shape, it)
        for i in range(W.ncoeffs):
            c = W.get_coeff(i)
            c = c if x.ndim == 3 else c[None]
            for b, O in enumerate(Os):
                assert bitexact(c[b].reshape(-1), np.asarray(O.get_coeff(i)).reshape(-1)), (shape, i, b)
