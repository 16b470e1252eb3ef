"""CPU: closed-form invariants of the algorithm checked on the oracle."""
import numpy as np
import pytest
from conftest import nerr

import oracle

ALL = ["haar"] + [f"db{i}" for i in range(2, 21)] + [f"sym{i}" for i in range(2, 21)] + \
      [f"coif{i}" for i in range(1, 6)] + \
      [f"{p}{s}" for p in ("bior", "rbio")
       for s in ("1.3", "1.5", "2.2", "2.4", "2.6", "2.8", "3.1", "3.3", "3.5", "3.7", "3.9", "4.4", "5.5", "6.8")]


def rnd(shape, seed=0):
    return (np.random.default_rng(seed).standard_normal(shape) * 50 + 128).astype(np.float32)


def test_table_has_72_banks_with_unit_dc_gain():
    assert len(ALL) == 72
    for n in ALL:
        hlen, L, H, IL, IH = oracle.filters(n, do_swt=1)
        assert hlen % 2 == 0 and 2 <= hlen <= 40
        assert abs(float(L.astype(np.float64).sum()) - np.sqrt(2)) < 1e-6     # sum L = sqrt 2 (P5)
        assert abs(float(H.astype(np.float64).sum())) < 1e-6


@pytest.mark.parametrize("wname", ALL)
def test_perfect_reconstruction_all_banks(wname):
    for shape, sep, swt in [((64, 64), 1, 0), ((37, 51), 1, 0), ((40, 36), 0, 0), ((32, 40), 1, 1)]:
        x = rnd(shape, 1)
        W = oracle.Wavelets(x, wname, 2, do_separable=sep, do_swt=swt)
        if W.info.nlevels < 1:
            continue
        W.forward()
        W.inverse()
        assert nerr(W.get_image(), x) < 2e-5, (wname, shape, sep, swt)


@pytest.mark.parametrize("wname", ["haar", "db2", "db7", "sym8", "coif3", "db20"])
def test_parseval_orthogonal_even_sizes(wname):
    x = rnd((256, 192), 2)
    W = oracle.Wavelets(x, wname, 2)
    W.forward()
    e = float((x.astype(np.float64) ** 2).sum())
    assert abs(W.norm2sq() - e) < 2e-5 * e


@pytest.mark.parametrize("wname", ["db2", "db7", "sym8"])
def test_impulse_response_pins_centring(wname):
    """x = delta_p  =>  a[k] = L[m], m = (2k + hlen/2 - p) mod N when m < hlen"""
    N = 64
    hlen, L, H, _, _ = oracle.filters(wname)
    for p in (0, 5, 31, 63):
        x = np.zeros((1, N), dtype=np.float32)
        x[0, p] = 1.0
        W = oracle.Wavelets(x, wname, 1, ndim=1)
        W.forward()
        a, d = W.get_coeff(0)[0], W.get_coeff(1)[0]
        for k in range(N // 2):
            m = (2 * k + hlen // 2 - p) % N
            assert a[k] == (L[m] if m < hlen else 0.0)
            assert d[k] == (H[m] if m < hlen else 0.0)


def test_constant_image_has_no_details():
    x = np.full((64, 64), 3.0, dtype=np.float32)
    W = oracle.Wavelets(x, "db7", 2)
    W.forward()
    assert np.allclose(W.get_coeff(0), 3.0 * 2 ** 2, rtol=1e-5)           # A_L = c * 2^L
    for i in range(1, W.ncoeffs):
        assert np.abs(W.get_coeff(i)).max() < 1e-4


def test_nonseparable_swaps_h_and_v():
    """non-separable H == separable V and vice versa (nonseparable.cu:72-79)"""
    x = rnd((64, 48), 3)
    Ws = oracle.Wavelets(x, "db3", 1)
    Wn = oracle.Wavelets(x, "db3", 1, do_separable=0)
    Ws.forward()
    Wn.forward()
    assert nerr(Wn.get_coeff(1), Ws.get_coeff(2)) < 1e-5
    assert nerr(Wn.get_coeff(2), Ws.get_coeff(1)) < 1e-5
    assert nerr(Wn.get_coeff(0), Ws.get_coeff(0)) < 1e-5
    assert nerr(Wn.get_coeff(3), Ws.get_coeff(3)) < 1e-5


def test_threshold_identities():
    x = rnd((64, 64), 4)
    W = oracle.Wavelets(x, "db2", 2)
    W.forward()
    det = [W.get_coeff(i).astype(np.float64) for i in range(1, W.ncoeffs)]
    a = W.get_coeff(0).astype(np.float64)
    beta = 7.5
    W.soft_threshold(beta)
    want = sum(np.maximum(np.abs(d) - beta, 0).sum() for d in det) + np.abs(a).sum()
    assert abs(W.norm1() - want) < 1e-5 * want
    W2 = oracle.Wavelets(x, "db2", 2)
    W2.forward()
    W2.hard_threshold(beta)
    for i, d in enumerate(det, 1):
        got = W2.get_coeff(i)
        assert np.array_equal(got != 0, np.abs(d.astype(np.float32)) > np.float32(beta))


def test_state_machine_and_clamp():
    x = rnd((40, 40), 5)
    W = oracle.Wavelets(x, "db7", 5)
    assert W.info.nlevels == 1           # ilog2(40 / 13) = 1, wt.cu:156-165
    W.forward()
    W.inverse()
    img = W.get_image()
    W.inverse()                          # refused, wt.cu:274-277
    assert np.array_equal(img, W.get_image())
    assert W.get_coeff(0) is None        # wt.cu:476-479
    bad = oracle.Wavelets(x, "nosuchwavelet", 1)
    assert bad.state == oracle.W_CREATION_ERROR
