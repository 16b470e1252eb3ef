"""CPU: the oracle (oracle/pdwt_oracle.c) against the golden vectors dumped from the reference's own CUDA build
on a B200 (tests/golden/*.npz, written by tests/golden/make_golden.py; pinning summary in pin_report.json).
Bar: transform buffers and thresholded coefficients bit-exact, norms 1e-5 relative (cuBLAS summation order)."""
import os

import numpy as np
import pytest
from cases import CASES, THRESH, THRESH_CASES, make_input
from conftest import bitexact, nerr

import oracle


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_cuda(case, golden_dir):
    name, shape, wname, levels, sep, swt, ndim = case
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    x = make_input(name, shape)
    W = oracle.Wavelets(x, wname, levels, do_separable=sep, do_swt=swt, ndim=ndim)
    assert [W.info.nlevels, W.info.hlen, W.info.ndims] == list(g["meta"])  # level clamp, hlen, 1-D detection
    W.forward()
    for i in range(W.ncoeffs):
        assert bitexact(W.get_coeff(i), g[f"c{i}"]), f"sub-band {i}"
    assert abs(W.norm1() - g["norm1"]) <= 1e-5 * abs(g["norm1"])
    assert abs(W.norm2sq(ref_1d_bug=1) - g["norm2sq"]) <= 1e-5 * abs(g["norm2sq"])
    W.inverse()
    assert bitexact(W.get_image(), g["recon"])
    assert nerr(W.get_image(), x) < 2e-5  # perfect reconstruction
    if name not in THRESH_CASES:
        return
    for tag, kind, beta, app, nrm in THRESH:
        W = oracle.Wavelets(x, wname, levels, do_separable=sep, do_swt=swt, ndim=ndim)
        W.forward()
        getattr(W, f"{kind}_threshold")(beta, app, nrm)
        for i in range(W.ncoeffs):
            assert bitexact(W.get_coeff(i), g[f"{tag}_c{i}"]), f"{tag} sub-band {i}"
        assert abs(W.norm1() - g[f"{tag}_norm1"]) <= 1e-5 * abs(g[f"{tag}_norm1"])
        if tag == "soft":
            W.inverse()
            assert bitexact(W.get_image(), g["soft_recon"])


def test_pin_report_is_green(golden_dir):
    """the committed pinning run: every transform buffer bit-exact, only norms differ (<= 1e-6)"""
    import json
    rep = json.load(open(os.path.join(golden_dir, "pin_report.json")))
    for name, r in rep.items():
        if name.startswith("_"):
            continue
        for key, err in r["not_bitexact"].items():
            assert "norm" in key, (name, key)
            assert err < 1e-6
