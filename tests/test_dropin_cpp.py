"""GPU: a C++ client written against the reference's class interface (examples/dropin_demo.cpp includes only "wt.h")
is compiled with g++, linked with libpdwt_b200.so and must reproduce the oracle -- the drop-in boundary of
INTEGRATION.md section 1, exercised from C++ rather than through ctypes."""
import os
import subprocess

import numpy as np
import pytest
from conftest import ROOT, bitexact

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def demo_binary(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("dropin") / "dropin_demo")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda, "include"),
           os.path.join(ROOT, "examples", "dropin_demo.cpp"), "-L" + os.path.join(ROOT, "pdwt_b200"), "-lpdwt_b200",
           "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + os.path.join(ROOT, "pdwt_b200"), "-o", out]
    subprocess.check_call(cmd)
    return out


@pytest.mark.parametrize("shape,wname,levels,sep,swt", [((512, 640), "db7", 3, 1, 0), ((200, 333), "sym4", 2, 1, 0),
                                                         ((128, 128), "haar", 3, 1, 0), ((96, 128), "db3", 2, 1, 1),
                                                         ((128, 160), "db2", 2, 0, 0)])
def test_cpp_client_matches_oracle(demo_binary, tmp_path, shape, wname, levels, sep, swt):
    x = (np.random.default_rng(3).standard_normal(shape) * 50 + 128).astype(np.float32)
    fin, fout = str(tmp_path / "in.f32"), str(tmp_path / "out.f32")
    x.tofile(fin)
    r = subprocess.run([demo_binary, str(shape[0]), str(shape[1]), wname, str(levels), str(sep), str(swt), "10.0", fin,
                        fout], capture_output=True, text=True, check=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][0].split()
    n0, n1 = float(line[1]), float(line[2])
    O = oracle.Wavelets(x, wname, levels, do_separable=sep, do_swt=swt)
    O.forward()
    assert abs(n0 - O.norm1()) <= 1e-5 * O.norm1()
    O.soft_threshold(10.0)
    assert abs(n1 - O.norm1()) <= 1e-5 * O.norm1()
    O.inverse()
    assert bitexact(np.fromfile(fout, np.float32).reshape(shape), O.get_image())
