"""Randomised parity sweep on the GPU: random shapes (odd, tiny, wide, tall), wavelets, level counts and transform
modes through the public class, every sub-band / reconstruction / proximal result compared BIT FOR BIT with the CPU
oracle.  usage: [PDWT_FUZZ_MODE=swt2] [PDWT_FUZZ_HI=1500] fuzz_gpu.py [cases] [seed]   (exit code 1 and the failing case on the first mismatch)"""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import oracle, pdwt_b200
from pdwt_b200 import Wavelets

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1234)
names = pdwt_b200.wavelet_names()
bitexact = lambda a, b: a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
t0, bad, ran = time.time(), 0, 0
for it in range(n_cases):
    mode = rng.choice(["dwt2", "dwt2", "dwt2", "swt2", "ns2", "nsswt2", "dwt1", "swt1"])
    if os.environ.get("PDWT_FUZZ_MODE"):     # one transform family only
        mode = os.environ["PDWT_FUZZ_MODE"]
    wname = names[rng.integers(len(names))]
    if mode in ("ns2", "nsswt2") and rng.random() < 0.7:
        wname = rng.choice(["db2", "db3", "db4", "sym4", "db7", "coif2", "haar"])
    small = rng.random() < 0.35
    hi = 96 if small else (260 if mode in ("ns2", "nsswt2", "swt2") else 700)
    if os.environ.get("PDWT_FUZZ_HI") and not small:   # larger planes: several column tiles / row chunks per level
        hi = int(os.environ["PDWT_FUZZ_HI"])
    Nr = int(rng.integers(1 if mode.endswith("1") else 8, hi))
    Nc = int(rng.integers(8, hi * (3 if mode.endswith("1") else 1)))
    if rng.random() < 0.5:
        Nr, Nc = (Nr + 3) & ~3, (Nc + 7) & ~7          # the aligned fast paths
    levels = int(rng.integers(1, 5))
    batch = int(rng.integers(1, 4)) if rng.random() < 0.3 else 1
    sep, swt, ndim = (0 if mode.startswith("ns") else 1), (1 if "swt" in mode else 0), (1 if mode.endswith("1") else 2)
    Nr = max(Nr, 1 if ndim == 1 else 2)
    x = (rng.standard_normal((batch, Nr, Nc) if batch > 1 else (Nr, Nc)) * 50 + 128).astype(np.float32)
    tag = f"#{it} {mode} {wname} L{levels} {Nr}x{Nc} b{batch}"
    try:
        W = Wavelets(x, wname, levels, do_separable=sep, do_swt=swt, ndim=ndim)
        Os = [oracle.Wavelets(x[p] if batch > 1 else x, wname, levels, do_separable=sep, do_swt=swt, ndim=ndim)
              for p in range(batch)]
        if W.state == pdwt_b200.W_CREATION_ERROR or W.info.nlevels < 1:
            assert all(O.state == oracle.W_CREATION_ERROR or O.info.nlevels < 1 for O in Os), tag + " creation state"
            continue
        assert W.info.nlevels == Os[0].info.nlevels, tag + " nlevels"
        ran += 1
        W.forward()
        for O in Os: O.forward()
        op = rng.choice(["none", "soft", "hard", "group", "shrink", "linf"])
        args = (float(rng.uniform(1, 60)), int(rng.integers(2)), int(rng.integers(2)))
        for obj in [W] + Os:
            if op == "soft": obj.soft_threshold(*args)
            elif op == "hard": obj.hard_threshold(*args)
            elif op == "group": obj.group_soft_threshold(*args)
            elif op == "shrink": obj.shrink(args[0] / 60, args[1])
            elif op == "linf": obj.proj_linf(*args[:2])
        for i in range(W.ncoeffs):
            c = W.get_coeff(i)
            for p, O in enumerate(Os):
                assert bitexact(c[p] if batch > 1 else c, O.get_coeff(i)), f"{tag} op={op}{args} sub-band {i} plane {p}"
        n1 = np.atleast_1d(W.norm1())
        for p, O in enumerate(Os):
            assert abs(n1[p] - O.norm1()) <= 1e-5 * max(abs(O.norm1()), 1e-20), f"{tag} norm1 plane {p}"
        W.inverse()
        rec = W.get_image()
        for p, O in enumerate(Os):
            O.inverse()
            assert bitexact(rec[p] if batch > 1 else rec, O.get_image()), f"{tag} op={op} reconstruction plane {p}"
    except AssertionError as e:
        print("MISMATCH:", e, flush=True)
        bad += 1
        if bad >= 5:
            break
    except Exception as e:
        print("ERROR:", tag, repr(e), flush=True)
        bad += 1
        if bad >= 5:
            break
print(f"fuzz: {it + 1} cases drawn, {ran} transformed, {bad} failures, {time.time() - t0:.1f} s", flush=True)
sys.exit(1 if bad else 0)
