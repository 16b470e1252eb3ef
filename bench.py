#!/usr/bin/env python3
"""bench.py -- the headline measurement of pdwt_b200 (BASELINE.json: "Mpixels/s fwd+inv 2D DWT db7 L3 4096^2;
achieved HBM GB/s vs B200 peak").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c5img]

One "step" = `forward(); inverse();` of ONE `Wavelets` object holding one 4096x4096 float32 image, db7, 3 levels,
separable (BASELINE.json configs[1]).  Steps rotate over R=4 objects with distinct images and distinct device
buffers (about 1 GiB touched between two uses of the same object, 8x the 126 MB L2), so no step finds its inputs
in L2 from the previous one.  Under torchrun (N>1) every rank runs the same per-GPU workload on its own images
(weak scaling; the path has no data-path collective -- DESIGN.md section 6); the time is the max over ranks.

JSON line (rank 0):
  value      Mpixels/s, device-resident inputs, CUDA events on the launching stream
  e2e        same metric through the public API with pinned HOST buffers: set_image (H2D) -> forward -> inverse
             -> get_image (D2H) inside the timed region
  e2e        ... `e2e.value`: 4 objects on 4 streams with set_async (copies and kernels of consecutive steps overlap);
             `e2e.sync`: the reference's blocking calls
  roofline   the dominant kernel (largest share of the step): algorithmic bytes per launch / average duration vs the
             measured HBM peak, duration = one CUDA-event pair around every launch inside the steps (library profiler,
             separate pass over the same steps); next to it the same kernel launched K times in a row between two
             events, in plain stream order and with programmatic dependent launch
  cpu_baseline  the CPU oracle (port of the reference kernels, OpenMP) on this box's host cores, bounded sample
`--impl reference` times the reference's own implementation: PDWT has no CPU path, its stock CUDA kernels compiled
unmodified for sm_100 (oracle/_ref/libpdwt_ref.so, built by `make -C oracle ref`) run on the same GPU through the
reference's own Wavelets class; if that library cannot be loaded, the CPU oracle port is timed instead.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (Nr, Nc, wavelet, levels, images per GPU and step)
    "c2": (4096, 4096, "db7", 3, 1),     # BASELINE.json configs[1] -- the configuration the metric is quoted on
    "c5img": (2048, 2048, "db7", 3, 1),  # one image of configs[4]
    "c5": (2048, 2048, "db7", 3, 64),    # configs[4]: 512 images of 2048^2 over 8 GPUs = 64 per GPU, one batched object
    "c2b8": (4096, 4096, "db7", 3, 8),   # north_star's "batched 4096x4096": 8 images of configs[1] in one batched object
}
ROTATE = int(os.environ.get("PDWT_BENCH_ROTATE", "4"))
METRIC = "Mpixels/s fwd+inv 2D DWT db7 L3 4096^2"


def config_of(workload):
    """the `config` object of the JSON line: the same for both arms (the driver compares them)"""
    Nr, Nc, wname, levels, nimg = WORKLOADS[workload]
    npx = Nr * Nc * nimg
    return {"workload": f"{workload}: {Nr}x{Nc} float32, {wname}, {levels} levels, separable DWT, "
                        f"forward()+inverse() of {nimg} image(s) per GPU and step "
                        f"(BASELINE.json configs[{1 if workload == 'c2' else 4}])",
            "l2": (f"steps rotate over {ROTATE} Wavelets objects with distinct images/buffers (~1 GiB touched "
                   f"between reuses, > 126 MB L2); no explicit flush") if nimg == 1 else
                  f"one batched object of {nimg} images (inputs {4 * npx >> 20} MiB > 126 MB L2)",
            "per_gpu": "every rank runs this workload on its own images; no data-path collective"}


def seeded_image(shape, seed):
    return (np.random.default_rng(seed).standard_normal(shape) * 50 + 128).astype(np.float32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed regions run (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self.t.join(timeout=5)

    def summary(self):
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm   # samples taken under load
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def dist_setup(n_gpus):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return world, rank, local


def max_over_ranks(ms, world):
    if world == 1:
        return ms
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def algorithmic_bytes(tag):
    """k_fwd2d[R x C] / k_inv2d[R x C]: one level over an R x C plane reads 4*R*C bytes and writes 4*R*C bytes
    (8 B per pixel of the level, fp32, compulsory traffic only)."""
    try:
        dims = tag[tag.index("[") + 1:tag.index("]")].split("x")
        return 8.0 * int(dims[0]) * int(dims[1])
    except ValueError:
        return None


# ============================================================================================ our arm
def run_ours(args):
    import torch
    import pdwt_b200
    from pdwt_b200 import Wavelets

    world, rank, local = dist_setup(args.gpus)
    Nr, Nc, wname, levels, nimg = WORKLOADS[args.workload]
    npx = Nr * Nc * nimg
    rotate = ROTATE if nimg == 1 else 1      # a 64-image batch (1 GiB) is 8x the L2 by itself
    L = pdwt_b200.lib()
    if L.pdwt_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- pdwt_b200 has no CPU fallback")

    shape = (Nr, Nc) if nimg == 1 else (nimg, Nr, Nc)
    imgs = [seeded_image(shape, 1000 * rank + i) for i in range(rotate)]
    Ws = [Wavelets(torch.from_numpy(im).cuda(), wname, levels) for im in imgs]
    stream = torch.cuda.current_stream()
    ROT = rotate

    def step(i):
        W = Ws[i % ROT]
        W.forward()
        W.inverse()

    for i in range(args.warmup):
        step(i)
    # parity spot check before anything is timed: perfect reconstruction of the rotating images
    rec = Ws[(args.warmup - 1) % ROT].get_image() if args.warmup else None
    if rec is not None:
        ref = imgs[(args.warmup - 1) % ROT]
        err = float(np.abs(rec - ref).max() / np.abs(ref).max())
        if not err < 1e-5:
            raise SystemExit(f"bench.py: fwd+inv does not reconstruct the image (err {err:.3e})")

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = L.pdwt_launch_count()
    with ClockSampler(local) as clk:
        barrier(world)
        e0.record(stream)
        for i in range(args.steps):
            step(i)
        e1.record(stream)
        barrier(world)
        ms_dev = max_over_ranks(e0.elapsed_time(e1), world)
        launches = L.pdwt_launch_count() - launches0

        # ---- end to end through the public API with pinned host buffers.  Every step copies its input image from
        # pinned host memory (H2D), transforms it and reads the reconstruction back into pinned host memory (D2H),
        # all inside the timed region.  `sync`: the reference's blocking calls, one object after the other.
        # `pipelined` (the e2e headline): the same four calls per step, but each of the ROT objects owns a stream and
        # its copies are only enqueued (Wavelets.set_async), so H2D of step i+1, the kernels of step i and D2H of
        # step i-1 overlap -- the way an iterative-reconstruction loop would drive several slices.
        h_in = [torch.from_numpy(im).pin_memory() for im in imgs]
        h_outs = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(ROT)]
        in_np = [t.numpy() for t in h_in]
        out_np = [t.numpy() for t in h_outs]

        def e2e_step(i):
            W = Ws[i % ROT]
            W.set_image(in_np[i % ROT])                # H2D (pinned), wt.cu:427
            W.forward()
            W.inverse()
            W.get_image(out_np[i % ROT])               # D2H (pinned), wt.cu:421

        def timed_e2e():
            barrier(world)
            e0.record(stream)
            t0 = time.perf_counter()
            for i in range(args.steps):
                e2e_step(i)
            e1.record(stream)
            barrier(world)
            wall = (time.perf_counter() - t0) * 1e3
            return max_over_ranks(max(e0.elapsed_time(e1), 0.0), world), wall

        for i in range(max(1, args.warmup)):
            e2e_step(i)
        ms_sync, ms_sync_wall = timed_e2e()

        streams = [torch.cuda.Stream() for _ in range(ROT)]
        for W, st in zip(Ws, streams):
            W.set_stream(st)
            W.set_async(True)
        for i in range(max(ROT, args.warmup)):
            e2e_step(i)
        barrier(world)
        e0.record(stream)
        for st in streams:
            st.wait_event(e0)
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(i)
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            stream.wait_event(ev)
        e1.record(stream)
        barrier(world)
        ms_e2e_wall = (time.perf_counter() - t0) * 1e3
        ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), 0.0), world)
        for k in range(ROT):     # every pipelined step really delivered its reconstruction to the host
            err = float(np.abs(out_np[k] - imgs[k]).max() / np.abs(imgs[k]).max())
            if not err < 1e-5:
                raise SystemExit(f"bench.py: pipelined e2e result {k} is wrong (err {err:.3e})")
        for W in Ws:
            W.set_async(False)
            W.set_stream(None)

        # ---- the device-resident steps again, driven over 2 streams (extra information, not `value`): the images are
        # independent, so the small levels of one step (which leave most SMs idle) overlap the level-1 kernels of the
        # next.  Rank-local (no collective inside), reported by rank 0.
        concurrent = None
        if nimg == 1 and ROT >= 2:
            try:
                NSTR = 2
                cstreams = [torch.cuda.Stream() for _ in range(NSTR)]
                for k, W in enumerate(Ws):
                    W.set_stream(cstreams[k % NSTR])
                for i in range(max(ROT, args.warmup)):
                    step(i)
                torch.cuda.synchronize()
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record(stream)
                for st in cstreams:
                    st.wait_event(c0)
                for i in range(args.steps):
                    step(i)
                for st in cstreams:
                    ev = torch.cuda.Event()
                    ev.record(st)
                    stream.wait_event(ev)
                c1.record(stream)
                torch.cuda.synchronize()
                ms_c = c0.elapsed_time(c1) / args.steps
                rec = Ws[(args.steps - 1) % ROT].get_image()
                err = float(np.abs(rec - imgs[(args.steps - 1) % ROT]).max() / np.abs(imgs[(args.steps - 1) % ROT]).max())
                concurrent = {"streams": NSTR, "ms_per_step": round(ms_c, 5), "value": round(npx / ms_c / 1e3, 1),
                              "unit": "Mpixels/s", "reconstruction_err": err,
                              "how": "the same forward()+inverse() steps, objects bound alternately to 2 streams: "
                                     "consecutive (independent) steps overlap; rank-local"}
            except Exception as e:   # extra information only: never fail the bench line over it
                concurrent = {"error": str(e)[:200]}
            finally:
                for W in Ws:
                    W.set_stream(None)
                torch.cuda.synchronize()
    clocks = clk.summary()

    # ---- per-kernel durations (separate pass; the event pairs perturb back-to-back launches slightly)
    L.pdwt_profile_begin()
    for i in range(args.steps):
        step(i)
    ents = (pdwt_b200.ProfileEntry * 64)()
    n = L.pdwt_profile_end(ents, 64)
    kernels = {}
    for k in range(max(n, 0)):
        e = ents[k]
        kernels[e.name.decode()] = {"launches": e.launches, "avg_us": 1e3 * e.ms_total / e.launches,
                                    "min_us": 1e3 * e.ms_min, "total_ms": e.ms_total}

    # ---- the level-1 kernels on their own: K consecutive launches of ONE kernel between two events.  The per-launch
    # event pairs above put an event record in front of every kernel: that adds 2-5 us of launch gap to each and breaks
    # the programmatic-dependent-launch chain.  Two variants: PDWT_PDL=0 (plain stream order: launch n+1 starts when
    # launch n has completed -- the kernel's own duration, used for `roofline`) and the library default (the next
    # launch's prologue overlaps the tail of the previous one -- what a step sees).
    def back_to_back(pdl_off):
        if pdl_off:
            os.environ["PDWT_PDL"] = "0"          # read by the library at every launch
        try:
            W1 = [Wavelets(torch.from_numpy(im).cuda(), wname, 1) for im in imgs]
            for W in W1:
                W.forward()
            K = max(20, args.steps)
            out = {}
            torch.cuda.synchronize()
            e0.record(stream)
            for i in range(K):
                W1[i % ROT].forward()
            e1.record(stream)
            torch.cuda.synchronize()
            out["fwd_level1_us"] = round(1e3 * e0.elapsed_time(e1) / K, 2)
            fh = C.c_void_p()
            if L.pdwt_filters_create(C.byref(fh), wname.encode(), 0) > 0:
                sets = []
                for W in W1:
                    cp = (C.c_void_p * W.ncoeffs)(*[W.coeff_int_ptr(k) for k in range(W.ncoeffs)])
                    sets.append((C.c_void_p(W.image_int_ptr()), cp, C.c_void_p(L.pdwt_wavelets_tmp_int_ptr(W._h)), W.info))
                for img_p, cp, tmp_p, info in sets:
                    L.pdwt_inverse_separable(fh, img_p, cp, tmp_p, info, nimg, None)
                torch.cuda.synchronize()
                e0.record(stream)
                for i in range(K):
                    img_p, cp, tmp_p, info = sets[i % ROT]
                    L.pdwt_inverse_separable(fh, img_p, cp, tmp_p, info, nimg, None)
                e1.record(stream)
                torch.cuda.synchronize()
                out["inv_level1_us"] = round(1e3 * e0.elapsed_time(e1) / K, 2)
                L.pdwt_filters_destroy(fh)
            out["launches_each"] = K
            return out
        finally:
            if pdl_off:
                del os.environ["PDWT_PDL"]

    b2b = back_to_back(True)
    b2b_pdl = back_to_back(False)
    peak, peak_src = peaks()
    roof = None
    if kernels:
        top = max(kernels, key=lambda k: kernels[k]["total_ms"])
        ab = algorithmic_bytes(top)
        if ab:
            ab *= nimg
            key = "inv_level1_us" if "inv" in top else "fwd_level1_us"
            dur_us = kernels[top]["avg_us"]      # one CUDA-event pair around every launch of it inside the steps
            ach = ab / (dur_us * 1e-6) / 1e9
            roof = {"bound": "hbm", "kernel": top, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": None, "algorithmic_bytes_per_launch": ab,
                    "avg_us": round(dur_us, 2), "peak_source": peak_src,
                    "how": "average of one CUDA-event pair around every launch of this kernel inside the steps, on the "
                           "launching stream (the event record in front of the launch adds part of the launch gap)",
                    "kernel_share_of_step": round(kernels[top]["total_ms"] / sum(v["total_ms"] for v in kernels.values()), 3)}
            if key in b2b and top.endswith(f"[{Nr}x{Nc}]"):   # the same kernel launched K times in a row, two events
                for name, d, how in (("back_to_back", b2b, "plain stream order (PDWT_PDL=0): launch n+1 starts when n is complete"),
                                     ("back_to_back_pdl", b2b_pdl, "library default: the next launch's prologue overlaps the tail")):
                    roof[name] = {"avg_us": d[key], "frac": round(ab / (d[key] * 1e-6) / 1e9 / peak, 4),
                                  "how": f"{d['launches_each']} consecutive launches of this kernel alone between two CUDA "
                                         f"events, rotating buffers, {how}"}
            # the second ceiling of this kernel (the separable DWT sits at the FP32 ridge): one level over R x C
            # pixels does hlen FMAs per pixel in each of its two passes = 4 * hlen flop per pixel, on the CUDA cores
            hlen = int(Ws[0].info.hlen)
            fl = 4.0 * hlen * (ab / 8.0)
            tf = fp32_peak_tflops()
            roof["fp32"] = {"bound": "fp32 CUDA cores (no tensor cores: north_star)", "achieved": round(fl / (dur_us * 1e-6) / 1e12, 2),
                            "peak": round(tf, 1), "unit": "TFLOP/s", "frac": round(fl / (dur_us * 1e-6) / 1e12 / tf, 4),
                            "algorithmic_flops_per_launch": fl,
                            "note": "the level kernels are co-limited by FP32 issue and HBM (DESIGN 3.4-3.6): packed FFMA2 is "
                                    "about half of their instruction stream, ncu shows the FMA pipe 50-63 % busy"}
            tr = os.path.join(ROOT, "profiles", "traffic.json")      # dram bytes per launch from the ncu capture
            if os.path.exists(tr) and args.workload == "c2":
                roof["traffic"] = json.load(open(tr)).get(top.split("[")[0])

    value = world * npx * args.steps / (ms_dev * 1e-3) / 1e6
    e2e_val = world * npx * args.steps / (ms_e2e * 1e-3) / 1e6
    whole = 16.0 * npx * args.steps / (ms_dev * 1e-3) / 1e9   # algorithmic GB/s of the whole fwd+inv step, per GPU

    out = {
        "metric": METRIC, "value": round(value, 1), "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_dev / args.steps, 5), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(args.workload),
        "e2e": {"value": round(e2e_val, 1), "unit": "Mpixels/s", "h2d_bytes_per_step": 4 * npx,
                "d2h_bytes_per_step": 4 * npx, "ms_per_step": round(ms_e2e / args.steps, 4),
                "wall_ms_per_step": round(ms_e2e_wall / args.steps, 4),
                "api": f"Wavelets.set_image(pinned host) -> forward -> inverse -> get_image(pinned host), {ROT} objects "
                       f"on {ROT} streams with set_async(True): copies and kernels of consecutive steps overlap",
                "sync": {"value": round(world * npx * args.steps / (ms_sync * 1e-3) / 1e6, 1),
                         "ms_per_step": round(ms_sync / args.steps, 4),
                         "api": "same four calls, blocking copies on one stream (the reference's semantics)"}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "step_algorithmic_gbs": {"achieved": round(whole, 1), "frac_of_peak": round(whole / peak, 4),
                                 "bytes_per_pixel": 16},
        "kernels": {k: {"launches": v["launches"], "avg_us": round(v["avg_us"], 2)} for k, v in kernels.items()},
        "level1_back_to_back": {"plain_stream_order": b2b, "pdl": b2b_pdl},
    }
    if concurrent is not None:
        out["concurrent_streams"] = concurrent
    if rank == 0 and world == 1 and args.workload == "c2":
        # north_star's target configuration ("batched 4096x4096 ... >= 70 % of HBM peak on 1 GPU"): the same transform
        # on 8 images held by ONE batched object, device-resident, same timing rules (inputs 512 MiB > L2)
        out["batched_4096x8"] = batched(L, peak, 4096, 4096, wname, levels, 8,
                                        "8 x 4096x4096 float32 in one batched Wavelets object (north_star's target configuration)")
    if rank == 0 and world == 1 and args.workload == "c2":
        # the other BASELINE.json configurations, each with its own roofline and the reference's CUDA build beside it
        out["configs"] = baseline_configs(L, peak)
    if world > 1 and args.workload == "c2":
        # BASELINE.json configs[4]: the batch sharded over the ranks THROUGH ShardedWavelets / Layer C over NCCL
        sh = sharded_c5(world, rank, peak)
        if rank == 0:
            out["c5_sharded"] = sh
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_port_baseline(Nr, Nc, wname, levels, iters=3)
    if args.workload != "c2":
        out["metric"] = f"Mpixels/s fwd+inv 2D DWT db7 L3 {Nr}^2 x {nimg} per GPU"
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def batched(L, peak, Nr, Nc, wname, levels, nimg, what, steps=10, warmup=3):
    import torch
    import pdwt_b200
    from pdwt_b200 import Wavelets
    x = torch.randn((nimg, Nr, Nc), device="cuda") * 50 + 128
    W = Wavelets(x, wname, levels)
    for _ in range(warmup):
        W.forward(); W.inverse()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        W.forward(); W.inverse()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    L.pdwt_profile_begin()
    for _ in range(steps):
        W.forward(); W.inverse()
    ents = (pdwt_b200.ProfileEntry * 64)()
    n = L.pdwt_profile_end(ents, 64)
    ks = {ents[k].name.decode(): 1e3 * ents[k].ms_total / ents[k].launches for k in range(max(n, 0))}
    npx = nimg * Nr * Nc
    res = {"workload": f"{what}, {wname}, {levels} levels, forward()+inverse()",
           "steps": steps, "warmup": warmup, "ms_per_step": round(ms, 4), "value": round(npx / (ms * 1e-3) / 1e6, 1),
           "unit": "Mpixels/s", "step_algorithmic_gbs": round(16.0 * npx / (ms * 1e-3) / 1e9, 1),
           "step_frac_of_peak": round(16.0 * npx / (ms * 1e-3) / 1e9 / peak, 4)}
    lvl1 = {k: v for k, v in ks.items() if k.endswith(f"[{Nr}x{Nc}]")}
    if lvl1:
        ab = 8.0 * npx
        res["level1_kernels"] = {k: {"avg_us": round(v, 2), "achieved": round(ab / (v * 1e-6) / 1e9, 1),
                                     "frac": round(ab / (v * 1e-6) / 1e9 / peak, 4)} for k, v in lvl1.items()}
        top = max(lvl1, key=lvl1.get)
        res["roofline"] = {"bound": "hbm", "kernel": top, "achieved": res["level1_kernels"][top]["achieved"], "peak": peak,
                           "unit": "GB/s", "frac": res["level1_kernels"][top]["frac"], "algorithmic_bytes_per_launch": ab}
    del W, x
    torch.cuda.empty_cache()
    return res


def fp32_peak_tflops():
    """148 SMs x 128 FP32 lanes x 2 flop x the maximum SM clock (MEASURED_PEAKS.json sm_max_mhz, else 1965 MHz)"""
    mhz = 1965.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        mhz = float(json.load(open(p)).get("sm_max_mhz", mhz))
    return 148 * 128 * 2 * mhz * 1e6 / 1e12


def _timed_us(fn, iters, warm=3):
    import torch
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters


def _load_ref():
    so = os.path.join(ROOT, "oracle", "_ref", "libpdwt_ref.so")
    if not os.path.exists(so):
        return None
    try:
        R = C.CDLL(so)
    except OSError:
        return None
    fp = C.POINTER(C.c_float)
    R.ref_create.restype = C.c_void_p
    R.ref_create.argtypes = [fp, C.c_int, C.c_int, C.c_char_p] + [C.c_int] * 6
    for n in ("ref_forward", "ref_inverse", "ref_destroy"):
        getattr(R, n).argtypes = [C.c_void_p]
    R.ref_soft_threshold.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int]
    R.ref_norm1.argtypes = [C.c_void_p]
    R.ref_norm1.restype = C.c_float
    return R


def baseline_configs(L, peak):
    """BASELINE.json configs[2], [3] and (per GPU) [4], device-resident, CUDA events on the launching stream; the
    reference's own CUDA build (oracle/_ref, same GPU, same inputs) beside each where it has a counterpart.
    Algorithmic bytes and flops per pixel: SURVEY section 8d / DESIGN.md section 3.4."""
    import torch
    import pdwt_b200
    from pdwt_b200 import Wavelets
    R = _load_ref()
    fp = C.POINTER(C.c_float)
    tf = fp32_peak_tflops()
    res = {}

    def one(tag, shape, wname, levels, sep, swt, seq, iters, b_px, flop_px, what):
        x = seeded_image(shape, 0)
        npx = x.size
        W = Wavelets(torch.from_numpy(x).cuda(), wname, levels, do_separable=sep, do_swt=swt)
        l0 = L.pdwt_launch_count()
        us = _timed_us(lambda: seq(W, None), iters)
        launches = (L.pdwt_launch_count() - l0) // (iters + 3)
        L.pdwt_profile_begin()
        for _ in range(iters):
            seq(W, None)
        ents = (pdwt_b200.ProfileEntry * 64)()
        n = L.pdwt_profile_end(ents, 64)
        ks = {ents[k].name.decode(): round(1e3 * ents[k].ms_total / ents[k].launches * (ents[k].launches / iters), 2)
              for k in range(max(n, 0))}
        gbs, tfl = b_px * npx / us / 1e3, flop_px * npx / us / 1e6
        d = {"workload": what, "steps": iters, "warmup": 3, "us_per_step": round(us, 1),
             "value": round(npx / us, 1), "unit": "Mpixels/s", "gpu_launches_per_step": int(launches),
             "roofline": {"hbm": {"bound": "hbm", "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s",
                                  "frac": round(gbs / peak, 4), "algorithmic_bytes_per_step": b_px * npx},
                          "fp32": {"bound": "fp32 CUDA cores (no tensor cores: north_star)", "achieved": round(tfl, 2),
                                   "peak": round(tf, 1), "unit": "TFLOP/s", "frac": round(tfl / tf, 4),
                                   "algorithmic_flops_per_step": flop_px * npx}},
             "us_per_step_by_kernel": ks}
        d["roofline"]["binding"] = "fp32" if d["roofline"]["fp32"]["frac"] > d["roofline"]["hbm"]["frac"] else "hbm"
        if R is not None:
            h = R.ref_create(x.ctypes.data_as(fp), shape[0], shape[1], wname.encode(), levels, 1, sep, 0, swt, 2)
            us_r = _timed_us(lambda: seq(None, h), max(3, iters // 2))
            R.ref_destroy(h)
            d["reference_cuda_us_per_step"] = round(us_r, 1)
            d["vs_reference_cuda"] = round(us_r / us, 2)
        del W
        torch.cuda.empty_cache()
        res[tag] = d

    def seq_fwd_inv(W, h):
        if W is not None:
            W.forward(); W.inverse()
        else:
            R.ref_forward(h); R.ref_inverse(h)

    def seq_c4(W, h):   # README.md:90-103
        if W is not None:
            W.forward(); W.norm1(); W.soft_threshold(10.0, 0, 0); W.norm1(); W.inverse()
        else:
            R.ref_forward(h); R.ref_norm1(h); R.ref_soft_threshold(h, 10.0, 0, 0); R.ref_norm1(h); R.ref_inverse(h)

    try:
        one("c3", (2048, 2048), "sym8", 4, 1, 1, seq_fwd_inv, 10, 112.0, 1536.0,
            "configs[2]: 2-D SWT sym8, 4 levels, 2048x2048 float32, forward()+inverse()")
        one("c4", (4096, 4096), "db7", 2, 0, 0, seq_c4, 6, 31.5, 980.0,
            "configs[3]: non-separable 2-D DWT db7, 2 levels, 4096x4096 float32, forward -> norm1 -> "
            "soft_threshold(10) -> norm1 -> inverse (README.md:90-103)")
        res["c5_per_gpu"] = batched(L, peak, 2048, 2048, "db7", 3, 64,
                                    "configs[4] per GPU: 64 of the 512 images of 2048x2048 in one batched object")
    except Exception as e:   # extra blocks never take the headline line down
        res["error"] = str(e)[:300]
    return res


def sharded_c5(world, rank, peak):
    """BASELINE.json configs[4] at N GPUs: 64 x N images of 2048^2 (512 at N = 8), db7, 3 levels, held on rank 0's device,
    scattered to the owners, transformed there, gathered back -- all through pdwt_b200.sharded.ShardedWavelets, i.e.
    Layer C of the C ABI over NCCL (device to device).  Scatter, compute and gather are timed separately with CUDA
    events (max over ranks); nothing here touches host memory."""
    import torch
    from pdwt_b200.sharded import ShardedWavelets
    try:
        Nr = Nc = 2048
        per = 64
        B = per * world
        g = torch.Generator(device="cuda").manual_seed(11)
        full = (torch.randn((B, Nr, Nc), device="cuda", generator=g) * 50 + 128) if rank == 0 else None
        S = ShardedWavelets(full, "db7", 3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def timed(fn, reps):
            barrier(world)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            barrier(world)
            return max_over_ranks(e0.elapsed_time(e1), world) / reps

        for _ in range(2):
            S.forward(); S.inverse()
        ms_scatter = timed(lambda: S.scatter_batch(full), 3)
        ms_compute = timed(lambda: (S.forward(), S.inverse()), 10)
        dst = torch.empty_like(full) if rank == 0 else None
        S.gather_image(out=dst)           # first use of the owners -> root direction sets the NCCL connections up
        ms_gather = timed(lambda: S.gather_image(out=dst), 3)
        rec = [dst]
        err = None
        if rank == 0:
            err = float(((rec[0] - full).abs().max() / full.abs().max()).item())
        S.forward()
        n1 = S.norm1()
        S.close()
        npx = B * Nr * Nc
        moved = 4.0 * (B - per) * Nr * Nc      # bytes that cross NVLink per direction (the root keeps its own block)
        out = {"workload": f"configs[4]: {B} images of {Nr}x{Nc} float32, db7, 3 levels, {per} per GPU on {world} GPUs, "
                           f"ShardedWavelets (Layer C: grouped ncclSend/ncclRecv, device to device)",
               "scatter_ms": round(ms_scatter, 3), "compute_ms_per_step": round(ms_compute, 4), "gather_ms": round(ms_gather, 3),
               "compute_only": {"value": round(npx / ms_compute / 1e3, 1), "unit": "Mpixels/s",
                                "per_gpu_step_algorithmic_gbs": round(16.0 * per * Nr * Nc / ms_compute / 1e6, 1),
                                "per_gpu_step_frac_of_peak": round(16.0 * per * Nr * Nc / ms_compute / 1e6 / peak, 4)},
               "scatter_compute_gather": {"value": round(npx / (ms_scatter + ms_compute + ms_gather) / 1e3, 1),
                                          "unit": "Mpixels/s"},
               "root_link_gbs": {"scatter": round(moved / ms_scatter / 1e6, 1), "gather": round(moved / ms_gather / 1e6, 1),
                                 "nominal_nvlink5_per_direction": 900.0},
               "reconstruction_err": err, "norm1_of_first_and_last_image": [float(n1[0]), float(n1[-1])],
               "nccl": "bound at run time by libpdwt_b200.so (dlopen libnccl.so.2)"}
        del S, full
        torch.cuda.empty_cache()
        return out
    except Exception as e:
        return {"error": str(e)[:300]}


def cpu_port_baseline(Nr, Nc, wname, levels, iters):
    """the CPU oracle (oracle/pdwt_oracle.c, OpenMP) on this box's host cores: `iters` fwd+inv of the same image"""
    import oracle
    x = seeded_image((Nr, Nc), 0)
    O = oracle.Wavelets(x, wname, levels)
    O.forward(); O.inverse()        # warm-up (page faults, OpenMP pool)
    t = []
    for _ in range(iters):
        O.set_image(x)
        t0 = time.perf_counter()
        O.forward(); O.inverse()
        t.append(time.perf_counter() - t0)
    s = float(np.median(t))
    return {"value": round(Nr * Nc / s / 1e6, 1), "unit": "Mpixels/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{iters} x (forward+inverse) of one {Nr}x{Nc} image, median; OpenMP over all host cores"}


# ====================================================================================== reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    Nr, Nc, wname, levels, nimg = WORKLOADS[args.workload]
    if nimg != 1:
        raise SystemExit("--impl reference: the reference has no batch dimension; use c2 or c5img")
    npx = Nr * Nc
    cfg = config_of(args.workload)
    base = {"impl": "reference", "metric": METRIC, "unit": "Mpixels/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg}
    so = os.path.join(ROOT, "oracle", "_ref", "libpdwt_ref.so")
    R = None
    try:
        import torch
        R = C.CDLL(so) if torch.cuda.is_available() else None
        if R is not None:
            torch.cuda.set_device(0)
    except OSError:
        R = None
    if R is None:
        # the reference's CUDA build is not usable here: time the CPU port of its kernels instead
        cb = cpu_port_baseline(Nr, Nc, wname, levels, iters=max(1, min(args.steps, 5)))
        base.update({"value": cb["value"], "ms_per_step": round(npx / cb["value"] / 1e3, 3), "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(base), flush=True)
        return
    fp = C.POINTER(C.c_float)
    R.ref_create.restype = C.c_void_p
    R.ref_create.argtypes = [fp, C.c_int, C.c_int, C.c_char_p] + [C.c_int] * 6
    R.ref_destroy.argtypes = [C.c_void_p]
    R.ref_time_rotating.restype = C.c_float
    R.ref_time_rotating.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]
    R.ref_time_e2e.restype = C.c_float
    R.ref_time_e2e.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.POINTER(fp), fp]
    R.ref_get_image.argtypes = [C.c_void_p, fp]
    import torch
    imgs = [seeded_image((Nr, Nc), i) for i in range(ROTATE)]
    hs = (C.c_void_p * ROTATE)(*[R.ref_create(im.ctypes.data_as(fp), Nr, Nc, wname.encode(), levels, 1, 1, 0, 0, 2)
                                 for im in imgs])
    with ClockSampler(0) as clk:
        R.ref_time_rotating(hs, ROTATE, 0, args.warmup)
        ms = R.ref_time_rotating(hs, ROTATE, args.warmup, args.steps)
        h_in = [torch.from_numpy(im).pin_memory() for im in imgs]
        h_out = torch.empty((Nr, Nc), dtype=torch.float32).pin_memory()
        ins = (fp * ROTATE)(*[C.cast(t.data_ptr(), fp) for t in h_in])
        outp = C.cast(h_out.data_ptr(), fp)
        R.ref_time_e2e(hs, ROTATE, 0, max(1, args.warmup), ins, outp)
        ms_e2e = R.ref_time_e2e(hs, ROTATE, args.warmup, args.steps, ins, outp)
    rec = h_out.numpy()
    src = imgs[(args.warmup + args.steps - 1) % ROTATE]
    err = float(np.abs(rec - src).max() / np.abs(src).max())
    for h in hs:
        R.ref_destroy(h)
    v = npx * args.steps / (ms * 1e-3) / 1e6
    ve = npx * args.steps / (ms_e2e * 1e-3) / 1e6
    base.update({
        "value": round(v, 1), "ms_per_step": round(ms / args.steps, 5),
        "e2e": {"value": round(ve, 1), "unit": "Mpixels/s", "h2d_bytes_per_step": 4 * npx, "d2h_bytes_per_step": 4 * npx,
                "ms_per_step": round(ms_e2e / args.steps, 4),
                "api": "reference Wavelets::set_image(host) -> forward -> inverse -> get_image(host)"},
        "cpu_baseline": {"value": round(v, 1), "unit": "Mpixels/s", "cores": 0, "kind": "reference",
                         "sample": "PDWT has no CPU implementation: this is its stock CUDA code (unmodified sources, "
                                   "nvcc -arch=sm_100, oracle/_ref/libpdwt_ref.so) on the same B200 through the "
                                   "reference's own Wavelets class; every step, rotating over 4 objects"},
        "clocks": clk.summary(), "reconstruction_err": err,
        "config": cfg,
    })
    print(json.dumps(base), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
