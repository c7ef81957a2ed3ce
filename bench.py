#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline: "GPR train+predict sec at n=50k; Cholesky FP64 TFLOP/s vs peak; test pts/s".

A step = one full pass of the hot path on BASELINE config 4: squared-exponential GPR,
d = 8, n = 50 000 training points, noise 0.01 -> kernel-matrix build, Cholesky, alpha + logp, then predictive
mean + variance of m = 1 000 000 test points.  With N GPUs the test points are split into N contiguous shards, one per
rank, and every rank factorises its own replica (no data-path collective; total work fixed => "strong" scaling).

  value   seconds per step with X, y, X_star already resident in HBM (device timed, max over ranks)
  e2e     the same through the host API (GPR(X, y, noise, k); $predict(X_star)) with pinned HOST buffers: H2D of
          X, y, X_star and D2H of mean/var inside the timed region
  roofline      the dominant kernel (variance pass: V = L^-1 K_star, fused column norms) against the FP64 tensor
                peak measured live with cuBLAS Dgemm (MEASURED_PEAKS.json has no FP64 figure).  With the INT8
                tensor-core pass (the default at this size) the fraction exceeds 1; roofline.int8 states the same
                time against the INT8 pipe (2 x the measured bf16 rate)
  parity        the first 4096 test points of the timed result re-predicted through the FP64 substitution on a
                fresh factor; asserted within the north-star tolerance (1e-9)
  cpu_baseline  the oracle (NumPy/SciPy/OpenBLAS restatement of the reference's R path) on the host cores, on a
                bounded sample scaled by flop count to the full job (R is not installed: kind = "port")

`--impl reference` times only that CPU path and prints the same JSON line with "impl": "reference".
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GPR train+predict sec at n=50k; Cholesky FP64 TFLOP/s vs peak; test pts/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--train-size", dest="n", type=int, default=50000)
    ap.add_argument("--test-size", dest="m", type=int, default=1000000)
    ap.add_argument("--dim", dest="d", type=int, default=8)
    ap.add_argument("--cpu-sample-n", type=int, default=8192)
    ap.add_argument("--cpu-sample-m", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--predict-path", type=int, default=0, choices=[0, 1, 2, 3, 4],
                    help="variance pass (GPRC_OPT_PREDICT_PATH): 0 auto, 1 inverse, 2 FP64 substitution, 3 persistent "
                         "FP64 substitution, 4 substitution on the INT8 tensor cores")
    ap.add_argument("--ozaki-digits", type=int, default=7, choices=[6, 7, 8])
    ap.add_argument("--int8-tile", type=int, default=64, choices=[64, 128],
                    help="test points per CTA of the INT8 pass (GPRC_OPT_INT8_TILE)")
    ap.add_argument("--parity-sample", type=int, default=4096,
                    help="test points re-predicted through the FP64 substitution (path 2) after the timed run and "
                         "compared with the timed result")
    ap.add_argument("--train", default="auto", choices=["auto", "replicated", "distributed"],
                    help="N > 1: every rank factorises its own replica, or the factorisation itself is distributed "
                         "(panel-cyclic, NCCL broadcasts) and assembled on every rank; auto = distributed")
    return ap.parse_args()


def make_inputs(n, m, d):
    """BASELINE config 4 generator (identical bits on every rank and in the CPU arm)."""
    rng = np.random.default_rng(4)
    X = rng.uniform(-1, 1, size=(d, n))
    y = np.sum(np.sin(math.pi * X), axis=0) + rng.normal(0, 0.1, n)
    Xs = rng.uniform(-1, 1, size=(d, m))
    return X, y, Xs


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle, bounded sample, scaled by algorithmic flops
# ---------------------------------------------------------------------------------------------------------------
def cpu_reference(n, m, d, ns, ms, steps=1, warmup=0):
    import scipy.linalg
    from oracle import gprc_oracle as o
    # all the host threads the box offers (torchrun exports OMP_NUM_THREADS=1 to its workers: undo that here)
    threads = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits, threadpool_info
        threadpool_limits(limits=threads)
        threads = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        pass
    X, y, Xs = make_inputs(ns, ms, d)
    k = o.cov_func(o.sqrexp, l=1.0)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        K = o.covariance_matrix(X, X, k)                              # R/GPRclass.R:138
        t1 = time.perf_counter()
        L = scipy.linalg.cholesky(K + 0.01 * np.eye(ns), lower=True)  # :142
        t2 = time.perf_counter()
        alpha = scipy.linalg.solve_triangular(L.T, scipy.linalg.solve_triangular(L, y, lower=True), lower=False)
        logp = -0.5 * (y @ alpha) - np.sum(np.log(np.diag(L))) - ns / 2 * math.log(2 * math.pi)
        t3 = time.perf_counter()
        Ks = o.covariance_matrix(X, Xs, k)                            # :160
        mean = Ks.T @ alpha
        t4 = time.perf_counter()
        v = scipy.linalg.solve_triangular(L, Ks, lower=True)          # :162 ("generous": substitution, not dgesv)
        var = k(Xs, Xs) - np.sum(v * v, axis=0)
        t5 = time.perf_counter()
        if it >= warmup:
            times.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4))
    tb, tc, ts, tk, tv = np.mean(np.array(times), axis=0)
    # scale each phase by its algorithmic work
    full = (tb * (n / ns) ** 2 + tc * (n / ns) ** 3 + ts * (n / ns) ** 2 + tk * (n * m) / (ns * ms)
            + tv * (n * n * m) / (ns * ns * ms))
    sample = ("oracle (NumPy/SciPy, OpenBLAS) on n=%d, m=%d of the same generator; phases build/chol/solve/Kstar/trsm "
              "= %.2f/%.2f/%.2f/%.2f/%.2f s, each scaled by its algorithmic work to n=%d, m=%d (scaled, not run)"
              % (ns, ms, tb, tc, ts, tk, tv, n, m))
    return dict(value=full, unit="s", cores=threads, kind="port", sample=sample,
                sample_seconds=float(tb + tc + ts + tk + tv),
                chol_gflops=ns ** 3 / 3 / tc / 1e9, trsm_gflops=ns * ns * ms / tv / 1e9)


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "500"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # samples under load: power above 40 % of the observed maximum
        if pw:
            thr = 0.4 * max(pw)
            load = [s for s, p in zip(sm, pw) if p >= thr] or sm
        else:
            load = sm
        return dict(sm_mhz=float(np.median(load)) if load else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=sorted(reasons))


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def fp64_tensor_peak_tflops(torch, dev):
    """cuBLAS Dgemm 8192^3, best of 5 (burst) -- the denominator of the tensor roofline, measured in this run."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty(n, n, dtype=torch.float64, device=dev)
    torch.mm(a, b, out=c)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.mm(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    torch.cuda.empty_cache()
    return 2 * n ** 3 / (best * 1e-3) / 1e12


def roofline_block(args, used_path, timers, achieved, peak_tf, flops_var, var_ms, sec_per_step, n):
    """The `roofline` object of the JSON line (pure function: unit-tested on the CPU tier)."""
    kname = ("gemm_kernel<TrsmLeftUpdatePolicy> (variance pass v = L^-1 K_star by blocked substitution, K_star^T "
             "updated in place; column norms fused into gemm_kernel<TrsmLeftDiagPolicy>)") if not timers["trtri"] \
        else "gemm_kernel<TrmmNormPolicy> (variance pass v = (L^-1) K_star, fused column norms)"
    int8 = None
    if used_path == 4:
        # the O(n^2 m) products run as S(S+1)/2 exact INT8 digit products per FP64 product (ozaki.cuh): the kernel's
        # own roofline is the INT8 tensor pipe.  MEASURED_PEAKS.json has no INT8 figure; INT8 dense is nominally
        # 2 x bf16 dense on B200 (4.5 vs 2.25 POP/s), so the denominator is 2 x the measured bf16 figure.
        S = args.ozaki_digits
        pairs = S * (S + 1) // 2
        bf16 = None
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")) as f:
                bf16 = float(json.load(f)["bf16_tflops_sustained"])
            src = "2 x bf16_tflops_sustained of MEASURED_PEAKS.json (of measured; INT8 dense = 2 x bf16 dense nominally)"
        except Exception:
            bf16, src = 1400.0, "2 x 1.4 PFLOP/s sustained bf16 (of fallback, B200_PROFILING.md)"
        kern = ("oz::update_kernel<%d>" if args.int8_tile == 64 else "oz::update128_kernel<%d>") % S
        int8 = dict(kernel=kern + " (tcgen05.mma kind::i8, TMEM accumulators)", digits=S, tile=args.int8_tile,
                    int8_products_per_fp64_product=pairs,
                    achieved_int8_tops=(achieved * pairs) if achieved else None, peak_int8_tops=2 * bf16,
                    frac_of_int8_peak=(achieved * pairs / (2 * bf16)) if achieved else None, peak_source=src)
        kname = (kern + " (variance pass v = L^-1 K_star by blocked substitution with the O(n^2 m) "
                 "products as exact INT8 digit products on tcgen05 / TMEM; FP64 diagonal solves and column norms in "
                 "gemm_kernel<TrsmLeftDiagPolicy>)")
    fp64_view = dict(achieved=achieved, peak=peak_tf, unit="TFLOP/s", frac=(achieved / peak_tf) if achieved else None,
                     peak_source="cuBLAS Dgemm fp64 8192^3 best of 5, measured in this run (MEASURED_PEAKS.json "
                                 "holds no FP64 figure; of measured)")
    common = dict(kernel=kname, bound="tensor", predict_path=used_path, algorithmic_flops_per_step=flops_var,
                  ms_per_step=var_ms, share_of_step=var_ms / (sec_per_step * 1e3))
    if used_path == 4:
        # The bounding unit of this kernel is the INT8 tensor pipe: algorithmic work = S(S+1)/2 INT8 multiply-adds
        # per FP64 multiply-add of the n^2 m pass, peak = 2 x the measured bf16 rate.
        # `fp64_equivalent` states the same time against the FP64 tensor roofline (cuBLAS Dgemm): a ratio above 1
        # there is the point of the INT8 path, not a measurement artefact.
        roofline = dict(common, achieved=int8["achieved_int8_tops"], peak=int8["peak_int8_tops"], unit="TFLOP/s",
                        unit_note="INT8 tensor-core operations per second / 1e12 (TOP/s; 2 per multiply-add)",
                        frac=int8["frac_of_int8_peak"], peak_source=int8["peak_source"],
                        algorithmic_int8_ops_per_step=flops_var * int8["int8_products_per_fp64_product"],
                        int8=int8, fp64_equivalent=fp64_view,
                        traffic=(1.1167e9 if args.int8_tile == 64 else None),
                        traffic_note="dram bytes of ONE captured launch of oz::update_kernel<7> (ncu --set full, "
                                     "K = 16 256, 9472 test points; algorithmic bytes of that launch 1.10e9), not "
                                     "per step; profiles/README.md")
    else:
        roofline = dict(common, **fp64_view,
                        # one ncu --set full capture of a mid-sweep launch of this kernel (profiles/README.md):
                        # dram__bytes_read + write = 4.95e9 B for a launch whose algorithmic bytes (V rows read
                        # once, L row panel) are 3.8e9 B
                        traffic=(4.946e9 if not timers["trtri"] and n == 50000 else None),
                        traffic_note="bytes of ONE captured launch (grid 148, block row ~197 of 392, 4.37 ms), "
                                     "not per step")
    return roofline


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n, m, d = args.n, args.m, args.d
    config = dict(workload="C4: GPR sqrexp(l=1)+iid noise 0.01, d=%d, n=%d train, m=%d test (mean+variance)" % (d, n, m),
                  n=n, m=m, d=d, sharding="test points split across ranks, replicated factorisation",
                  l2="inputs larger than L2 (L is %.1f GB)" % (8.0 * n * n / 1e9))

    if args.impl == "reference":
        if rank != 0:
            return 0
        ref = cpu_reference(n, m, d, args.cpu_sample_n, args.cpu_sample_m, steps=max(1, args.steps),
                            warmup=min(args.warmup, 1))
        line = dict(metric=METRIC, value=ref["value"], unit="s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ref["value"] * 1e3, higher_is_better=False, scaling="strong", vs_baseline=None,
                    dtype="f64", data="synthetic", config=config, impl="reference",
                    cpu_baseline=dict(value=ref["value"], unit="s", cores=ref["cores"], kind=ref["kind"],
                                      sample=ref["sample"]),
                    e2e=dict(value=ref["value"], unit="s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    test_pts_per_s=m / ref["value"], cholesky_tflops=ref["chol_gflops"] / 1e3)
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    os.environ.setdefault("GPRC_DEVICE", str(local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import gprc_b200 as g
    ctx = g.Context(local)
    lib = ctx.lib
    ctx.set_option(g._lib.OPT_PREDICT_PATH, args.predict_path)
    ctx.set_option(g._lib.OPT_OZAKI_DIGITS, args.ozaki_digits)
    ctx.set_option(g._lib.OPT_INT8_TILE, args.int8_tile)
    dist_train = world > 1 and args.train in ("auto", "distributed")
    D = None
    if dist_train:
        from importlib import import_module
        D = import_module("gaussian-process-regression_b200.dist").DistGPR(ctx)
        config["sharding"] = "factorisation distributed by 512-column panels (NCCL broadcast), factor assembled on every rank; test points split across ranks"
    dist_phase = dict(build=0.0, factor=0.0, solve=0.0)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak_tf = fp64_tensor_peak_tflops(torch, dev)

    X, y, Xs_all = make_inputs(n, m, d)
    lo, hi = rank * m // world, (rank + 1) * m // world
    m_local = hi - lo
    xp = np.ascontiguousarray(X.T)                      # the ABI layout: points contiguous
    xsp = np.ascontiguousarray(Xs_all[:, lo:hi].T)
    del Xs_all
    mean = np.empty(m_local)
    var = np.empty(m_local)
    for a in (xp, y, xsp, mean, var):                   # pinned host memory for the e2e leg
        lib.gprc_host_register(a.ctypes.data_as(C.c_void_p), a.nbytes)
    dX, dy, dXs = ctx.upload(xp), ctx.upload(y), ctx.upload(xsp)
    dmean, dvar = ctx.malloc(8 * m_local), ctx.malloc(8 * m_local)
    spec = g.KernelSpec("sqrexp", l=1.0)
    kc, _keep = spec.to_c()

    def step_device():
        if dist_train:
            # collective train (X, y are 3.6 MB: uploaded by the call), then predict on this rank's shard
            h, lp, inf, ph = D.fit_replicated(xp, d, n, y, 0.01, spec)
            assert inf == 0, "not positive definite: %d" % inf
            for kk in dist_phase:
                dist_phase[kk] += ph[kk]
            g._lib.check(lib.gprc_gpr_predict_dev(h, dXs, m_local, dmean, dvar))
            ctx.sync()
            lib.gprc_gpr_free(h)
            return lp
        h = C.c_void_p()
        logp, info = C.c_double(0.0), C.c_long(0)
        g._lib.check(lib.gprc_gpr_fit_dev(ctx.handle, kc, dX, d, n, dy, 0.01, C.byref(h), C.byref(logp), C.byref(info)))
        assert info.value == 0, "not positive definite: %d" % info.value
        g._lib.check(lib.gprc_gpr_predict_dev(h, dXs, m_local, dmean, dvar))
        ctx.sync()
        lib.gprc_gpr_free(h)
        return logp.value

    def step_host():
        if dist_train:
            h, lp, inf, ph = D.fit_replicated(xp, d, n, y, 0.01, spec)
            assert inf == 0
            g._lib.check(lib.gprc_gpr_predict(h, g._lib.dptr(xsp), m_local, g._lib.dptr(mean), g._lib.dptr(var)))
            lib.gprc_gpr_free(h)
            return lp
        h = C.c_void_p()
        logp, info = C.c_double(0.0), C.c_long(0)
        g._lib.check(lib.gprc_gpr_fit(ctx.handle, kc, g._lib.dptr(xp), d, n, g._lib.dptr(y), 0.01, C.byref(h),
                                      C.byref(logp), C.byref(info)))
        assert info.value == 0
        g._lib.check(lib.gprc_gpr_predict(h, g._lib.dptr(xsp), m_local, g._lib.dptr(mean), g._lib.dptr(var)))
        lib.gprc_gpr_free(h)
        return logp.value

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    ctx.reset_timers()
    for kk in dist_phase:
        dist_phase[kk] = 0.0
    ctx.mark(0)
    t0 = time.perf_counter()
    logp = 0.0
    for _ in range(args.steps):
        logp = step_device()
    ctx.mark(1)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ctx.elapsed_ms(0, 1)
    timers, launches = ctx.timers()
    used_path = ctx.last_predict_path()
    if dist_train:
        timers["build_k"], timers["chol"], timers["solve"] = dist_phase["build"], dist_phase["factor"], dist_phase["solve"]
    clocks = sampler.stop() if rank == 0 else None
    sec_per_step = max_over_ranks(dev_ms / 1e3 / args.steps)
    wall_per_step = max_over_ranks(wall / args.steps)

    # sanity of the result that was timed (finite, variance within [0, k**])
    v = np.empty(min(m_local, 4096))
    ctx.d2h(v, dvar)
    assert np.all(np.isfinite(v)) and v.min() > -1e-8 and v.max() <= 1.0 + 1e-9, (v.min(), v.max())

    # parity of the timed result at full size: the first test points of this rank's shard are predicted again through
    # the FP64 blocked substitution (path 2) on a fresh factor and compared (mean, variance) with what was timed
    parity = None
    ps = min(args.parity_sample, m_local)
    if ps > 0:
        got_m, got_v = np.empty(ps), np.empty(ps)
        ctx.d2h(got_m, dmean)
        ctx.d2h(got_v, dvar)
        ctx.set_option(g._lib.OPT_PREDICT_PATH, 2)
        if dist_train:
            h = D.fit_replicated(xp, d, n, y, 0.01, spec)[0]
        else:
            h = C.c_void_p()
            lp_, inf_ = C.c_double(0.0), C.c_long(0)
            g._lib.check(lib.gprc_gpr_fit_dev(ctx.handle, kc, dX, d, n, dy, 0.01, C.byref(h), C.byref(lp_), C.byref(inf_)))
        dm2, dv2 = ctx.malloc(8 * ps), ctx.malloc(8 * ps)
        g._lib.check(lib.gprc_gpr_predict_dev(h, dXs, ps, dm2, dv2))
        ctx.sync()
        ref_m, ref_v = np.empty(ps), np.empty(ps)
        ctx.d2h(ref_m, dm2)
        ctx.d2h(ref_v, dv2)
        lib.gprc_gpr_free(h)
        ctx.free(dm2)
        ctx.free(dv2)
        ctx.set_option(g._lib.OPT_PREDICT_PATH, args.predict_path)
        parity = dict(sample=int(ps), against="FP64 blocked substitution (predict path 2) on a fresh factor",
                      max_abs_dvar=float(np.max(np.abs(got_v - ref_v))),
                      max_abs_dmean=float(np.max(np.abs(got_m - ref_m))),
                      max_rel_dvar=float(np.max(np.abs(got_v - ref_v) / np.maximum(np.abs(ref_v), 1e-300))),
                      tolerance="1e-9 relative on mean and variance (BASELINE.json north_star)")
        assert parity["max_abs_dvar"] <= 1e-9 and parity["max_abs_dmean"] <= 1e-9 * max(1.0, float(np.max(np.abs(ref_m)))), parity

    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, 2))  # a step is > 1 min at full size: two are enough for this leg
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host()
        barrier()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        e2e = dict(value=e2e_s, unit="s", h2d_bytes_per_step=int(xp.nbytes + y.nbytes + xsp.nbytes) * world,
                   d2h_bytes_per_step=int(mean.nbytes + var.nbytes) * world,
                   steps=e2e_steps,
                   note="host API gprc_gpr_fit + gprc_gpr_predict with pinned host buffers; bytes summed over ranks")

    if rank == 0:
        var_ms = timers["var"] / args.steps
        flops_var = float(n) * n * m_local            # algorithmic: n^2 m  
        achieved = flops_var / (var_ms * 1e-3) / 1e12 if var_ms > 0 else None
        n_pad = (n + 127) // 128 * 128
        chol_tf = n ** 3 / 3 / (timers["chol"] / args.steps * 1e-3) / 1e12  # aggregate over ranks when distributed
        roofline = roofline_block(args, used_path, timers, achieved, peak_tf, flops_var, var_ms, sec_per_step, n)
        line = dict(metric=METRIC, value=sec_per_step, unit="s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=sec_per_step * 1e3, higher_is_better=False, scaling="strong", vs_baseline=None,
                    dtype="f64", data="synthetic", config=config, clocks=clocks, e2e=e2e, parity=parity,
                    gpu_launches=int(launches), roofline=roofline,
                    cholesky_tflops=chol_tf, cholesky_frac_of_peak=chol_tf / (peak_tf * (world if dist_train else 1)),
                    test_pts_per_s=m / sec_per_step, wall_s_per_step=wall_per_step, logp=logp,
                    phase_ms_per_step={k: v / args.steps for k, v in timers.items() if v})
        if not args.no_cpu_baseline and world == 1:
            ref = cpu_reference(n, m, d, args.cpu_sample_n, args.cpu_sample_m)
            line["cpu_baseline"] = dict(value=ref["value"], unit="s", cores=ref["cores"], kind=ref["kind"],
                                        sample=ref["sample"])
        print(json.dumps(line))
    if D is not None:
        D.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
