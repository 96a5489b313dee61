#!/usr/bin/env python
"""bench.py -- MU iterations/s on BASELINE.json configs[1] (dense fp32 100k x 10k, k=64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one multiplicative-update iteration (H update, W update, column normalisation) over the whole
synthetic matrix.  Rank 0 prints ONE JSON line:
  value      iterations/s with V already resident in HBM (CUDA events on the engine's stream, max over ranks), issued the
             way the reference's run loop issues them: the residual is evaluated on every 10th and on the last iteration
             (SingleGpuDispatcher.cpp:173) and its host work is inside the events, as it is inside the reference's
             elapsedTime; strong scaling: the same 100k x 10k problem is column-sharded over the N GPUs
  e2e        the same metric through the reference-facing C ABI (nmfgpu_compute_single) with PAGEABLE host buffers (what
             an R caller hands over): H2D of V, W0, H0 and D2H of W, H inside the timed region; e2e_pinned: the same
             call from page-locked buffers.  The reference arm gets the same two kinds of buffers.
  roofline   the slower of the two V-streaming kernels against the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the fp64 OpenMP oracle timed on a bounded column sample of the same workload
`--impl reference` times the UNMODIFIED reference (oracle/_ref/libnmfgpu64_ref.so, compiled from
/root/reference by oracle/build_ref.sh: its cuBLAS fp32 path on the same B200) through its own public API;
if that library cannot be loaded the oracle port on the host cores stands in (cpu_baseline.kind = "port").
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M, N, K = 100_000, 10_000, 64          # BASELINE.json configs[1]
SEED_V, SEED_W, SEED_H = 42, 43, 44
NVSMI_QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
               "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
               "clocks_event_reasons.sw_power_cap")


def flops_per_iteration(m, n, k):
    return 4.0 * m * n * k + 4.0 * k * k * (m + n)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi in a loop (20 ms) from before the warm-up; stop(t0, t1) keeps the samples whose timestamps lie in
    [t0, t1] (time.time() of this host), the window the caller was running the timed iterations in."""

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + NVSMI_QUERY, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self, t0=None, t1=None):
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                stamp = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if t0 is not None and not (t0 <= stamp <= t1):
                    continue
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def cpu_baseline(sample_cols=2000, iters=8):
    """fp64 OpenMP oracle on the first `sample_cols` columns of the workload (bounded CPU time)."""
    from nmfgpu_b200.workloads import uniform_block
    from oracle import binding as orc
    V = uniform_block(SEED_V, M, sample_cols, total_rows=M)
    W0 = uniform_block(SEED_W, M, K)
    H0 = uniform_block(SEED_H, K, sample_cols, total_rows=K)
    orc.time_mu_iterations(V[:, :64], W0, H0[:, :64], 1)   # warm the library
    secs = orc.time_mu_iterations(V, W0, H0, iters)
    # the W-side Gram/update work does not shrink with the column sample; scale only the V-sized part
    per_iter = secs / iters
    full = per_iter * (N / sample_cols)
    return {"value": 1.0 / full, "unit": "iterations/s", "cores": orc.num_threads(), "kind": "port",
            "sample": "fp64 OpenMP oracle, %d MU iterations on the first %d of %d columns (%.2f s), scaled x%d to the full matrix"
                      % (iters, sample_cols, N, secs, N // sample_cols)}


def run_reference(args, rank, world, local):
    """--impl reference: the unmodified reference library on one B200 (rank 0 only)."""
    if rank != 0:
        return
    from nmfgpu_b200 import api
    from nmfgpu_b200.workloads import uniform_block
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libnmfgpu64_ref.so")
    line = {"impl": "reference", "metric": "MU iterations/s (dense fp32 100000x10000, k=64)", "unit": "iterations/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": "dense fp32 100000x10000, k=64, MU Frobenius, CopyExisting init",
                                                              "l2": "input (4 GB) larger than L2"}}
    try:
        REF = api.Library(ref_so)
        import torch  # noqa: F401  (only to make sure a CUDA runtime / GPU is usable in this process)
        ok = REF.initialize() == 0 and REF.number_of_gpu() > 0
    except Exception as e:  # noqa: BLE001
        ok = False
        sys.stderr.write("reference library unavailable (%s): timing the oracle port instead\n" % e)
    if ok:
        REF.set_verbosity(api.Verbosity.NoOutput)
        # host inputs; V is generated column-block-wise to bound temporaries
        V = np.empty((M, N), dtype=np.float32, order="F")
        for c0 in range(0, N, 500):
            V[:, c0:c0 + 500] = uniform_block(SEED_V, M, 500, total_rows=M, col0=c0)
        W0 = uniform_block(SEED_W, M, K)
        H0 = uniform_block(SEED_H, K, N)
        REF.compute(V, K, W0=W0, H0=H0, iterations=max(1, args.warmup))
        sampler = ClockSampler(0)
        time.sleep(0.5)
        w0 = time.time()
        t0 = time.perf_counter()
        r = REF.compute(V, K, W0=W0, H0=H0, iterations=args.steps)
        wall = time.perf_counter() - t0
        clocks = sampler.stop(w0, time.time())
        assert r["rc"] == 0, r["rc"]
        # the same call from page-locked memory (what our own e2e_pinned leg gets)
        wall_pinned = None
        try:
            import torch
            Vp = torch.empty((N, M), dtype=torch.float32, pin_memory=True)      # (N, M) C-order = (M, N) column-major
            Vp.numpy()[:, :] = V.T
            Vpin = Vp.numpy().T
            REF.compute(Vpin, K, W0=W0, H0=H0, iterations=2)
            t0 = time.perf_counter()
            rp = REF.compute(Vpin, K, W0=W0, H0=H0, iterations=args.steps)
            wall_pinned = time.perf_counter() - t0
            assert rp["rc"] == 0
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("pinned leg of the reference arm skipped: %s\n" % e)
        REF.finalize()
        inner = max(r["elapsed"], 1e-3)        # ExecutionRecord.elapsedTime: host clock, excludes setup (Dispatcher.cpp:166,181,218)
        value = args.steps / inner
        line.update(value=value, ms_per_step=1000.0 * inner / args.steps, clocks=clocks, gpu_launches=None,
                    cpu_baseline={"value": value, "unit": "iterations/s", "cores": 1, "kind": "reference",
                                  "sample": "unmodified reference (cuBLAS fp32 on the same B200), %d iterations, ExecutionRecord.elapsedTime" % args.steps},
                    e2e={"value": args.steps / wall, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                         "host_buffers": "pageable", "h2d_bytes_per_call": (M * N + M * K + K * N) * 4, "d2h_bytes_per_call": (M * K + K * N) * 4},
                    e2e_pinned=None if wall_pinned is None else {"value": args.steps / wall_pinned, "unit": "iterations/s", "host_buffers": "page-locked"},
                    effective_tflops=flops_per_iteration(M, N, K) * value / 1e12, final_frobenius=r["frobenius"])
    else:
        cb = cpu_baseline()
        line.update(value=cb["value"], ms_per_step=1000.0 / cb["value"], cpu_baseline=cb, gpu_launches=0,
                    e2e={"value": cb["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line))


def run_ours(args, rank, world, local):
    import torch
    import torch.distributed as dist
    from nmfgpu_b200 import api
    from nmfgpu_b200.workloads import shard_columns, uniform_block

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    sampler = ClockSampler(local) if rank == 0 else None      # started first: nvidia-smi needs up to a second for its first line on an 8-GPU box
    L = api.Library()
    L.set_verbosity(api.Verbosity.NoOutput)
    assert L.initialize() == 0
    assert L.choose_gpu(local) == 0
    c0, c1 = shard_columns(N, world, rank)
    nloc = c1 - c0
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (ctypes.c_ubyte * 128)()
            assert L.lib.nmfgpu_b200_dist_unique_id(buf) == 0
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        raw = bytes(uid.cpu().tolist())
        assert L.lib.nmfgpu_b200_dist_init(rank, world, raw) == 0
        assert L.lib.nmfgpu_b200_dist_set_shard(N, c0) == 0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident leg: V generated on the device (no 4 GB H2D), factors from the host generator
    ld = M
    dev = L.lib.nmfgpu_b200_device_alloc(ld * nloc * 4)
    assert dev, "device allocation of V failed"
    assert L.lib.nmfgpu_b200_device_uniform_f32(dev, M, nloc, ld, SEED_V, M, 0, c0) == 0
    W0 = uniform_block(SEED_W, M, K)
    H0 = uniform_block(SEED_H, K, nloc, total_rows=K, col0=c0)
    s = api.Session(L, "mu", M, nloc, K, device_ptr=dev, ld_v=ld)
    s.set_factors(W0, H0)
    s.iterate(args.warmup)                      # W >= 4 iterations also record the CUDA graph the timed batches replay
    s.iterate_with_error()                      # ... and the residual path (pinned buffers, the host combine) is warm
    s.synchronize()
    info0 = s.info()
    barrier()
    wall0 = time.time()
    ms, f_timed = s.time_run(args.steps)        # CUDA events on the engine's stream around exactly K iterations, residual cadence included
    barrier()
    info1 = s.info()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # clocks under this load: the timed region lasts 6-35 ms at 20 steps, less than two nvidia-smi samples, so the same
    # iterations keep running (untimed) until the sampling window is about 0.15 s long.  The number of extra batches
    # follows from the reduced time, so every rank runs the same number (the ranks of a sharded run iterate in lockstep:
    # a rank that decided by its own clock to do one batch more would wait for signals that never come)
    extra = int(np.ceil(max(0.0, 150.0 - ms) / max(ms / args.steps * 10.0, 1e-3)))
    for _ in range(extra):
        s.iterate(10)
        s.synchronize()
    barrier()
    clocks = sampler.stop(wall0, time.time()) if sampler else None
    if clocks is not None:
        clocks["window"] = "the timed iterations and %d more of the same iterations (about 0.15 s in all)" % (10 * extra)
    launches = int(info1.kernel_launches - info0.kernel_launches)
    collectives = int(info1.collective_calls - info0.collective_calls)
    f_final, _ = s.iterate_with_error()

    # ---- roofline of the two V-streaming kernels (each reads its block of V exactly once per launch).  Every rank
    # launches them: with row blocks the W^T V kernel stores its tiles into the other ranks' memory.
    roof = None
    reps, t_wtv, t_vht = 8, [], []
    barrier()
    s.iterate(10)     # queued ahead of the first product: the launches below find the GPU at the clocks of the iteration loop
                      # (timed from idle, a run measured 1.08 ms for the kernel that takes 0.74 ms inside the loop)
    for _ in range(reps):
        _, _, a, b = s.products(want_wtv=False, want_vht=False)
        t_wtv.append(a)
        t_vht.append(b)
    barrier()
    if rank == 0:
        a, b = float(np.mean(t_wtv)), float(np.mean(t_vht))
        peak, how = peaks()
        # algorithmic bytes per launch: V once + the small operand (hi and lo) + the partial outputs
        # (N GPUs: a rank streams 1/N of V either way -- a column shard, or the row block it was regrouped into)
        bytes_wtv = 4.0 * M * nloc + 8.0 * M * K / world + 4.0 * K * N * info1.splits_wtv / (world if info1.row_owners else 1)
        bytes_vht = 4.0 * M * nloc + 8.0 * K * N + 4.0 * M * K * info1.splits_vht / world
        name, tms, by = ("gemm_vht (V H^T)", b, bytes_vht) if b >= a else ("gemm_wtv (W^T V)", a, bytes_wtv)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(name.split()[0]) if world == 1 else None   # the capture is of the full 100k x 10k launch
        achieved = by / (tms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "peak_source": how, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "bytes_per_launch": by, "ms_per_launch": tms,
                "ms_gemm_wtv": a, "ms_gemm_vht": b, "ms_gemm_wtv_launches": [round(x, 4) for x in t_wtv],
                "ms_gemm_vht_launches": [round(x, 4) for x in t_vht], "uses_tensor_cores": bool(info1.uses_tensor_cores)}
    s.close()
    L.lib.nmfgpu_b200_device_free(dev)

    # ---- end-to-end legs through the reference-facing C ABI with host buffers (rank-local shard): first from ordinary
    # pageable memory -- what an nmfgpu4R caller hands over -- then from page-locked memory
    e2e_iters = args.steps
    nbytes = M * nloc * 4
    tmp = L.lib.nmfgpu_b200_device_alloc(nbytes)
    assert L.lib.nmfgpu_b200_device_uniform_f32(tmp, M, nloc, M, SEED_V, M, 0, c0) == 0

    def timed_call(Vhost):
        r = L.compute(Vhost, K, W0=W0, H0=H0, iterations=2)          # warm: module load, allocator
        assert r["rc"] == 0, r["rc"]
        barrier()
        t0 = time.perf_counter()
        r = L.compute(Vhost, K, W0=W0, H0=H0, iterations=e2e_iters)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        assert r["rc"] == 0, r["rc"]
        tw = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        return float(tw.item()), r

    Vpage = np.empty((nloc, M), dtype=np.float32)                    # (nloc, M) C-order = (M, nloc) column-major
    assert L.lib.nmfgpu_b200_device_download(Vpage.ctypes.data, tmp, nbytes) == 0
    wall, r = timed_call(Vpage.T)
    del Vpage
    hp = L.lib.nmfgpu_b200_host_alloc(nbytes)
    assert hp, "pinned host allocation failed"
    Vh = np.ctypeslib.as_array(ctypes.cast(hp, ctypes.POINTER(ctypes.c_float)), shape=(nloc, M)).T   # (M, nloc) Fortran view
    assert L.lib.nmfgpu_b200_device_download(hp, tmp, nbytes) == 0
    L.lib.nmfgpu_b200_device_free(tmp)
    wall_pinned, _ = timed_call(Vh)
    del Vh
    L.lib.nmfgpu_b200_host_free(hp)
    h2d = (M * nloc + M * K + K * nloc) * 4 * world
    d2h = (M * K + K * nloc) * 4 * world

    if rank == 0:
        value = args.steps / (ms * 1e-3)
        line = {
            "metric": "MU iterations/s (dense fp32 100000x10000, k=64)", "value": value, "unit": "iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "dense fp32 100000x10000, k=64, MU Frobenius, CopyExisting init (BASELINE configs[1])",
                       "arithmetic": "3xTF32 tcgen05 (fp32-equivalent)" if roof and roof["uses_tensor_cores"] else "fp32 SIMT",
                       "parallelism": ("single GPU" if world == 1 else
                                       "column shards x%d regrouped into row blocks; per iteration k x n partials of W^T V, columns of H and k*k+k "
                                       "statistics move as NVLink peer stores from inside the kernels (no NCCL call in the iteration)" % world
                                       if info1.row_owners else "column shards x%d, NCCL all-reduce of V H^T and H H^T" % world),
                       "cadence": "residual on every 10th and the last iteration inside the timed region (reference run loop)",
                       "l2": "input (4 GB) larger than the 126 MB L2; no flush needed"},
            "effective_tflops": flops_per_iteration(M, N, K) * value / 1e12,
            "hbm_roofline_iterations_per_s": 1.0 / ((8.0 * M * N + 16.0 * K * (M + N)) / (peaks()[0] * 1e9)) * world,
            "e2e": {"value": e2e_iters / wall, "unit": "iterations/s", "h2d_bytes_per_step": h2d / e2e_iters, "d2h_bytes_per_step": d2h / e2e_iters,
                    "host_buffers": "pageable",
                    "call": "nmfgpu_compute_single(numIterations=%d) from pageable host buffers, wall clock incl. H2D of V" % e2e_iters},
            "e2e_pinned": {"value": e2e_iters / wall_pinned, "unit": "iterations/s", "host_buffers": "page-locked"},
            "gpu_launches": launches, "collective_calls": collectives, "clocks": clocks, "roofline": roof,
            "timed_frobenius": f_timed, "final_frobenius": f_final, "e2e_frobenius": r["frobenius"],
        }
        try:
            line["cpu_baseline"] = cpu_baseline() if world == 1 else None
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"error": str(e)}
        print(json.dumps(line))
    L.finalize()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    rank, world, local = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world, local)
    else:
        run_ours(args, rank, world, local)


if __name__ == "__main__":
    main()
