#!/usr/bin/env python
"""run_config.py -- BASELINE.json configs[3] and configs[4] on 1..8 B200 (one process per GPU under torchrun), one JSON line
per run on stdout (rank 0).  GPU box only.

  cfg4   dense 50 000 x 20 000, k = 128, k-means initialisation (timed separately), GDCLS / AHCLS / nsNMF (or any algorithm)
         python tools/run_config.py cfg4 [--algo gdcls,ahcls,nsnmf] [--iters 20]
  cfg5   sparse CSR/CSC 1 000 000 x 100 000 at 0.1 % density, k = 100, MU in compressed execution (csrc/spmm.cu)
         python tools/run_config.py cfg5 [--iters 20]
  N GPUs: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_config.py ...
The matrix is column-sharded; every rank generates its own shard.  Times are CUDA-event times on the
engine's stream (cfg4, sessions with V resident) or the library's ExecutionRecord (cfg5, nmfgpu_compute_single), max over ranks.
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmfgpu_b200 import api                                        # noqa: E402
from nmfgpu_b200.workloads import shard_columns, uniform_block     # noqa: E402

PARAMS = {"mu": {}, "gdcls": {"lambda": 0.01}, "als": {}, "acls": {"lambdaW": 0.01, "lambdaH": 0.01},
          "ahcls": {"lambdaW": 0.01, "lambdaH": 0.01, "alphaW": 0.01, "alphaH": 0.01}, "nsnmf": {"theta": 0.5}}

ap = argparse.ArgumentParser()
ap.add_argument("config", choices=["cfg4", "cfg5"])
ap.add_argument("--algo", default=None)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--m", type=int, default=None)
ap.add_argument("--n", type=int, default=None)
ap.add_argument("--k", type=int, default=None)
ap.add_argument("--density", type=float, default=1e-3)
ap.add_argument("--init", default="kmeans", choices=["kmeans", "random"])
a = ap.parse_args()

rank, world, local = (int(os.environ.get(v, d)) for v, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
import torch                                                         # noqa: E402
import torch.distributed as dist                                     # noqa: E402
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

L = api.Library()
L.set_verbosity(api.Verbosity.NoOutput)
assert L.initialize() == 0 and L.choose_gpu(local) == 0


def max_over_ranks(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def dist_setup(n, c0):
    if world == 1:
        return
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (ctypes.c_ubyte * 128)()
        assert L.lib.nmfgpu_b200_dist_unique_id(buf) == 0
        uid = torch.tensor(list(buf), dtype=torch.uint8)
    uid = uid.cuda()
    dist.broadcast(uid, 0)
    assert L.lib.nmfgpu_b200_dist_init(rank, world, bytes(uid.cpu().tolist())) == 0
    assert L.lib.nmfgpu_b200_dist_set_shard(n, c0) == 0


if a.config == "cfg4":
    m, n, k = a.m or 50_000, a.n or 20_000, a.k or 128
    algos = (a.algo or "gdcls,ahcls,nsnmf").split(",")
    c0, c1 = shard_columns(n, world, rank)
    nloc = c1 - c0
    dist_setup(n, c0)
    dev = L.lib.nmfgpu_b200_device_alloc(m * nloc * 4)
    assert dev and L.lib.nmfgpu_b200_device_uniform_f32(dev, m, nloc, m, 42, m, 0, c0) == 0
    for algo in algos:
        s = api.Session(L, algo, m, nloc, k, device_ptr=dev, ld_v=m, params=PARAMS[algo])
        barrier()
        method = api.NmfInitializationMethod.KMeansAndRandomValues if a.init == "kmeans" else api.NmfInitializationMethod.AllRandomValues
        ms_init = max_over_ranks(s.initialize(method, 7))
        s.iterate(6)                         # graphs recorded, kernels loaded
        f0, _ = s.iterate_with_error()
        s.synchronize()
        barrier()
        ms, f1 = s.time_run(a.iters)         # reference cadence: residual on every 10th and the last iteration
        ms = max_over_ranks(ms)
        info = s.info()
        flops = 4.0 * m * n * k + 4.0 * k * k * (m + n)
        if rank == 0:
            print(json.dumps({"config": "cfg4: dense %d x %d, k=%d, %s, %s init" % (m, n, k, algo, a.init), "n_gpus": world, "iterations": a.iters,
                              "ms_per_iteration": ms / a.iters, "iterations_per_s": 1000.0 * a.iters / ms, "init_ms": ms_init,
                              "effective_tflops": flops * a.iters / ms / 1e9, "tensor_cores": bool(info.uses_tensor_cores),
                              "row_blocks": bool(info.row_owners), "frobenius_before": f0, "frobenius_after": f1}), flush=True)
        s.close()
    L.lib.nmfgpu_b200_device_free(dev)
else:
    m, n, k = a.m or 1_000_000, a.n or 100_000, a.k or 100
    c0, c1 = shard_columns(n, world, rank)
    nloc = c1 - c0
    dist_setup(n, c0)
    t0 = time.time()
    # CSC shard: Poisson(density * m) entries per column at distinct-ish uniform rows (duplicates merged)
    rng = np.random.default_rng(1000 + rank)
    counts = rng.poisson(a.density * m, size=nloc).astype(np.int64)
    cols = np.repeat(np.arange(nloc, dtype=np.int64), counts)
    rows = rng.integers(0, m, size=int(counts.sum()), dtype=np.int64)
    key = np.unique(cols * m + rows)
    cols, rows = (key // m).astype(np.int32), (key % m).astype(np.int32)
    nnz = len(key)
    vals = (1.0 - rng.random(nnz, dtype=np.float32)).astype(np.float32)
    indptr = np.zeros(nloc + 1, dtype=np.int32)
    np.cumsum(np.bincount(cols, minlength=nloc), out=indptr[1:])
    build_s = time.time() - t0
    desc = api.sparse_description(api.StorageFormat.CSC, m, nloc, vals, indptr, rows, 0)
    W0 = uniform_block(43, m, k)
    H0 = uniform_block(44, k, nloc, total_rows=k, col0=c0)
    r = L.compute(None, k, W0=W0, H0=H0, iterations=10, sparse=(desc, np.dtype(np.float32)))     # pays context, kernel loading, pools
    assert r["rc"] == 0, r
    barrier()
    t0 = time.time()
    r = L.compute(None, k, W0=W0, H0=H0, iterations=a.iters, sparse=(desc, np.dtype(np.float32)))
    wall = max_over_ranks(time.time() - t0)
    assert r["rc"] == 0, r
    inner = max_over_ranks(max(r["elapsed"], 1e-3))
    total_nnz = max_over_ranks(float(nnz)) if world == 1 else None
    t = torch.tensor([float(nnz)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t)
    total_nnz = float(t.item())
    if rank == 0:
        print(json.dumps({"config": "cfg5: sparse %d x %d, nnz %.4g (density %.3g), k=%d, MU, compressed execution" % (m, n, total_nnz, total_nnz / (m * n), k),
                          "n_gpus": world, "iterations": a.iters, "ms_per_iteration": 1000.0 * inner / a.iters, "iterations_per_s": a.iters / inner,
                          "wall_s_whole_call": wall, "host_build_s": build_s, "frobenius": r["frobenius"],
                          "gather_bytes_per_iteration": 2.0 * total_nnz * 4 * k,
                          "algorithmic_bytes_per_iteration": 2 * (8.0 * total_nnz + 4.0 * (m + 1)) + 16.0 * k * (m + n)}), flush=True)
L.finalize()
if world > 1:
    dist.destroy_process_group()
