#!/usr/bin/env python
"""dist_check.py -- column-sharded runs against the single-GPU run of the same problem, on the GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py

Every rank first factorises the whole matrix on its own GPU (no communicator), then its column shard with
NMFGPU_DIST_MODE = allreduce and with the default (row blocks with peer-store exchange, dist.h), one process per GPU
(NCCL + CUDA IPC transport; tests/test_sharded_gpu.py runs the same dataflows with the ranks as threads).  W, the gathered H and the residual of both
sharded runs must agree with the single-GPU run within the tolerance of tests/test_parity_gpu.py.
"""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmfgpu_b200 import api                                       # noqa: E402
from nmfgpu_b200.workloads import dense_inputs, shard_columns     # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = api.Library()
L.set_verbosity(api.Verbosity.NoOutput)
assert L.initialize() == 0
assert L.choose_gpu(local) == 0
shapes = [(6000, 4096, 32, 30), (20000, 2048 * world, 64, 20)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
failed = False
for (m, n, k, iters) in shapes:
    V, W0, H0 = dense_inputs(m, n, k, seed=5)
    s = api.Session(L, "mu", m, n, k, V=V)
    s.set_factors(W0, H0)
    s.iterate(iters)
    f1, _ = s.iterate_with_error()
    W1, H1 = s.get_factors()
    s.close()

    c0, c1 = shard_columns(n, world, rank)
    uid = torch.zeros(128, dtype=torch.uint8)
    for mode in ("allreduce", "rowblocks"):
        os.environ["NMFGPU_DIST_MODE"] = mode
        if rank == 0:
            buf = (ctypes.c_ubyte * 128)()
            assert L.lib.nmfgpu_b200_dist_unique_id(buf) == 0
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        u = uid.cuda()
        dist.broadcast(u, 0)
        assert L.lib.nmfgpu_b200_dist_init(rank, world, bytes(u.cpu().tolist())) == 0
        assert L.lib.nmfgpu_b200_dist_set_shard(n, c0) == 0
        s = api.Session(L, "mu", m, c1 - c0, k, V=np.asfortranarray(V[:, c0:c1]))
        s.set_factors(W0, np.asfortranarray(H0[:, c0:c1]))
        before = s.info().collective_calls
        s.iterate(iters)
        f2, _ = s.iterate_with_error()
        W2, H2 = s.get_factors()
        calls = s.info().collective_calls - before
        s.close()
        assert L.lib.nmfgpu_b200_dist_finalize() == 0
        eW = np.linalg.norm(W2 - W1) / np.linalg.norm(W1)
        eH = np.linalg.norm(H2 - H1[:, c0:c1]) / np.linalg.norm(H1[:, c0:c1])
        ef = abs(f2 - f1) / f1
        ok = eW <= 1e-5 and eH <= 1e-5 and ef <= 5e-6
        failed |= not ok
        print("rank %d %s %-9s W %.2e  H %.2e  residual %.9g vs %.9g (%.1e)  collectives %d  %s"
              % (rank, (m, n, k), mode, eW, eH, f2, f1, ef, calls, "ok" if ok else "MISMATCH"), flush=True)

# ---- compressed execution of a sparse input (csrc/spmm.cu) over column shards: all-reduce dataflow, through the reference API
import scipy.sparse as sp                                          # noqa: E402
from nmfgpu_b200.workloads import planted_inputs                   # noqa: E402
m, n, k, iters = 3000, 2048, 24, 20
rng = np.random.default_rng(5)
D = ((rng.random((m, n)) < 0.01) * (0.1 + rng.random((m, n)))).astype(np.float32)
_, W0, H0 = planted_inputs(m, n, k, seed=2)


def csr_desc(block):
    S = sp.csr_matrix(block)
    keep = (S.data.astype(np.float32), S.indptr.astype(np.int32), S.indices.astype(np.int32))
    return api.sparse_description(api.StorageFormat.CSR, block.shape[0], block.shape[1], *keep), keep


os.environ["NMFGPU_SPARSE"] = "1"
desc, keep = csr_desc(D)
single = L.compute(None, k, W0=W0, H0=H0, iterations=iters, sparse=(desc, np.dtype(np.float32)))
c0, c1 = shard_columns(n, world, rank)
if rank == 0:
    buf = (ctypes.c_ubyte * 128)()
    assert L.lib.nmfgpu_b200_dist_unique_id(buf) == 0
    uid = torch.tensor(list(buf), dtype=torch.uint8)
u = uid.cuda()
dist.broadcast(u, 0)
assert L.lib.nmfgpu_b200_dist_init(rank, world, bytes(u.cpu().tolist())) == 0
assert L.lib.nmfgpu_b200_dist_set_shard(n, c0) == 0
desc, keep = csr_desc(np.ascontiguousarray(D[:, c0:c1]))
shard = L.compute(None, k, W0=W0, H0=np.asfortranarray(H0[:, c0:c1]), iterations=iters, sparse=(desc, np.dtype(np.float32)))
assert L.lib.nmfgpu_b200_dist_finalize() == 0
assert single["rc"] == 0 and shard["rc"] == 0, (single["rc"], shard["rc"])
eW = np.linalg.norm(shard["W"] - single["W"]) / np.linalg.norm(single["W"])
eH = np.linalg.norm(shard["H"] - single["H"][:, c0:c1]) / np.linalg.norm(single["H"][:, c0:c1])
ef = abs(shard["frobenius"] - single["frobenius"]) / single["frobenius"]
ok = eW <= 2e-4 and eH <= 2e-4 and ef <= 2e-5
failed |= not ok
print("rank %d sparse %s W %.2e  H %.2e  residual %.9g vs %.9g (%.1e)  %s"
      % (rank, (m, n, k), eW, eH, shard["frobenius"], single["frobenius"], ef, "ok" if ok else "MISMATCH"), flush=True)
L.finalize()
dist.destroy_process_group()
sys.exit(1 if failed else 0)
