"""The two multi-GPU dataflows of csrc/engine.cu (dist.h), restated in numpy over a world_size-2 gloo group.

Each rank holds a column shard of V (and, for the row-owner dataflow, the row block it would receive from the
grouped send/recv at setup) and exchanges exactly what the engine exchanges:
  all-reduce  : the m x k partial V H^T and the k x k partial H H^T;
  row owners  : all-gather of H together with the statistics of the local columns (H_g H_g^T), all-gather of the
                un-normalised row blocks of W together with their k*k + k statistics (Gram matrix -- its diagonal gives
                the column norms -- and column sums); every rank adds the gathered statistics up in rank order.
Both must reproduce the single-process MU restatement (tests/np_restatement.py).
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nmfgpu_b200.workloads import dense_inputs, shard_columns   # noqa: E402
from tests import np_restatement                                 # noqa: E402

M, N, K, ITERS = 300, 128, 7, 12
EPS = float(np.finfo(np.float32).eps)


def _all_reduce(x):
    t = torch.from_numpy(np.ascontiguousarray(x))
    dist.all_reduce(t)
    return t.numpy()


def _all_gather(x, world):
    t = torch.from_numpy(np.ascontiguousarray(x))
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [o.numpy() for o in out]


def _worker(rank, world, port, mode, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        V, W0, H0 = dense_inputs(M, N, K)
        V = V.astype(np.float64)
        c0, c1 = shard_columns(N, world, rank)
        Vc = V[:, c0:c1]
        W = W0.astype(np.float64)
        H = H0.astype(np.float64)[:, c0:c1]
        pad = 64                                            # row blocks are padded to a common size (256 in the engine)
        block = -(-(-(-M // world)) // pad) * pad
        r0 = rank * block
        r1 = min(M, r0 + block)
        Vr = V[r0:r1, :]                                    # what the grouped send/recv of the column shards delivers
        G = W.T @ W
        for _ in range(ITERS):
            Nn = W.T @ Vc
            H = H * Nn / (G @ H + EPS)
            if mode == "allreduce":
                B = _all_reduce(H @ H.T)
                P = _all_reduce(Vc @ H.T)
                W = W * P / (W @ B + EPS)
                s = (W * W).sum(axis=0)
                W = W / np.where(s > 0, np.sqrt(s), 1.0)
                G = W.T @ W
            else:
                Hfull = np.concatenate(_all_gather(H, world), axis=1)
                B = sum(_all_gather(H @ H.T, world))         # statistics of the local columns, added up in rank order
                Wb = W[r0:r1] * (Vr @ Hfull.T) / (W[r0:r1] @ B + EPS)
                stat = sum(_all_gather(np.concatenate([(Wb.T @ Wb).ravel(), Wb.sum(axis=0)]), world))
                gram = stat[:K * K].reshape(K, K)
                d = np.diag(gram)
                norm = np.where(d > 0, np.sqrt(d), 1.0)
                G = gram / np.outer(norm, norm)
                padded = np.zeros((block, K))
                padded[:r1 - r0] = Wb                       # the blocks travel un-normalised, the receiver scales
                W = np.concatenate(_all_gather(padded, world), axis=0)[:M] / norm
                colsum = stat[K * K:] / norm                # centring term of the next W^T V: column sums of the unit-column W
                assert np.allclose(colsum, W.sum(axis=0), rtol=1e-12)
        Hall = np.concatenate(_all_gather(H, world), axis=1)
        if rank == 0:
            ret["W"], ret["H"] = W, Hall
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["allreduce", "rowowners"])
def test_world2_matches_single_process(mode):
    V, W0, H0 = dense_inputs(M, N, K)
    Wref, Href, _ = np_restatement.run("mu", V, W0, H0, ITERS, EPS)
    with mp.Manager() as manager:
        ret = manager.dict()
        port = 29500 + (os.getpid() % 2000)
        mp.spawn(_worker, args=(2, port, mode, ret), nprocs=2, join=True)
        np.testing.assert_allclose(ret["W"], Wref, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(ret["H"], Href, rtol=1e-9, atol=1e-12)


def test_shards_cover_every_column_once():
    for n, world in ((10000, 8), (1001, 4), (7, 8), (128, 2)):
        cols = []
        for r in range(world):
            c0, c1 = shard_columns(n, world, r)
            cols.extend(range(c0, c1))
        assert cols == list(range(n))
