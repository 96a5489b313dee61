"""The two multi-GPU dataflows of csrc/engine.cu (dist.h), restated in numpy over a world_size-2 gloo group.

Each rank holds a column shard of V (and, for the row-block dataflow, the row block it receives from the grouped
send/recv at setup) and exchanges exactly what the engine exchanges:
  all-reduce  : the m x k partial V H^T and the k x k partial H H^T;
  row blocks  : (csrc/fused.h) the k x N partial W[I_g]^T V[I_g, :] of every rank goes to the owners of the columns (a
                reduce-scatter; the engine does it with peer stores), the owners update their columns of H and every rank
                receives them (all-gather), V[I_g, :] H^T needs no reduction, every rank updates its own rows of W and
                keeps them UN-NORMALISED: what travels instead are k*k + k statistics per rank (Gram matrix of the rows --
                its diagonal gives the column norms -- and column sums, resp. H_g H_g^T and row sums), added up in rank
                order on every rank; the column scale is applied where W is read.  W is gathered only at the end.
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
        if mode == "allreduce":
            G = W.T @ W
            for _ in range(ITERS):
                Nn = W.T @ Vc
                H = H * Nn / (G @ H + EPS)
                B = _all_reduce(H @ H.T)
                P = _all_reduce(Vc @ H.T)
                W = W * P / (W @ B + EPS)
                s = (W * W).sum(axis=0)
                W = W / np.where(s > 0, np.sqrt(s), 1.0)
                G = W.T @ W
            Hall = np.concatenate(_all_gather(H, world), axis=1)
        else:
            # owners of the columns of H: blocks of `own` columns (a multiple of 128 in the engine), independent of the caller's shards
            own = -(-(-(-N // world)) // 16) * 16
            o0, o1 = min(N, rank * own), min(N, (rank + 1) * own)
            Hfull = np.concatenate(_all_gather(H, world), axis=1)           # setup: every rank starts from the whole H
            Wun = W[r0:r1].copy()                                           # this rank's rows, un-normalised from now on

            def w_statistics(normalise):
                stat = sum(_all_gather(np.concatenate([(Wun.T @ Wun).ravel(), Wun.sum(axis=0)]), world))   # rank order
                gram, colsum = stat[:K * K].reshape(K, K), stat[K * K:]
                d = np.diag(gram)
                inv = np.where(d > 0, 1.0 / np.sqrt(d), 1.0) if normalise else np.ones(K)
                return inv, gram * np.outer(inv, inv), colsum * inv

            inv, G, colsum = w_statistics(False)                            # the initial W is used as it is (MU.h:247)
            for _ in range(ITERS):
                # W^T V: the partial of the own row block, reduce-scattered to the owners of the columns
                partial = Wun.T @ Vr                                        # k x N, of the UN-NORMALISED rows
                parts = _all_gather(partial, world)
                Nown = inv[:, None] * sum(p[:, o0:o1] for p in parts)       # the column scale is applied to the sum
                Hown = Hfull[:, o0:o1] * Nown / (G @ Hfull[:, o0:o1] + EPS)
                # the owners' columns and their statistics go to every rank
                padded = np.zeros((K, own))
                padded[:, :o1 - o0] = Hown
                Hfull = np.concatenate(_all_gather(padded, world), axis=1)[:, :N]
                statH = sum(_all_gather(np.concatenate([(Hown @ Hown.T).ravel(), Hown.sum(axis=1)]), world))
                B = statH[:K * K].reshape(K, K)
                # own rows of W: scale applied as W is read, no reduction for V[I_g, :] H^T
                Wn = Wun * inv
                Wun = Wn * (Vr @ Hfull.T) / (Wn @ B + EPS)
                inv, G, colsum = w_statistics(True)
            # the factors the caller gets: unit columns, gathered rows; its own columns of H
            padded = np.zeros((block, K))
            padded[:r1 - r0] = Wun * inv
            W = np.concatenate(_all_gather(padded, world), axis=0)[:M]
            assert np.allclose(colsum, W.sum(axis=0), rtol=1e-12)           # centring term of the next W^T V
            Hall = Hfull
        if rank == 0:
            ret["W"], ret["H"] = W, Hall
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["allreduce", "rowblocks"])
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
