"""The stream-K work split of the tensor-core products (csrc/tc_gemm.cu: reduction chunks, equal shares of every chunk per
CTA, one partial product -- "slot" -- per CTA, tile and chunk), enumerated on the host through nmfgpu_b200_plan_segments.
Checked without a GPU:
  * every (tile, stage) of the product is covered exactly once,
  * the slots a tile receives are exactly 0 .. count-1, each written by exactly one segment (a consumer that adds up `count`
    partial products never reads one that nobody wrote; a segment in the padding of the last chunk still owns its slot),
  * the count the consumers are told (per 128-wide tile) is that number and fits the unsigned char it travels in,
  * the CTAs' shares differ by at most one unit per chunk.
"""
import ctypes
import os

import numpy as np
import pytest

from nmfgpu_b200 import _lib

SMS = 148
SHAPES = [
    # (A rows, reduction length, kp)                       what it is
    (10000, 100000, 64),     # W^T V of BASELINE configs[1]: 40 tiles, 3125 stages, 4 chunks
    (100000, 10000, 64),     # V H^T of the same
    (1250, 100000, 64),      # W^T V of an 8-GPU column shard: 5 tiles, one chunk
    (12500, 10000, 64),      # V H^T of an 8-GPU row block
    (20000, 50000, 128),     # configs[3]
    (50000, 20000, 128),
    (500, 1000, 16),         # configs[0]
    (33, 17, 16), (128, 128, 128), (129, 5000, 80), (3000, 64, 16), (1, 1, 16), (257, 33, 16), (70000, 3200000, 64),
]


def plan(lib, rows_a, reduce_len, kp, sms=SMS):
    info = (ctypes.c_uint * 5)()
    tiles128 = (rows_a + 127) // 128
    slots = (ctypes.c_ubyte * tiles128)()
    lib.nmfgpu_b200_plan_segments.restype = ctypes.c_uint
    n = lib.nmfgpu_b200_plan_segments(rows_a, reduce_len, kp, sms, None, 0, info, slots, tiles128)
    seg = (ctypes.c_uint * (5 * n))()
    assert lib.nmfgpu_b200_plan_segments(rows_a, reduce_len, kp, sms, seg, n, info, slots, tiles128) == n
    return np.array(seg, dtype=np.int64).reshape(n, 5), list(info), np.array(slots, dtype=np.int64)


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


@pytest.mark.parametrize("chunks_env", [None, "1", "3", "8"])
@pytest.mark.parametrize("shape", SHAPES)
def test_work_split_invariants(lib, monkeypatch, shape, chunks_env):
    if chunks_env is None:
        monkeypatch.delenv("NMFGPU_TC_CHUNKS", raising=False)
    else:
        monkeypatch.setenv("NMFGPU_TC_CHUNKS", chunks_env)
    rows_a, reduce_len, kp = shape
    seg, (tiles, stages_per_tile, chunks, chunk_stages, grid), slots128 = plan(lib, rows_a, reduce_len, kp)
    assert tiles == (rows_a + 255) // 256 and stages_per_tile == (reduce_len + 31) // 32
    assert 1 <= grid <= SMS and chunks * chunk_stages >= stages_per_tile > (chunks - 1) * chunk_stages
    # coverage: every (tile, stage) exactly once
    covered = np.zeros((tiles, stages_per_tile), dtype=np.int32)
    for cta, tile, stage0, length, slot in seg:
        assert tile < tiles and (length == 0 or stage0 + length <= stages_per_tile)     # a segment in the padding of the last chunk has no stages
        covered[tile, stage0:stage0 + length] += 1
    assert covered.min() == 1 and covered.max() == 1
    # slots: per tile exactly 0 .. count-1, count as announced to the consumers
    for tile in range(tiles):
        mine = np.sort(seg[seg[:, 1] == tile][:, 4])
        count = slots128[2 * tile]
        assert count <= 255 and list(mine) == list(range(count)), (tile, mine, count)
        if 2 * tile + 1 < len(slots128):
            assert slots128[2 * tile + 1] == count
    # balance: per chunk the shares differ by at most one unit, so overall by at most `chunks`
    work = np.bincount(seg[:, 0], weights=seg[:, 3], minlength=grid)
    padding = chunks * chunk_stages - stages_per_tile
    assert work.max() - work.min() <= chunks + padding * tiles


def test_default_chunks_follow_the_operand_size(lib, monkeypatch):
    monkeypatch.delenv("NMFGPU_TC_CHUNKS", raising=False)
    _, (_, _, chunks, _, _), slots = plan(lib, 10000, 100000, 64)      # W hi/lo 51 MB: 4 chunks of ~12 MB
    assert chunks == 4 and slots.max() <= 24
    _, (_, _, chunks, _, _), _ = plan(lib, 100000, 10000, 64)          # H^T hi/lo 5 MB: one chunk
    assert chunks == 1
    _, (_, _, chunks, _, _), slots = plan(lib, 1250, 100000, 64)       # 5 tiles on 148 CTAs: 30 partial products per chunk already
    assert chunks == 1
