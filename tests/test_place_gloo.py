"""world_size-2 gloo test (CPU) of the place-recognition exchange: partition -> local top-k -> all-gather ->
deterministic merge must equal a single-shard brute force.  The local top-k is a numpy stand-in here (the CUDA
kernel is covered by tests/test_gpu_match.py::test_db_top2_shard); this test covers the N>1 host logic."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _numpy_local_topk(shard, q, k, th_votes):
    q = np.asarray(q, np.uint8)
    d = np.unpackbits(q[:, None, :] ^ shard.desc[None, :, :], axis=2).sum(2).astype(np.int64)
    gidx = shard.first_kf * shard.desc_per_kf + np.arange(shard.n_desc, dtype=np.int64)
    keys = (d << 48) | gidx[None, :]
    keys = np.sort(keys, axis=1)[:, :k]
    if keys.shape[1] < k:
        keys = np.concatenate([keys, np.full((len(q), k - keys.shape[1]), -1, np.int64)], 1)
    votes = np.zeros(shard.n_kf, np.int32)
    best = keys[:, 0]
    ok = (best >> 48) <= th_votes
    np.add.at(votes, ((best[ok] & ((1 << 48) - 1)) - shard.first_kf * shard.desc_per_kf) // shard.desc_per_kf, 1)
    return torch.from_numpy(keys), torch.from_numpy(votes)


def _worker(rank, world, port, q, db, dpk, ret):
    sys.path.insert(0, ROOT)
    from swarmmap_b200 import place
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parts = place.partition(len(db) // dpk, world)
    first, n = parts[rank]
    shard = place.PlaceShard(db[first * dpk:(first + n) * dpk], dpk, first, local_topk=_numpy_local_topk)
    keys, votes = shard.query(q, 2, 50)
    ret[rank] = (keys.numpy().copy(), votes.numpy().copy(), first, n)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_topk_equals_single_shard(world):
    sys.path.insert(0, ROOT)
    from swarmmap_b200 import place
    rng = np.random.default_rng(42)
    dpk = 4
    db = rng.integers(0, 256, (4 * 101, 32), dtype=np.uint8)  # 101 keyframes: uneven split
    q = rng.integers(0, 256, (37, 32), dtype=np.uint8)
    q[3] = db[50]
    db[333] = db[50]  # exact tie across shards -> lower global index must win
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, q, db, dpk, ret), nprocs=world, join=True)
    whole = place.PlaceShard(db, dpk, 0, local_topk=_numpy_local_topk)
    ref_keys, ref_votes = whole.query_local(q, 2, 50)
    votes = np.zeros(len(db) // dpk, np.int32)
    for r in range(world):
        keys, v, first, n = ret[r]
        np.testing.assert_array_equal(keys, ref_keys.numpy())  # every rank holds the same merged result
        votes[first:first + n] += v
    d, i = place.unpack_keys(ref_keys)
    assert int(d[3, 0]) == 0 and int(i[3, 0]) == 50 and int(i[3, 1]) == 333 and int(d[3, 1]) == 0
    np.testing.assert_array_equal(votes, ref_votes.numpy())  # shard votes sum to the single-shard histogram


def test_partition_and_merge_units():
    sys.path.insert(0, ROOT)
    from swarmmap_b200 import place
    assert place.partition(10, 3) == [(0, 4), (4, 3), (7, 3)]
    assert place.partition(100000, 8)[-1] == (87500, 12500)
    g = torch.tensor([[[5, 9]], [[3, -1]], [[5, 7]]], dtype=torch.int64)  # (world=3, nq=1, k=2)
    assert place.merge_topk(g, 2).tolist() == [[3, 5]]
    g = torch.full((2, 1, 2), -1, dtype=torch.int64)
    assert place.merge_topk(g, 2).tolist() == [[-1, -1]]
