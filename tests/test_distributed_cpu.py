"""CPU, world_size 2, gloo: fragment sharding and the descriptor all-gather (the only exchange step of the path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from d3feat_b200.distributed import all_gather_descriptors, shard_fragments


def test_shard_fragments_partitions_everything():
    for n, w in [(64, 8), (7, 2), (3, 4), (0, 2)]:
        owned = [shard_fragments(n, r, w) for r in range(w)]
        flat = sorted(f for o in owned for f in o)
        assert flat == list(range(n))
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frags = shard_fragments(5, rank, world)          # rank 0: 0,2,4   rank 1: 1,3
        rows = [0 if f == 3 else 10 + 3 * f for f in frags]   # fragment 3 is empty: keeps its slot
        desc = torch.cat([torch.full((r, 4), float(f)) + torch.arange(r)[:, None] / 100.0 for f, r in zip(frags, rows)], 0)
        all_desc, all_rows, owner = all_gather_descriptors(desc, rows)
        np.save(os.path.join(out_dir, "d%d.npy" % rank), all_desc.numpy())
        np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array(all_rows))
        np.save(os.path.join(out_dir, "o%d.npy" % rank), np.array(owner))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_all_gather_descriptors_gloo_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    d0, d1 = np.load(tmp_path / "d0.npy"), np.load(tmp_path / "d1.npy")
    assert np.array_equal(d0, d1)                       # every rank ends with the same gathered matrix
    rows = np.load(tmp_path / "r0.npy").tolist()
    owner = np.load(tmp_path / "o0.npy").tolist()
    assert rows == [10, 16, 22, 13, 0] and owner == [0, 0, 0, 1, 1]   # unequal counts per rank, zero-row fragment kept
    assert d0.shape == (sum(rows), 4)
    # fragment ids are encoded in the integer part: rank 0's fragments first, then rank 1's, padding removed
    frag_of_row = np.floor(d0[:, 0]).astype(int)
    expect = np.concatenate([np.full(r, f) for f, r in zip([0, 2, 4, 1, 3], rows)])
    assert np.array_equal(frag_of_row, expect)


def test_all_gather_without_process_group_is_identity():
    d = torch.randn(7, 3)
    a, rows, owner = all_gather_descriptors(d, [3, 4])
    assert a is d and rows == [3, 4] and owner == [0, 0]
