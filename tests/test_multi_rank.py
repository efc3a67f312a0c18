"""CPU tier: the N>1 host logic (index sharding + VRF min-reduction) on a world-size-2 gloo group.
Each rank computes its shard with the ORACLE (no GPU here); the reduction code is the product's."""
import importlib
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, start, count, n, num_labels, q):
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from oracle import pyoracle as orc
    sharding = importlib.import_module("go-spacemesh_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = orc.py_commitment(bytes(32), bytes(range(32)))
    s, k = sharding.shard_range(start, count, world, rank)
    labels, found, idx, l32 = orc.c_labels_range(c, n, s, k, orc.py_vrf_difficulty(num_labels), threads=2)
    best = sharding.allgather_vrf((idx, l32) if found else None)
    q.put((rank, s, k, labels.tobytes(), best))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("count,num_labels", [(1001, 64), (7, 4), (300, 10**9)])
def test_two_rank_shards_and_vrf_reduce(orc, count, num_labels):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, start, n, world = _free_port(), 2**32 - 500, 4, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, start, count, n, num_labels, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c = orc.py_commitment(bytes(32), bytes(range(32)))
    exp, found, idx, l32 = orc.c_labels_range(c, n, start, count, orc.py_vrf_difficulty(num_labels), threads=2)
    # shards tile the range exactly, in order
    assert res[0][1] == start and res[0][1] + res[0][2] == res[1][1] and res[1][1] + res[1][2] == start + count
    assert res[0][3] + res[1][3] == exp.tobytes()
    # every rank holds the same global minimum, equal to the single-process scan
    for r in res:
        assert r[4] == ((idx, l32) if found else None)


def test_shard_range_properties():
    sharding = importlib.import_module("go-spacemesh_b200.sharding")
    for world in (1, 2, 3, 4, 8):
        for count in (0, 1, 7, 8, 9, 2**34, 2**37 + 5):
            cover = 0
            nxt = 123
            for r in range(world):
                s, k = sharding.shard_range(123, count, world, r)
                assert s == nxt or k == 0
                nxt = s + k
                cover += k
            assert cover == count


def test_vrf_record_roundtrip_and_order():
    sharding = importlib.import_module("go-spacemesh_b200.sharding")
    a = (2**63 + 5, bytes([0, 1] + [255] * 30))
    b = (3, bytes([0, 1] + [255] * 29 + [254]))
    c = (2, a[1])
    assert sharding.decode_vrf(sharding.encode_vrf(a)) == a
    assert sharding.decode_vrf(sharding.encode_vrf(None)) is None
    recs = [sharding.encode_vrf(x) for x in (a, None, b, c)]
    assert sharding.reduce_vrf(recs) == b                       # smallest label wins
    assert sharding.reduce_vrf([recs[0], recs[3]]) == c          # equal labels: lowest index wins
    assert sharding.reduce_vrf([recs[1]]) is None
