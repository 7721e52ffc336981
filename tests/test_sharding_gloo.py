"""CPU tier: the N>1 host logic (round-robin shard, host-side gather, max-over-ranks) on a world_size-2 gloo group."""
import os
import socket

import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from oai_analysis_2_b200 import sharding
    r, w, _ = sharding.init_process_group("gloo")
    mine = sharding.shard_indices(7, r, w)
    recs = [dict(index=i, rank=r, value=i * i) for i in mine]
    merged = sharding.gather_records(recs, r, w)
    slow = sharding.max_over_ranks(1.0 + r, w)
    sharding.barrier(w)
    q.put((r, mine, merged, slow))


def test_two_rank_shard_and_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == [0, 2, 4, 6] and got[1][1] == [1, 3, 5]
    assert [r["index"] for r in got[0][2]] == list(range(7)) and got[1][2] is None
    assert all(r["value"] == r["index"] ** 2 for r in got[0][2])
    assert got[0][3] == 2.0 and got[1][3] == 2.0


def test_single_rank_is_passthrough():
    from oai_analysis_2_b200 import sharding
    assert sharding.shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]
    assert sharding.gather_records([dict(index=1), dict(index=0)], 0, 1) == [dict(index=1), dict(index=0)]
    assert sharding.max_over_ranks(3.5, 1) == 3.5
