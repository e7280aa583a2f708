"""World-size-2 run of the library's sharder on CPU (gloo): one process per shard as under torchrun
(nutsb_multi_create_rank on the emulator build), every rank given the SAME global population and batch -- with
shouts, broadcasts and write_level ops that are replicated to every shard.  Rank 0 checks that the ranks' users
partition the population and that their streams, put together, are the oracle's for the whole batch; and that the
timing / throughput reduction bench.py uses (max of times, sum of deliveries) behaves."""
import os
import socket
import sys
from pathlib import Path

import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    import ctypes
    import torch
    import oracle_lib as O
    from cpusim.build_sim import build_sim
    from nuts333_b200 import api
    from test_multi import make_case
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = api.bind(ctypes.CDLL(str(build_sim())))
    c = make_case(71, 40, 6, 160)                                  # the same on every rank
    o, us = c["ops"], c["users"]
    m = api.MultiContext(lib=lib, rank=(world, rank, 0))
    m.set_users(us["room"], us["flags"], us["level"], c["n_rooms"])
    _, ush, _ = m.plan()
    got, total, deliv = m.write_batch(o)
    mine = {u: got[u] for u in range(40) if ush[u] == rank}
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(rank=rank, streams=mine, deliveries=deliv, others_empty=all(got[u] == b"" for u in range(40) if ush[u] != rank)))
    t = torch.tensor([1.0 + rank], dtype=torch.float64)          # pretend per-rank step times
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    d = torch.tensor([float(deliv)], dtype=torch.float64)
    dist.all_reduce(d, op=dist.ReduceOp.SUM)
    if rank == 0:
        P = O.port()
        off, data, nd = P.write_batch(o, us, verdict=o["verdict"])
        users_seen = sorted(u for g in gathered for u in g["streams"])
        ok = users_seen == list(range(40)) and all(g["others_empty"] for g in gathered)
        for g in gathered:
            for u, s in g["streams"].items():
                ok = ok and s == data[int(off[u]):int(off[u + 1])].tobytes()
        q.put(dict(ok=ok, t_max=float(t[0]), d_sum=float(d[0]), d_whole=int(nd.sum()), ranks_with_users=sum(1 for g in gathered if g["streams"])))
    m.close()
    dist.barrier()
    dist.destroy_process_group()


def test_room_sharding_two_ranks_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["ok"], "the ranks' streams differ from the whole batch"
    assert res["t_max"] == 2.0 and res["d_sum"] == res["d_whole"] and res["ranks_with_users"] == world
