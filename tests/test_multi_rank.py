"""World-size-2 run of the sharding logic on CPU (gloo): each rank renders its own rooms,
rank 0 checks that the gathered per-user digests equal those of the whole batch rendered in
one piece -- i.e. sharding by room needs no exchange step -- and that the timing/throughput
reduction bench.py uses (max of times, sum of deliveries) behaves."""
import hashlib
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
N_MSGS, N_USERS, UPR = 300, 120, 20


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import oracle_lib as O
    from nuts333_b200 import shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = O.port()
    words = synth.swear_words(64)
    sh = shard.shard_inputs(rank, N_MSGS, N_USERS, UPR, words)
    bt, bo = sh["bodies"]
    v = P.contains_swearing_batch(bt, bo, words)
    off, data, nd = P.write_batch(sh["ops"], sh["users"], verdict=v)
    digests = [hashlib.sha256(data[int(off[u]):int(off[u + 1])].tobytes()).hexdigest() for u in range(N_USERS)]
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(rank=rank, digests=digests, shard=sh, deliveries=int(nd.sum())))
    t = torch.tensor([1.0 + rank], dtype=torch.float64)          # pretend per-rank step times
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    d = torch.tensor([float(nd.sum())], dtype=torch.float64)
    dist.all_reduce(d, op=dist.ReduceOp.SUM)
    if rank == 0:
        shards = [g["shard"] for g in sorted(gathered, key=lambda g: g["rank"])]
        ops, users, n_rooms, (gbt, gbo) = shard.to_global(shards)
        gv = P.contains_swearing_batch(gbt, gbo, words)
        goff, gdata, gnd = P.write_batch(ops, users, verdict=gv)
        whole = [hashlib.sha256(gdata[int(goff[u]):int(goff[u + 1])].tobytes()).hexdigest() for u in range(world * N_USERS)]
        parts = sum((g["digests"] for g in sorted(gathered, key=lambda g: g["rank"])), [])
        q.put(dict(ok=whole == parts, t_max=float(t[0]), d_sum=float(d[0]), d_whole=int(gnd.sum()),
                   distinct=len(set(parts)) > world, n_rooms=n_rooms))
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
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["ok"], "sharded streams differ from the whole batch"
    assert res["t_max"] == 2.0 and res["d_sum"] == res["d_whole"] and res["distinct"]
    assert res["n_rooms"] == world * (N_USERS // UPR)
