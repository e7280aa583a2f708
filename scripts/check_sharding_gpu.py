"""SURVEY.md 8(e) on real GPUs: every rank renders its own rooms on its own B200; rank 0 then renders the WHOLE
batch (all ranks' rooms, users and messages in one population) on its GPU and checks that the per-user stream
digests are the same -- sharding by room needs no exchange step -- and that sampled streams equal the oracle's.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/check_sharding_gpu.py
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from nuts333_b200 import api, build, shard, synth  # noqa: E402

N_MSGS, N_USERS, UPR = 200_000, 5_000, 100


def render(ctx, ops, users, n_rooms, bodies, words):
    bt, bo = bodies
    ctx.set_swear_words(words)
    v = ctx.contains_swearing_batch(bt, bo)
    ctx.set_users(users["room"], users["flags"], users["level"], n_rooms)
    st = ctx.write_batch(dict(ops, verdict=v))
    return st, ctx.stream_digests().copy(), v


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    build.build()
    words = synth.swear_words(64)
    sh = shard.shard_inputs(rank, N_MSGS, N_USERS, UPR, words)
    ctx = api.Context(local)
    st, dg, v = render(ctx, sh["ops"], sh["users"], sh["n_rooms"], sh["bodies"], words)
    # digests travel as int64 bit patterns (NCCL has no uint64)
    mine = torch.from_numpy(dg.view(np.int64)).cuda()
    allg = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allg, mine)
    tot = torch.tensor([float(st.n_deliveries)], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot)
    ok = True
    if rank == 0:
        shards = [shard.shard_inputs(r, N_MSGS, N_USERS, UPR, words) for r in range(world)]
        ops, users, n_rooms, bodies = shard.to_global(shards)
        gst, gdg, gv = render(ctx, ops, users, n_rooms, bodies, words)
        parts = np.concatenate([t.cpu().numpy().view(np.uint64) for t in allg])
        same = bool((parts == gdg).all())
        import oracle_lib as O
        P = O.port()
        pick = sorted(np.random.RandomState(9).choice(len(users["room"]), 12, replace=False).tolist())
        o, d, nd = P.write_batch(ops, users, verdict=gv, only_users=pick)
        oracle_ok = all(gst.user(u) == d[int(o[u]):int(o[u + 1])].tobytes() for u in pick)
        ok = same and oracle_ok and int(tot.item()) == int(gst.n_deliveries)
        print(json.dumps(dict(check="room sharding on GPUs", n_gpus=world, users=len(users["room"]), rooms=n_rooms,
                              msgs=world * N_MSGS, deliveries=int(gst.n_deliveries), sharded_digests_equal_whole_batch=same,
                              sampled_streams_equal_oracle=oracle_ok, ok=ok)))
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
