"""Synthetic per-rank inputs for the weak-scaling benchmark and the sharding checks.

The sharder itself is in the library: nutsb_multi (include/nutsb200.h, csrc/nutsb_multi.cuh) takes one global
population and one batch, deals the rooms to the shards, routes the ops (all-room and level ops replicated) and
returns the streams in global user order; api.MultiContext binds it.  This module only *generates* inputs: rank g's
own rooms, users and messages (bench.py's weak scaling: per-GPU work fixed), and `to_global`, which puts such shards
together into the one population / one batch that a single context -- or nutsb_multi -- renders as a whole."""
from __future__ import annotations

import numpy as np

from . import synth

SEED = synth.SEED


def shard_inputs(rank: int, n_msgs: int, n_users: int, users_per_room: int, words, gated: bool = True):
    """The rank's shard in LOCAL indices (what its own Context sees)."""
    seed = SEED + 0x1000 * rank
    users, n_rooms = synth.users(n_users, users_per_room, seed=seed)
    bt, bo = synth.bodies(n_msgs, words if gated else None, seed=seed)
    ops, spk, rm = synth.say_ops(n_msgs, n_users, users_per_room, bt, bo, gated=gated, seed=seed)
    return dict(users=users, n_rooms=n_rooms, bodies=(bt, bo), ops=ops, speaker=spk)


def to_global(shards):
    """Concatenates per-rank shards into one batch over the global population: user u of
    rank g becomes g*n_users+u, room r becomes g*n_rooms+r, swear gate m becomes
    (messages before rank g) + m.  -> ops, users, n_rooms, bodies"""
    texts, offs, base = [], [np.zeros(1, np.uint64)], 0
    kind, target, exc, flags, gate = [], [], [], [], []
    room, uflags, level = [], [], []
    bts, bos, bbase = [], [np.zeros(1, np.uint64)], 0
    u0 = r0 = m0 = 0
    for sh in shards:
        o, us = sh["ops"], sh["users"]
        nU, nR = len(us["room"]), sh["n_rooms"]
        texts.append(o["text"]); offs.append(o["off"][1:] + np.uint64(base)); base += int(o["off"][-1])
        k = o["kind"]
        t = o["target"].copy()
        t[k == 0] += u0                               # write_user: a user index
        t[(k == 1) & (o["target"] >= 0)] += r0        # room ops: a room index (-1 stays "all rooms")
        e = o["except_user"].copy(); e[e >= 0] += u0
        kind.append(k); target.append(t); exc.append(e); flags.append(o["flags"])
        if "gate" in o:
            g = o["gate"].copy(); g[g >= 0] += m0; gate.append(g)
        rr = us["room"].copy(); rr[rr >= 0] += r0
        room.append(rr); uflags.append(us["flags"]); level.append(us["level"])
        bt, bo = sh["bodies"]
        bts.append(bt); bos.append(bo[1:] + np.uint64(bbase)); bbase += int(bo[-1])
        u0 += nU; r0 += nR; m0 += len(bo) - 1
    ops = dict(text=np.concatenate(texts), off=np.concatenate(offs), kind=np.concatenate(kind),
               target=np.concatenate(target), except_user=np.concatenate(exc), flags=np.concatenate(flags))
    if gate:
        ops["gate"] = np.concatenate(gate)
    users = dict(room=np.concatenate(room), flags=np.concatenate(uflags), level=np.concatenate(level))
    return ops, users, r0, (np.concatenate(bts), np.concatenate(bos))
