#!/usr/bin/env python
"""bench.py -- rendered messages/s of the NUTS message path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input per rank:
BASELINE config 3 (1M say() messages x 10k users in 100 rooms of 100, 64-word swear
list, ban_swearing on: contains_swearing -> gated write_user / write_room_except ->
colour render + fan-out) plus config 4's ban verdicts (100k sites + 100k names vs
10k-entry lists).  Ranks own disjoint rooms (weak scaling, no data-path collective).

value = deliveries/s ("rendered messages/s") with inputs resident in HBM, timed with
CUDA events on the stream the kernels run on, max over ranks.  e2e = the same through
the host-buffer C-ABI (nutsb_*_batch with pinned host inputs, per-user streams copied
back to pinned host memory).  roofline = the render+fan-out kernel's algorithmic bytes
over its own CUDA-event time vs the measured HBM copy bandwidth.  cpu_baseline = the
reference's own C routines (oracle/_ref, built from nuts333.c) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "rendered messages/sec (colour+swear+ban)"
UNIT = "deliveries/s"

# per-rank shard of the workload
N_MSGS = 1_000_000
N_USERS = 10_000
USERS_PER_ROOM = 100
N_SWEAR = 64
KERNEL_TIMING_STEPS = 3
IN_FLIGHT = 1
N_BAN_QUERIES = 100_000
N_BAN_ENTRIES = 10_000
SEED = 0x333


def workload_name():
    return ("C3+C4 per rank: %d say() msgs x %d users (%d rooms of %d), %d-word swear list, ban_swearing; "
            "%d sites + %d names vs %d-entry ban lists" % (N_MSGS, N_USERS, N_USERS // USERS_PER_ROOM, USERS_PER_ROOM,
                                                           N_SWEAR, N_BAN_QUERIES, N_BAN_QUERIES, N_BAN_ENTRIES))


def make_inputs(rank: int, n_msgs: int):
    """The rank's shard (its own rooms, users and messages: nuts333_b200/shard.py), generated
    on the host once; the same generator feeds every leg."""
    from nuts333_b200 import shard, synth
    seed = SEED + 0x1000 * rank
    words = synth.swear_words(N_SWEAR)
    sh = shard.shard_inputs(rank, n_msgs, N_USERS, USERS_PER_ROOM, words, gated=True)
    st, so = synth.sites(N_BAN_QUERIES, seed=seed)
    nt, no = synth.names(N_BAN_QUERIES)
    sfile = synth.ban_file(0, N_BAN_ENTRIES, N_BAN_QUERIES, N_BAN_QUERIES, True, seed=seed)
    ufile = synth.ban_file(1, N_BAN_ENTRIES, N_BAN_QUERIES, N_BAN_QUERIES, True, seed=seed)
    return dict(words=words, users=sh["users"], n_rooms=sh["n_rooms"], bodies=sh["bodies"], ops=sh["ops"],
                speaker=sh["speaker"], sites=(st, so), names=(nt, no), sfile=sfile, ufile=ufile)


# ---------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.proc, self.th = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        """Keep the samples taken inside [t0, t1] (the timed region); if the region was
        shorter than the sampling period, keep the nearest ones taken under the same load."""
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        if not inside:
            inside = sorted(self.samples, key=lambda s: min(abs(s[0] - t0), abs(s[0] - t1)))[:5]
        self.samples = inside

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for _, s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, val in zip(names, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own C routines on host cores
# ---------------------------------------------------------------------------------------
def _ref_worker(args):
    """One process = one single-threaded reference instance (its state is global) on one
    room-shard of the sample.  Returns (deliveries, bytes, seconds, kind)."""
    rank, shard, n_shards, n_msgs, n_ban = args
    import tempfile
    import oracle_lib as O
    inp = make_inputs(rank, n_msgs)
    words, users, ops = inp["words"], inp["users"], inp["ops"]
    R = O.ref()
    if R is not None:
        R._tmp = tempfile.mkdtemp(prefix="nutsref_cwd_")     # per process: the reference reads ./datafiles/
    kind = "reference" if R is not None else "port"
    P = O.port()
    bt, bo = inp["bodies"]
    # this worker's share: messages whose room % n_shards == shard
    room_of_op = ops["target"].copy()
    k1 = ops["kind"] == 1
    msg_room = room_of_op[k1]
    mine_msg = (msg_room % n_shards) == shard
    mine_op = np.repeat(mine_msg, 3)
    sel = np.nonzero(mine_op)[0]
    lens = np.diff(ops["off"].astype(np.int64))
    sub_texts = [ops["text"][int(ops["off"][i]):int(ops["off"][i + 1])].tobytes() for i in sel]
    text, off = O.pack(sub_texts)
    msg_idx = np.nonzero(mine_msg)[0]
    remap = -np.ones(len(mine_msg), np.int64)
    remap[msg_idx] = np.arange(len(msg_idx))
    sub = dict(text=text, off=off, kind=ops["kind"][sel], target=ops["target"][sel], except_user=ops["except_user"][sel],
               flags=ops["flags"][sel], gate=remap[ops["gate"][sel]].astype(np.int32))
    bsel = [bt[int(bo[i]):int(bo[i + 1])].tobytes() for i in msg_idx]
    b2, bo2 = O.pack(bsel)
    st, so = inp["sites"]
    nt, no = inp["names"]
    qs = [st[int(so[i]):int(so[i + 1])].tobytes() for i in range(shard, n_ban, n_shards)]
    qn = [nt[int(no[i]):int(no[i + 1])].tobytes() for i in range(shard, n_ban, n_shards)]
    qst, qso = O.pack(qs)
    qnt, qno = O.pack(qn)
    t0 = time.perf_counter()
    if R is not None:
        R.set_swear_words(words[:-1])
        R.set_ban_file(0, inp["sfile"]); R.set_ban_file(1, inp["ufile"])
        v = R.contains_swearing_batch(b2, bo2)
        R.ban_batch(0, qst, qso); R.ban_batch(1, qnt, qno)
        R.write_batch(sub, inp["n_rooms"], users, verdict=v, sink_mode=1)
        deliveries = None
        nbytes = int(R.lib.ref_total_write_bytes())
        dt = time.perf_counter() - t0
        deliveries, _ = P.write_batch_count(sub, users, verdict=v)      # counted outside the timed region
    else:
        v = P.contains_swearing_batch(b2, bo2, words)
        P.ban_batch(0, inp["sfile"], qst, qso); P.ban_batch(1, inp["ufile"], qnt, qno)
        deliveries, nbytes = P.write_batch_count(sub, users, verdict=v)
        dt = time.perf_counter() - t0
    return deliveries, nbytes, dt, kind


def reference_step(rank: int, n_msgs: int, n_ban: int, procs: int):
    """Runs the reference's CPU path over a bounded sample with `procs` host processes."""
    import multiprocessing as mp
    import oracle_lib as O
    O.port(); O.ref()                       # build / load the checkers once, before forking
    args = [(rank, s, procs, n_msgs, n_ban) for s in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        res = [_ref_worker(args[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_ref_worker, args)
    wall = time.perf_counter() - t0
    d = sum(r[0] for r in res)
    busy = max(r[2] for r in res)
    return d, sum(r[1] for r in res), busy, wall, res[0][3]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    procs = max(1, os.cpu_count() or 1)
    procs = min(procs, 64)
    n_msgs = min(20_000 * max(1, procs // 4), N_MSGS)
    n_ban = max(procs, n_msgs * N_BAN_QUERIES // N_MSGS)        # the sample keeps the step's msgs : ban-queries ratio
    times, deliv = [], 0
    for i in range(args.warmup + args.steps):
        d, nbytes, busy, wall, kind = reference_step(0, n_msgs, n_ban, procs)
        if i >= args.warmup:
            times.append(busy); deliv += d
    total = sum(times)
    value = deliv / total
    sample = ("%d of %d msgs (all %d users, same generator) + %d of %d ban queries per list, per step; "
              "%d single-threaded reference processes sharded by room; write(2) hooked to a byte counter"
              % (n_msgs, N_MSGS, N_USERS, n_ban, N_BAN_QUERIES, procs))
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * total / max(1, args.steps), higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="u8", data="synthetic",
                config=dict(workload=workload_name(), note="reference CPU path on a bounded sample of the workload"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=procs, kind=kind, sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from nuts333_b200 import api, build

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    build.build()
    dev = torch.device("cuda", local)

    inp = make_inputs(rank, N_MSGS)
    ops, users = inp["ops"], inp["users"]
    ctx = api.Context(local)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_profiling(True)
    ctx.set_swear_words(inp["words"])
    ctx.set_ban_files(inp["sfile"], inp["ufile"])
    ctx.set_users(users["room"], users["flags"], users["level"], inp["n_rooms"])

    def to_dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def pad16(a):                      # packed text buffers are read in 16-byte vectors
        return np.concatenate([a, np.zeros(32, np.uint8)])

    bt, bo = inp["bodies"]
    st_, so_ = inp["sites"]
    nt_, no_ = inp["names"]
    d = dict(bt=to_dev(pad16(bt)), bo=to_dev(bo.view(np.int64)), text=to_dev(pad16(ops["text"])),
             off=to_dev(ops["off"].view(np.int64)), kind=to_dev(ops["kind"]), target=to_dev(ops["target"]),
             exc=to_dev(ops["except_user"]), flags=to_dev(ops["flags"]), gate=to_dev(ops["gate"]),
             st=to_dev(pad16(st_)), so=to_dev(so_.view(np.int64)), nt=to_dev(pad16(nt_)), no=to_dev(no_.view(np.int64)),
             verdict=torch.zeros(N_MSGS, dtype=torch.uint8, device=dev),
             vs=torch.zeros(N_BAN_QUERIES, dtype=torch.uint8, device=dev),
             vu=torch.zeros(N_BAN_QUERIES, dtype=torch.uint8, device=dev))
    n_ops = len(ops["kind"])
    input_bytes = sum(int(v.numel() * v.element_size()) for k, v in d.items() if k not in ("verdict", "vs", "vu"))
    torch.cuda.synchronize()

    state = dict(deliv=0, bytes=0, launches=0, fan_ms=0.0, fan_in=0, fan_out=0, plan_ms=0.0, direct_ms=0.0, render_ms=0.0, ksteps=0)
    state_lock = threading.Lock()

    # Batches in flight on this GPU: every lane is a context of its own (own stream, own scratch and stream
    # buffers) fed by its own host thread; the inputs in HBM are shared, the verdict outputs are per lane.
    # With two lanes one batch's planning / rendering (issue-bound) runs beside the other's fan-out (HBM-bound).
    lanes = [dict(ctx=ctx, stream=stream, verdict=d["verdict"], vs=d["vs"], vu=d["vu"])]

    def add_lane():
        c2 = api.Context(local)
        s2 = torch.cuda.Stream(device=dev)
        c2.set_stream(s2.cuda_stream); c2.set_profiling(True)
        c2.set_swear_words(inp["words"]); c2.set_ban_files(inp["sfile"], inp["ufile"])
        c2.set_users(users["room"], users["flags"], users["level"], inp["n_rooms"])
        lanes.append(dict(ctx=c2, stream=s2, verdict=torch.zeros_like(d["verdict"]), vs=torch.zeros_like(d["vs"]),
                          vu=torch.zeros_like(d["vu"])))

    for _ in range(1, max(1, args.in_flight)):
        add_lane()

    def step_device(L=None):
        L = L or lanes[0]
        c = L["ctx"]
        c.verdicts_dev("contains_swearing", N_MSGS, d["bt"].data_ptr(), d["bo"].data_ptr(), L["verdict"].data_ptr())
        c.verdicts_dev("site_banned", N_BAN_QUERIES, d["st"].data_ptr(), d["so"].data_ptr(), L["vs"].data_ptr())
        c.verdicts_dev("user_banned", N_BAN_QUERIES, d["nt"].data_ptr(), d["no"].data_ptr(), L["vu"].data_ptr())
        s = c.write_batch_dev(n_ops, d["text"].data_ptr(), d["off"].data_ptr(), d["kind"].data_ptr(),
                              d["target"].data_ptr(), d["exc"].data_ptr(), d["flags"].data_ptr(),
                              d["gate"].data_ptr(), L["verdict"].data_ptr())
        t = c.timing()
        with state_lock:
            state["deliv"] += int(s.n_deliveries); state["bytes"] += int(s.total_bytes)
            state["launches"] += int(t.launches) + 3
            state["fan_ms"] += float(t.fanout_ms); state["fan_in"] += int(t.fanout_bytes_in); state["fan_out"] += int(t.fanout_bytes_out)
            state["plan_ms"] += float(t.plan_ms); state["direct_ms"] += float(t.direct_ms); state["render_ms"] += float(t.render_ms)
            state["ksteps"] += 1
        return s

    def run_steps(k):
        """exactly k steps, dealt to the lanes"""
        if len(lanes) == 1:
            for _ in range(k):
                step_device()
            return
        share = [k // len(lanes) + (1 if i < k % len(lanes) else 0) for i in range(len(lanes))]
        th = [threading.Thread(target=lambda L=L, m=m: [step_device(L) for _ in range(m)]) for L, m in zip(lanes, share)]
        for t in th:
            t.start()
        for t in th:
            t.join()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if os.environ.get("BENCH_CLOCKS", "rank0") == "all" or rank == 0:    # rank 0 prints the line; nvidia-smi loops on every
        clocks.start()                      # nvidia-smi needs ~0.1 s to start: launched before the warm-up
    run_steps(args.warmup * len(lanes))
    for k in state:
        state[k] = 0 if not isinstance(state[k], float) else 0.0
    barrier()
    marker = torch.cuda.Stream(device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    e0.record(marker)                   # the device is idle here (barrier above): the lanes' first kernels come after
    run_steps(args.steps)
    for L in lanes:
        marker.wait_stream(L["stream"])
    e1.record(marker)
    barrier()
    tw1 = time.perf_counter()
    clocks.window(tw0, tw1)
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    dev_state = dict(state)
    # ---- per-kernel durations: the same step with every kernel on one stream (no k_render / k_direct beside
    #      the plan / the fan-out), so that each kernel is timed alone with CUDA events on its own stream
    ctx.set_overlap(False)
    step_device()                       # warm-up of the serial schedule
    for k in state:
        state[k] = 0 if not isinstance(state[k], float) else 0.0
    for _ in range(KERNEL_TIMING_STEPS):
        step_device()
    torch.cuda.synchronize()
    kstate = dict(state)
    ctx.set_overlap(os.environ.get('NUTSB_OVERLAP', '1') != '0')

    # ---- beside the headline (one batch at a time): the same K steps with two batches in flight on the GPU
    #      (a second context and host thread), one batch's planning under the other's fan-out
    two_ms = 0.0
    if len(lanes) == 1 and not args.no_e2e:
        add_lane()
        run_steps(2 * args.warmup)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(marker)
        run_steps(args.steps)
        for L in lanes:
            marker.wait_stream(L["stream"])
        f1.record(marker)
        barrier()
        two_ms = f0.elapsed_time(f1)
        lanes.pop().get("ctx").close()

    # ---- e2e: host buffers through the C-ABI, H2D and D2H inside the timed region.  The step's inputs sit in
    #      pinned host memory (the library copies straight from the caller's buffers); the result lands in the
    #      library's own pinned buffers.
    pinned = []

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        pinned.append(t)
        return t.numpy()

    bt, bo, st_, so_, nt_, no_ = (pin(a) for a in (bt, bo, st_, so_, nt_, no_))
    hops = {k: (pin(v) if isinstance(v, np.ndarray) else v) for k, v in ops.items()}
    # the speech tier takes the input lines themselves (verb, speaker, body): say() composed on the device
    from nuts333_b200 import synth
    un, uo = synth.names(N_USERS)
    ctx.set_user_names([un[int(uo[u]):int(uo[u + 1])].tobytes() for u in range(N_USERS)], np.zeros(N_USERS, np.uint8))
    ctx.set_ban_swearing(True)
    sp_verb, sp_spk = pin(np.zeros(N_MSGS, np.uint8)), pin(inp["speaker"].astype(np.int32))

    def e2e_leg(mode):
        """One leg through the host-buffer entry points: verdict batches, then nutsb_write_batch (mode "streams":
        streams in pinned host memory), nutsb_write_batch_iov ("iov": gather lists into the pool of renderings) or
        nutsb_speech_batch_iov ("speech_iov": the 1M input lines of say() in -- swear verdicts, composition and
        rendering on the device -- gather lists out)."""
        iov = mode != "streams"
        r = dict(h2d_ms=0.0, d2h_ms=0.0, ms=0.0, deliv=0, h2d=0, d2h=0, extra={})
        for i in range(1 + e2e_steps):                     # first pass allocates the pinned result buffers
            barrier()
            t0 = time.perf_counter()
            vs = ctx.site_banned_batch(st_, so_)
            vu = ctx.user_banned_batch(nt_, no_)
            keep = []
            if mode == "speech_iov":
                v = np.zeros(0, np.uint8)                  # the verdicts stay on the device
                s = api._IovStreams()
                ctx._ck(ctx.lib.nutsb_speech_batch_iov(ctx._h, N_MSGS, api._addr(sp_verb), api._addr(sp_spk), api._addr(bt),
                                                       api._addr(bo), s))
            else:
                v = ctx.contains_swearing_batch(bt, bo)
                o = ctx._ops_struct(dict(hops, verdict=v), keep)
            if mode == "iov":
                s = api._IovStreams()
                ctx._ck(ctx.lib.nutsb_write_batch_iov(ctx._h, o, s))
            if iov:
                touch = int(np.ctypeslib.as_array(api.C.cast(s.pool, api.u8p), shape=(16,))[0])      # touch the result
                touch += int(np.ctypeslib.as_array(api.C.cast(s.iov, api.u64p), shape=(2,))[1])
                out_bytes = int(s.pool_bytes) + int(s.pool2_bytes) + 16 * int(s.n_iov) + 12 * N_USERS
            else:
                s = api._Streams()
                ctx._ck(ctx.lib.nutsb_write_batch(ctx._h, o, s))
                touch = int(np.ctypeslib.as_array(api.C.cast(s.bytes, api.u8p), shape=(16,))[0])
                out_bytes = int(s.total_bytes)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if i == 0 and iov:
                # outside the timed passes: the lists describe every byte of every stream
                lens = np.ctypeslib.as_array(api.C.cast(s.iov, api.u64p), shape=(int(s.n_iov), 2))[:, 1]
                assert int(lens.sum()) == int(s.total_bytes), "gather lists do not add up to the streams"
                r["extra"] = dict(pool_bytes=int(s.pool_bytes) + int(s.pool2_bytes), n_iov=int(s.n_iov), stream_bytes=int(s.total_bytes))
            if i > 0:
                tt = ctx.timing()
                r["h2d_ms"] += float(tt.h2d_ms); r["d2h_ms"] += float(tt.d2h_ms)
                r["ms"] += dt * 1e3; r["deliv"] += int(s.n_deliveries)
                r["h2d"] = int(bt.nbytes + bo.nbytes + st_.nbytes + so_.nbytes + nt_.nbytes + no_.nbytes)
                r["h2d"] += (5 * N_MSGS if mode == "speech_iov" else
                             int(ops["text"].nbytes + ops["off"].nbytes + 2 * n_ops + 12 * n_ops + v.nbytes))
                r["d2h"] = out_bytes + 8 * (N_USERS + 1) + v.nbytes + vs.nbytes + vu.nbytes
        return r

    e2e_steps = max(1, min(args.steps, 3))
    zero = dict(h2d_ms=0.0, d2h_ms=0.0, ms=0.0, deliv=0, h2d=0, d2h=0, extra={})
    leg_s = zero if args.no_e2e else e2e_leg("streams")
    leg_v = zero if args.no_e2e else e2e_leg("iov")
    leg_p = zero if args.no_e2e else e2e_leg("speech_iov")
    if not args.no_e2e:                                    # the three legs deliver the same streams
        assert leg_v["deliv"] == leg_s["deliv"] == leg_p["deliv"], (leg_s["deliv"], leg_v["deliv"], leg_p["deliv"])
        assert leg_v["extra"]["stream_bytes"] == leg_p["extra"]["stream_bytes"]
    e2e_h2d_ms, e2e_d2h_ms, e2e_ms, e2e_deliv, h2d, d2h = (leg_s[k] for k in ("h2d_ms", "d2h_ms", "ms", "deliv", "h2d", "d2h"))

    # ---- the last leg again with two calls in flight (two contexts, two host threads): one call's H2D and
    #      kernels run under the other's D2H (PCIe is full duplex: scripts/probes/pcie_probe.py)
    pipe_ms, pipe_deliv = 0.0, 0
    if not args.no_e2e:
        c2 = api.Context(local)
        c2.set_swear_words(inp["words"]); c2.set_ban_files(inp["sfile"], inp["ufile"])
        c2.set_users(users["room"], users["flags"], users["level"], inp["n_rooms"])
        c2.set_user_names([un[int(uo[u]):int(uo[u + 1])].tobytes() for u in range(N_USERS)], np.zeros(N_USERS, np.uint8))
        c2.set_ban_swearing(True)
        got = [0, 0]

        def pipe_worker(k, cx, reps):
            for _ in range(reps):
                cx.site_banned_batch(st_, so_); cx.user_banned_batch(nt_, no_)
                s = api._IovStreams()
                cx._ck(cx.lib.nutsb_speech_batch_iov(cx._h, N_MSGS, api._addr(sp_verb), api._addr(sp_spk), api._addr(bt), api._addr(bo), s))
                got[k] += int(s.n_deliveries)

        for reps in (1, e2e_steps):                        # first round allocates the second context's buffers
            got[0] = got[1] = 0
            barrier()
            t0 = time.perf_counter()
            th = [threading.Thread(target=pipe_worker, args=(k, cx, reps)) for k, cx in enumerate((ctx, c2))]
            for t in th:
                t.start()
            for t in th:
                t.join()
            torch.cuda.synchronize()
            pipe_ms = (time.perf_counter() - t0) * 1e3
        pipe_deliv = got[0] + got[1]
        c2.close()
    # ---- reduce over ranks
    vals = torch.tensor([ms, e2e_ms, leg_v["ms"], leg_p["ms"], two_ms, pipe_ms], dtype=torch.float64, device=dev)
    sums = torch.tensor([dev_state["deliv"], e2e_deliv, dev_state["launches"], leg_v["deliv"], pipe_deliv], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms_max, e2e_ms_max, iov_ms_max, sp_ms_max, two_ms_max, pipe_ms_max = (float(vals[i]) for i in range(6))
    total_deliv, total_e2e_deliv, total_launch, total_iov_deliv = float(sums[0]), float(sums[1]), int(sums[2]), float(sums[3])

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        ks = max(1, kstate["ksteps"])
        fan_bytes = (kstate["fan_in"] + kstate["fan_out"]) / ks
        fan_ms = kstate["fan_ms"] / ks
        achieved = fan_bytes / (fan_ms * 1e-3) / 1e9 if fan_ms > 0 else 0.0
        traffic = None
        tp = ROOT / "profiles" / "fanout_traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = dict(metric=METRIC, value=total_deliv / (ms_max * 1e-3), unit=UNIT, n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms_max / max(1, args.steps), higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="u8", data="synthetic",
                    config=dict(workload=workload_name(), sharding="rooms per rank, no collective",
                                l2="inputs (%.0f MB) and outputs (%.1f GB) per step exceed the 126 MB L2; no flush needed"
                                   % (input_bytes / 1e6, dev_state["bytes"] / max(1, args.steps) / 1e9),
                                batches_in_flight=len(lanes),
                                two_batches_in_flight=(None if two_ms_max <= 0 else dict(
                                    value=total_deliv / (two_ms_max * 1e-3),
                                    ms_per_step=two_ms_max / max(1, args.steps),
                                    note="same K steps dealt to two contexts / host threads on the GPU; not the headline")),
                                source_msgs_per_s=world * N_MSGS * args.steps / (ms_max * 1e-3),
                                ban_queries_per_step=2 * N_BAN_QUERIES,
                                kernel_ms_alone=dict(plan=kstate["plan_ms"] / ks, render=kstate["render_ms"] / ks, fanout=fan_ms,
                                                     direct=kstate["direct_ms"] / ks, steps=ks,
                                                     how="extra steps after the timed region with nutsb_set_overlap(0): "
                                                         "every kernel on one stream, CUDA events around each"),
                                timed_region_ms=dict(plan_until_fanout=dev_state["plan_ms"] / max(1, dev_state["ksteps"]),
                                                     render_span=dev_state["render_ms"] / max(1, dev_state["ksteps"]),
                                                     fanout_and_direct=dev_state["fan_ms"] / max(1, dev_state["ksteps"]),
                                                     how="CUDA events of the timed steps themselves (side stream on)"),
                                plan_ms=kstate["plan_ms"] / ks, render_ms=kstate["render_ms"] / ks, fanout_ms=fan_ms,
                                direct_ms=kstate["direct_ms"] / ks),
                    roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                                  traffic=traffic, kernel="k_fanout", peak_source=peak_src,
                                  algorithmic_bytes_per_launch=fan_bytes,
                                  frac_of_spec_8000=achieved / 8000.0,          # both percentages
                                  whole_step=dict(achieved=(dev_state["bytes"] / max(1, args.steps) + input_bytes) / (ms_max / max(1, args.steps) * 1e-3) / 1e9,
                                                  unit="GB/s", note="stream bytes written + input bytes read per step / ms_per_step (rank 0): "
                                                                    "what the planning, rendering and verdict kernels cost on top of the fan-out")),
                    e2e=(None if args.no_e2e else
                         dict(value=total_e2e_deliv / (e2e_ms_max * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d,
                              d2h_bytes_per_step=d2h, steps=e2e_steps, ms_per_step=e2e_ms_max / e2e_steps,
                              write_batch_h2d_ms=e2e_h2d_ms / e2e_steps, write_batch_d2h_ms=e2e_d2h_ms / e2e_steps,
                              host_memory="pinned (inputs: caller's pinned buffers; streams: the library's pinned buffer)")),
                    gpu_launches=total_launch, clocks=clk)
        if not args.no_e2e:
            # the same step with gather lists as the result (nutsb_write_batch_iov): what a host that writev()s needs.
            # Reported beside `e2e`, which stays the full per-recipient streams.
            line["e2e_iov"] = dict(value=total_iov_deliv / (iov_ms_max * 1e-3), unit=UNIT, h2d_bytes_per_step=leg_v["h2d"],
                                   d2h_bytes_per_step=leg_v["d2h"], steps=e2e_steps, ms_per_step=iov_ms_max / e2e_steps,
                                   write_batch_h2d_ms=leg_v["h2d_ms"] / e2e_steps, write_batch_d2h_ms=leg_v["d2h_ms"] / e2e_steps,
                                   result="per-user struct iovec lists into one pinned pool: every room op rendered once per "
                                          "colour setting + every write_user op rendered once; byte-identical to `e2e`'s streams "
                                          "when gathered (tests/test_iov.py)", **leg_v["extra"])
            # ... and from the input lines themselves: say(user, inpstr) x 1M through nutsb_speech_batch_iov
            line["e2e_speech_iov"] = dict(value=total_iov_deliv / (sp_ms_max * 1e-3), unit=UNIT, h2d_bytes_per_step=leg_p["h2d"],
                                          d2h_bytes_per_step=leg_p["d2h"], steps=e2e_steps, ms_per_step=sp_ms_max / e2e_steps,
                                          input="verb, speaker and body of every line (nutsb_speech_batch_iov): swear verdicts, "
                                                "say()'s composition and the rendering all on the device", **leg_p["extra"])
            line["e2e_speech_iov"]["two_calls_in_flight"] = dict(
                value=float(sums[4]) / (pipe_ms_max * 1e-3), ms_per_step=pipe_ms_max / (2 * e2e_steps),
                note="two contexts / host threads per GPU, %d calls each: one call's H2D and kernels under the other's D2H" % e2e_steps)
        if world == 1 and not args.no_cpu_baseline:
            procs = 1
            dcpu, nbytes, busy, wall, kind = reference_step(0, 40_000, 4_000, procs)
            line["cpu_baseline"] = dict(value=dcpu / busy, unit=UNIT, cores=procs, kind=kind,
                                        sample="40000 of 1000000 msgs (all 10000 users) + 4000 of 100000 ban queries per "
                                               "list, one single-threaded reference process, write(2) hooked to a byte counter")
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    for L in lanes:
        L["ctx"].close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--in-flight", type=int, default=IN_FLIGHT, help="batches in flight per GPU (one context + host thread each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
