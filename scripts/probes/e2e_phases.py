"""Where the host-buffer legs spend their time: wall clock around each C-ABI call of one e2e step
(bench.py's workload, one GPU), next to the library's own device-side timings."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np, torch
import bench
from nuts333_b200 import api, build, synth

build.build()
inp = bench.make_inputs(0, bench.N_MSGS)
ops, users = inp["ops"], inp["users"]
ctx = api.Context(0); ctx.set_profiling(True)
ctx.set_swear_words(inp["words"]); ctx.set_ban_files(inp["sfile"], inp["ufile"])
ctx.set_users(users["room"], users["flags"], users["level"], inp["n_rooms"])
un, uo = synth.names(bench.N_USERS)
ctx.set_user_names([un[int(uo[u]):int(uo[u + 1])].tobytes() for u in range(bench.N_USERS)], np.zeros(bench.N_USERS, np.uint8))
ctx.set_ban_swearing(True)
keepalive = []
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); keepalive.append(t); return t.numpy()
bt, bo = (pin(a) for a in inp["bodies"]); st_, so_ = (pin(a) for a in inp["sites"]); nt_, no_ = (pin(a) for a in inp["names"])
hops = {k: (pin(v) if isinstance(v, np.ndarray) else v) for k, v in ops.items()}
verb, spk = pin(np.zeros(bench.N_MSGS, np.uint8)), pin(inp["speaker"].astype(np.int32))
out = {}
def timed(name, fn, reps=4):
    ts = []
    for i in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    out[name] = dict(ms=min(ts[1:]), all=[round(t, 3) for t in ts])
    return r
v = timed("contains_swearing_batch", lambda: ctx.contains_swearing_batch(bt, bo))
timed("site_banned_batch", lambda: ctx.site_banned_batch(st_, so_))
timed("user_banned_batch", lambda: ctx.user_banned_batch(nt_, no_))
def wb_iov():
    keep = []; o = ctx._ops_struct(dict(hops, verdict=v), keep); s = api._IovStreams()
    ctx._ck(ctx.lib.nutsb_write_batch_iov(ctx._h, o, s)); return s
timed("write_batch_iov", wb_iov)
t = ctx.timing(); out["write_batch_iov"]["lib"] = dict(h2d=t.h2d_ms, d2h=t.d2h_ms, dev_total=t.total_ms, plan=t.plan_ms, direct=t.direct_ms, render=t.render_ms, launches=t.launches)
def sp_iov():
    s = api._IovStreams()
    ctx._ck(ctx.lib.nutsb_speech_batch_iov(ctx._h, bench.N_MSGS, api._addr(verb), api._addr(spk), api._addr(bt), api._addr(bo), s)); return s
timed("speech_batch_iov", sp_iov)
t = ctx.timing(); out["speech_batch_iov"]["lib"] = dict(h2d=t.h2d_ms, d2h=t.d2h_ms, dev_total=t.total_ms, plan=t.plan_ms, direct=t.direct_ms, render=t.render_ms, launches=t.launches)
print(json.dumps(out, indent=1))
