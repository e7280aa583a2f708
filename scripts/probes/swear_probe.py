"""contains_swearing alone: the verdict kernel's duration (CUDA events on the library's stream, best of 5, inputs in
HBM) for the stock three-word list (k_ac_pair<true>: two bytes a step) and the 64-word list of config 3
(k_ac_pair<false>), 1M bodies each, verdicts checked against the oracle.  -> gpurun_out/r2_swear_probe.json"""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch
from nuts333_b200 import api, synth, build
import oracle_lib as O

build.build()
N = 1_000_000
port = O.port()
ctx = api.Context(0)
stream = torch.cuda.Stream(device=0)
ctx.set_stream(stream.cuda_stream)
out = {}
for name, words in (("stock_3_words", ["fuck", "shit", "cunt", "*"]), ("config3_64_words", synth.swear_words(64))):
    bt, bo = synth.bodies(N, words)
    ctx.set_swear_words(words)
    dt = torch.from_numpy(np.concatenate([bt, np.zeros(64, np.uint8)])).cuda(); do = torch.from_numpy(bo.view(np.int64)).cuda()
    dv = torch.zeros(N, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); ctx.verdicts_dev("contains_swearing", N, dt.data_ptr(), do.data_ptr(), dv.data_ptr()); b.record(stream)
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    want = port.contains_swearing_batch(bt, bo, words)
    got = dv.cpu().numpy()
    assert (got == want).all(), name
    ms = min(ts[3:])
    out[name] = dict(ms=ms, text_bytes=int(bo[-1]), gb_s=float(bo[-1]) / ms / 1e6, msgs_per_s=N / ms * 1e3, dirty=int(want.sum()), parity="bit-exact vs oracle port, 1M verdicts")
ctx.close()
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "r2_swear_probe.json").write_text(json.dumps(out, indent=1))
print(json.dumps(out))
