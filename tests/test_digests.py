"""The parity digests of SURVEY.md 8(d): d = FNV-1a of one delivery's bytes; per user the fold over his deliveries in
call order, per op the fold over its recipients in user-list order.  The restatement is pinned to the reference's own
write(2) calls (ref_harness folds them as they happen); the device computes them from its slab and streams."""
import numpy as np
import pytest

from nuts333_b200 import api
from test_multi import make_case


def _case(seed, U, NR, N):
    c = make_case(seed, U, NR, N)
    # (the reference has no gate: a gated-off op is simply not called) -- keep the gates, both sides honour them
    return c


def test_digests_port_vs_reference(port, ref):
    for seed in (81, 82):
        c = _case(seed, 30, 4, 200)
        o, us = c["ops"], c["users"]
        pu, po = port.delivery_digests(o, us, verdict=o["verdict"])
        ru, ro = ref.delivery_digests(o, c["n_rooms"], us, verdict=o["verdict"])
        assert (pu == ru).all() and (po == ro).all()
        assert len(set(po.tolist())) > 50 and (po != 0).sum() > 100


def _check_device(ctx, port, seed, U, NR, N):
    c = _case(seed, U, NR, N)
    o, us = c["ops"], c["users"]
    pu, po = port.delivery_digests(o, us, verdict=o["verdict"])
    ctx.set_users(us["room"], us["flags"], us["level"], c["n_rooms"])
    ctx.write_batch(o)
    du, do = ctx.delivery_digests(len(o["kind"]))
    bad_u, bad_o = np.nonzero(du != pu)[0], np.nonzero(do != po)[0]
    assert bad_u.size == 0 and bad_o.size == 0, (bad_u[:5], bad_o[:5], [int(o["kind"][i]) for i in bad_o[:5]])


def test_digests_on_emulator(sim_lib, port):
    ctx = api.Context(0, sim_lib)
    _check_device(ctx, port, 83, 24, 3, 150)
    ctx.close()


@pytest.mark.gpu
def test_digests_on_gpu(gpu_ctx, port):
    _check_device(gpu_ctx, port, 84, 600, 9, 6000)
    _check_device(gpu_ctx, port, 85, 50, 1, 3000)
