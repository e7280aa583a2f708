"""nutsb_pipe: host-buffer calls in flight on one device (two contexts, a worker thread each).  Every call's gather
lists must describe the same streams as the call made directly -- whatever else is in flight."""
import numpy as np
import pytest

import oracle_lib as O
from nuts333_b200 import api, synth


def _check(lib, port, device, n_msgs, n_users, upr, rounds):
    words = synth.swear_words(64)
    us, n_rooms = synth.users(n_users, upr)
    un, uo = synth.names(n_users)
    names = [un[int(uo[u]):int(uo[u + 1])].tobytes() for u in range(n_users)]
    p = api.Pipe(device, 2, lib)
    p.set_swear_words(words); p.set_users(us["room"], us["flags"], us["level"], n_rooms)
    p.set_user_names(names, np.zeros(n_users, np.uint8)); p.set_ban_swearing(True)
    batches, want = [], []
    for k in range(rounds):
        bt, bo = synth.bodies(n_msgs, words, m0=k * n_msgs)
        sops, spk, rm = synth.say_ops(n_msgs, n_users, upr, bt, bo, gated=True, m0=k * n_msgs)
        v = port.contains_swearing_batch(bt, bo, words)
        off, data, nd = port.write_batch(sops, us, verdict=v)
        batches.append((np.zeros(n_msgs, np.uint8), spk.astype(np.int32), bt, bo, dict(sops, verdict=v)))
        want.append((off, data))
    # speech lines in flight: submit k+1 before waiting for k
    t_prev = p.submit_speech_iov(*batches[0][:4])
    for k in range(1, rounds + 1):
        t_next = p.submit_speech_iov(*batches[k][:4]) if k < rounds else None
        r = p.wait(t_prev)
        off, data = want[k - 1]
        assert (r.off == off).all()
        for u in range(0, n_users, max(1, n_users // 23)):
            assert r.user(u) == data[int(off[u]):int(off[u + 1])].tobytes(), (k, u)
        t_prev = t_next
    # composed ops through the same pipe
    ta = p.submit_write_iov(batches[0][4]); tb = p.submit_write_iov(batches[1][4])
    for t, k in ((ta, 0), (tb, 1)):
        r = p.wait(t)
        off, data = want[k]
        assert (r.off == off).all() and r.user(1) == data[int(off[1]):int(off[2])].tobytes()
    tc = p.submit_write_iov(batches[2][4])
    with pytest.raises(api.NutsbError):
        p.wait(ta)                                          # given up: its context has taken the next call
    assert (p.wait(tc).off == want[2][0]).all()
    p.close()


def test_pipe_on_emulator(sim_lib, port):
    _check(sim_lib, port, 0, 60, 40, 10, 3)


@pytest.mark.gpu
def test_pipe_on_gpu(gpu_ctx, port):
    _check(gpu_ctx.lib, port, 0, 20000, 1000, 100, 4)
