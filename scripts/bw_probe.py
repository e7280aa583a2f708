"""Pure-write vs copy bandwidth probe (context for the fan-out roofline: that kernel is ~99% writes)."""
import torch, json
dev = torch.device("cuda", 0)
n = 6 * 1024**3
a = torch.empty(n, dtype=torch.uint8, device=dev)
b = torch.empty(n, dtype=torch.uint8, device=dev)
res = {}
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
t = timeit(lambda: a.fill_(7)); res["fill_u8_GBps"] = n / t / 1e6
a32 = a.view(torch.int32)
t = timeit(lambda: a32.fill_(7)); res["fill_i32_GBps"] = n / t / 1e6
t = timeit(lambda: a.zero_()); res["memset_zero_GBps"] = n / t / 1e6
t = timeit(lambda: b.copy_(a)); res["copy_rw_GBps"] = 2 * n / t / 1e6
t = timeit(lambda: a.sum()); res["read_sum_GBps"] = n / t / 1e6
print(json.dumps(res))
