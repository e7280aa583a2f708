"""PCIe copy rates on the box: pinned host <-> device, one copy against 2 / 4 concurrent copies on separate
streams (do several copy engines add up?), and both directions at once."""
import json, time, torch
dev = torch.device("cuda", 0)
N = 1 << 30
d = torch.empty(N, dtype=torch.uint8, device=dev)
h = torch.empty(N, dtype=torch.uint8).pin_memory()
out = {}
def run(name, parts, direction, both=False):
    streams = [torch.cuda.Stream(device=dev) for _ in range(parts * (2 if both else 1))]
    chunk = N // parts
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(parts):
            with torch.cuda.stream(streams[i]):
                a, b = d[i * chunk:(i + 1) * chunk], h[i * chunk:(i + 1) * chunk]
                if direction == "d2h": b.copy_(a, non_blocking=True)
                else: a.copy_(b, non_blocking=True)
        if both:
            for i in range(parts):
                with torch.cuda.stream(streams[parts + i]):
                    a, b = d2[i * chunk:(i + 1) * chunk], h2[i * chunk:(i + 1) * chunk]
                    a.copy_(b, non_blocking=True)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    out[name] = round(N * (2 if both else 1) / best / 1e9, 2)
d2 = torch.empty(N, dtype=torch.uint8, device=dev); h2 = torch.empty(N, dtype=torch.uint8).pin_memory()
for p in (1, 2, 4):
    run(f"d2h_{p}_streams_GBps", p, "d2h")
    run(f"h2d_{p}_streams_GBps", p, "h2d")
run("d2h_plus_h2d_GBps_total", 1, "d2h", both=True)
print(json.dumps(out))
