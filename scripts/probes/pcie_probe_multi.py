"""Aggregate PCIe copy rates with N GPUs copying AT THE SAME TIME (one process per GPU, torchrun): does the box give
every GPU its own 57 GB/s, or do they share a host memory / PCIe path?  (the host-buffer legs of
bench.py lose most of their efficiency at 8 GPUs.)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/probes/pcie_probe_multi.py

Each rank copies 1 GiB pinned host <-> its device, all ranks between two barriers; the aggregate = N x bytes / slowest
rank's time.  Variants: the pinned buffer allocated by a thread bound to the GPU's NUMA node / to the other node
(NUTSB_PROBE_BIND=local|remote|none, read from nvidia-smi topo), write-combined pinned memory for the H2D source."""
import ctypes, json, os, subprocess, sys, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = 1 << 30


def numa_info():
    """NUMA node count and this GPU's node / CPU affinity as the driver reports them."""
    info = dict(nodes=None, gpu_node=None, gpu_cpus=None)
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
        info["nodes"] = len(nodes)
        info["node_cpus"] = {d: open(f"/sys/devices/system/node/{d}/cpulist").read().strip() for d in sorted(nodes)}
    except OSError:
        pass
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip().lower()
        bus = bus[4:] if bus.startswith("0000") and len(bus) > 12 else bus
        p = f"/sys/bus/pci/devices/{bus}/numa_node"
        info["gpu_node"] = int(open(p).read()) if os.path.exists(p) else None
        p = f"/sys/bus/pci/devices/{bus}/local_cpulist"
        info["gpu_cpus"] = open(p).read().strip() if os.path.exists(p) else None
    except Exception:
        pass
    return info


def cpus_of(spec):
    out = set()
    for part in spec.split(","):
        if "-" in part:
            a, b = part.split("-"); out |= set(range(int(a), int(b) + 1))
        elif part.strip():
            out.add(int(part))
    return out


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn):
    best = 1e9
    for _ in range(4):
        barrier(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t[0]))
    return best


info = numa_info()
res = dict(n_gpus=world, numa=info if rank == 0 else None)
d = torch.empty(N, dtype=torch.uint8, device=dev)
d2 = torch.empty(N, dtype=torch.uint8, device=dev)
for bind in ("none", "local", "remote"):
    if bind != "none":
        if not info.get("nodes") or info["nodes"] < 2 or info.get("gpu_node") is None or info["gpu_node"] < 0:
            res[bind] = "no NUMA choice on this box (nodes=%s, gpu_node=%s)" % (info.get("nodes"), info.get("gpu_node"))
            continue
        node = info["gpu_node"] if bind == "local" else (info["gpu_node"] + 1) % info["nodes"]
        try:
            os.sched_setaffinity(0, cpus_of(info["node_cpus"]["node%d" % node]))       # first touch decides where the pages live
        except Exception as e:
            res[bind] = "sched_setaffinity failed: %s" % e
            continue
    h = torch.empty(N, dtype=torch.uint8).pin_memory(); h.fill_(1)
    h2 = torch.empty(N, dtype=torch.uint8).pin_memory(); h2.fill_(2)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def d2h():
        with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)

    def h2d():
        with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)

    def both():
        with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
        with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    r = {}
    for name, fn, mult in (("d2h", d2h, 1), ("h2d", h2d, 1), ("both", both, 2)):
        t = timed(fn)
        r[name + "_GBps_aggregate"] = round(world * mult * N / t / 1e9, 1)
        r[name + "_GBps_per_gpu"] = round(mult * N / t / 1e9, 1)
    res[bind] = r
    del h, h2
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
