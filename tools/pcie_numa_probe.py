"""Host<->device copy bandwidth per NUMA node of the pinned buffer (probe for the e2e path; not part of the product)."""
import glob, os, time
import torch

def node_cpus():
    out = {}
    for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        n = int(d.rsplit("node", 1)[1]); cp = set()
        for part in open(d + "/cpulist").read().strip().split(","):
            if not part: continue
            a, _, b = part.partition("-"); cp.update(range(int(a), int(b or a) + 1))
        out[n] = cp
    return out

def gpu_node(i=0):
    try:
        bus = torch.cuda.get_device_properties(i).pci_bus_id
        dom = torch.cuda.get_device_properties(i).pci_domain_id
        devid = torch.cuda.get_device_properties(i).pci_device_id
        p = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        return int(open(p).read()), p
    except Exception as ex:
        return None, str(ex)

print("gpu numa:", gpu_node(), "nodes:", {k: len(v) for k, v in node_cpus().items()})
N = 1 << 30
dev = torch.empty(N, dtype=torch.uint8, device="cuda"); dev2 = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
all_cpus = os.sched_getaffinity(0)
for node, cpus in node_cpus().items():
    cp = cpus & all_cpus
    if not cp: continue
    os.sched_setaffinity(0, cp)
    h1 = torch.empty(N, dtype=torch.uint8).pin_memory(); h1.fill_(1)
    h2 = torch.empty(N, dtype=torch.uint8).pin_memory(); h2.fill_(2)
    res = []
    for mode in ("h2d", "d2h", "both"):
        best = 0
        for it in range(4):
            torch.cuda.synchronize(); t = time.perf_counter()
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1): dev.copy_(h1, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2): h2.copy_(dev2, non_blocking=True)
            torch.cuda.synchronize(); dt = time.perf_counter() - t
            best = max(best, (2 if mode == "both" else 1) * N / dt / 1e9)
        res.append(f"{mode} {best:.1f} GB/s")
    print(f"node {node}: " + "  ".join(res))
    del h1, h2
os.sched_setaffinity(0, all_cpus)
