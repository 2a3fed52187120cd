"""Pinned-memory H2D / D2H bandwidth per host NUMA node (affinity set before the pinned allocation)."""
import os, glob, json, subprocess, sys
import torch

def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part: continue
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out

p = torch.cuda.get_device_properties(0)
bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
try:
    gpu_node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
except Exception as e:
    gpu_node = f"? ({e})"
print("gpu", bdf, "numa_node", gpu_node, "cpus visible", len(os.sched_getaffinity(0)))
try:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout[:1500])
except Exception as e:
    print("topo failed", e)
all_cpus = sorted(os.sched_getaffinity(0))
nodes = {}
for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    n = int(d.rsplit("node", 1)[1])
    cpus = [c for c in cpulist(open(d + "/cpulist").read()) if c in all_cpus]
    if cpus: nodes[n] = cpus
print("nodes", {n: f"{c[0]}..{c[-1]} ({len(c)})" for n, c in nodes.items()})
dev = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
def bw(host):
    res = {}
    for name, fn in (("h2d", lambda: dev.copy_(host, non_blocking=True)), ("d2h", lambda: host.copy_(dev, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10): fn()
        e.record(); torch.cuda.synchronize()
        res[name] = round(10 * host.numel() / s.elapsed_time(e) / 1e6, 1)
    return res
for n, cpus in nodes.items():
    os.sched_setaffinity(0, cpus)
    host = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    print(json.dumps({"node": n, "GB/s": bw(host)}))
    del host
os.sched_setaffinity(0, all_cpus)
host = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
print(json.dumps({"node": "any", "GB/s": bw(host)}))
# write-combined pinned memory through cudart
rt = torch.cuda.cudart()
try:
    import ctypes
    lib = ctypes.CDLL("libcudart.so.12")
    ptr = ctypes.c_void_p()
    rc = lib.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(64 << 20), ctypes.c_uint(4))   # cudaHostAllocWriteCombined
    print("cudaHostAlloc WC rc", rc)
    if rc == 0:
        buf = (ctypes.c_uint8 * (64 << 20)).from_address(ptr.value)
        ctypes.memset(ptr, 1, 64 << 20)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lib.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3): lib.cudaMemcpyAsync(dev.data_ptr(), ptr, 64 << 20, 1, st)
        torch.cuda.synchronize()
        s.record()
        for _ in range(10): lib.cudaMemcpyAsync(dev.data_ptr(), ptr, 64 << 20, 1, st)
        e.record(); torch.cuda.synchronize()
        print(json.dumps({"write_combined_h2d_GB/s": round(10 * (64 << 20) / s.elapsed_time(e) / 1e6, 1)}))
except Exception as ex:
    print("WC test failed:", ex)
