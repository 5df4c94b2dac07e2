"""Host-side probe for the multi-rank e2e leg: NUMA nodes, the cpuset's allowed CPUs / memory nodes, the GPU's NUMA node, and the
pinned-copy rate from memory bound to each allowed node (set_mempolicy + first touch, then cudaHostRegister through torch)."""
import ctypes, os, time, glob
import torch

print("cpus allowed:", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)))
for l in open("/proc/self/status"):
    if l.startswith(("Mems_allowed_list", "Cpus_allowed_list")):
        print(l.strip())
nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
print("nodes:", nodes)
for n in nodes:
    print(" node", n, "cpus", open(f"/sys/devices/system/node/node{n}/cpulist").read().strip(),
          [l.strip() for l in open(f"/sys/devices/system/node/node{n}/meminfo") if "MemTotal" in l or "MemFree" in l])
os.system("nvidia-smi topo -m 2>&1 | head -20")
os.system("nvidia-smi --query-gpu=index,pci.bus_id --format=csv 2>&1")
for d in glob.glob("/sys/bus/pci/devices/*/numa_node"):
    try:
        if open(d.replace("numa_node", "vendor")).read().strip() == "0x10de" and open(d.replace("numa_node", "class")).read().startswith("0x0302"):
            print(d, open(d).read().strip())
    except OSError:
        pass
libc = ctypes.CDLL(None, use_errno=True)
SYS_set_mempolicy = 238
MPOL_DEFAULT, MPOL_BIND = 0, 2
dev = torch.device("cuda", 0)
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device=dev)
for node in nodes + [None]:
    if node is not None:
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(SYS_set_mempolicy, MPOL_BIND, ctypes.byref(mask), 64)
        if rc != 0:
            print("set_mempolicy node", node, "failed errno", ctypes.get_errno()); continue
    else:
        libc.syscall(SYS_set_mempolicy, MPOL_DEFAULT, None, 0)
    h = torch.empty(n, dtype=torch.uint8)
    h.fill_(1)                      # first touch under the policy
    torch.cuda.cudart().cudaHostRegister(h.data_ptr(), n, 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("memory on node", node, ": H2D %.1f GB/s" % (5 * n / dt / 1e9))
    torch.cuda.cudart().cudaHostUnregister(h.data_ptr())
    del h
