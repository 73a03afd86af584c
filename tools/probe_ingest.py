"""Development aid (SURVEY.md 8f-3 / VERDICT item 8): what limits the host-frame path on this box?
  * is an NVDEC user-mode library present (libnvcuvid.so) -- without it there is no on-GPU video decode to build on;
  * PCIe topology / NUMA as nvidia-smi and /sys report them;
  * H2D bandwidth of ONE cudaMemcpyAsync per batch from pinned memory: one GPU alone, and all visible GPUs at once
    (one thread + one pinned pool + one stream per GPU), which is what the multi-GPU e2e numbers run into.
"""
import ctypes, glob, json, os, subprocess, sys, threading, time

out = {}
libs = []
for pat in ("/usr/lib/x86_64-linux-gnu/libnvcuvid*", "/usr/lib64/libnvcuvid*", "/usr/local/cuda/lib64/libnvcuvid*", "/usr/lib/x86_64-linux-gnu/libnvidia-encode*"):
    libs += glob.glob(pat)
out["nvdec_libs"] = libs
try:
    ctypes.CDLL("libnvcuvid.so.1")
    out["libnvcuvid_loadable"] = True
except OSError as e:
    out["libnvcuvid_loadable"] = False
    out["libnvcuvid_error"] = str(e)
for name, cmd in (("topo", ["nvidia-smi", "topo", "-m"]), ("lscpu", ["bash", "-c", "lscpu | egrep 'NUMA|Socket|Model name|^CPU\\(s\\)'"]),
                  ("pcie", ["bash", "-c", "nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max,pcie.link.width.max --format=csv"])):
    try:
        out[name] = subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout
    except Exception as e:
        out[name] = repr(e)
import torch
n = torch.cuda.device_count()
SZ = 60 * 3840 * 2160 * 3          # one bench batch
host = [torch.empty(SZ, dtype=torch.uint8).pin_memory() for _ in range(n)]
dev = [torch.empty(SZ, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]
streams = [torch.cuda.Stream(device=i) for i in range(n)]


def copy_loop(i, reps, res):
    torch.cuda.set_device(i)
    with torch.cuda.stream(streams[i]):
        dev[i].copy_(host[i], non_blocking=True)
        streams[i].synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            dev[i].copy_(host[i], non_blocking=True)
        streams[i].synchronize()
        res[i] = reps * SZ / (time.perf_counter() - t0) / 1e9


res = {}
copy_loop(0, 6, res)
out["h2d_gbs_gpu0_alone"] = res[0]
if n > 1:
    for group in ([0, 1], list(range(min(n, 4))), list(range(n))):
        res = {}
        th = [threading.Thread(target=copy_loop, args=(i, 6, res)) for i in group]
        [t.start() for t in th]; [t.join() for t in th]
        out[f"h2d_gbs_concurrent_{len(group)}"] = {"per_gpu": [round(res[i], 1) for i in group], "sum": round(sum(res.values()), 1)}
print(json.dumps(out, indent=1))
