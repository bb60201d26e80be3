"""Host <-> device link microbenchmark: what the box's PCIe fabric and host memory give to 1..N GPUs at once.
The end-to-end fit is a 4 GB host-to-device copy with ~10 ms of kernels under it, so these numbers ARE its
ceiling (state the platform ceiling with a measurement, not an inference).
For every device set: aggregate GB/s of concurrent page-locked copies, one stream and one 1 GiB buffer per device.
Also the host's own memcpy bandwidth (1 and all threads) and `nvidia-smi topo -m`.  One JSON line."""
import json, os, subprocess, sys, threading, time
import numpy as np
import torch

nd = torch.cuda.device_count()
GB = 1 << 30
host = [torch.empty(GB // 8, dtype=torch.float64).pin_memory() for _ in range(nd)]
devb = [torch.empty(GB // 8, dtype=torch.float64, device=f"cuda:{d}") for d in range(nd)]
streams = [torch.cuda.Stream(device=d) for d in range(nd)]


def run(devs, to_device=True, reps=3):
    best = 0.0
    for _ in range(reps + 1):
        for d in devs:
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for d in devs:
            with torch.cuda.stream(streams[d]):
                if to_device:
                    devb[d].copy_(host[d], non_blocking=True)
                else:
                    host[d].copy_(devb[d], non_blocking=True)
        for d in devs:
            streams[d].synchronize()
        dt = time.perf_counter() - t0
        best = max(best, len(devs) * GB / dt * 1e-9)
    return best


out = {"devices": nd, "h2d_single": {}, "d2h_single": {}, "h2d_sets": {}, "d2h_sets": {}, "duplex_sets": {}}
for d in range(nd):
    out["h2d_single"][str(d)] = run([d])
    out["d2h_single"][str(d)] = run([d], False)
sets = {"all": list(range(nd))}
if nd >= 2:
    sets["0,1"] = [0, 1]
if nd >= 4:
    sets["0-3"] = [0, 1, 2, 3]
    sets["0,2"] = [0, 2]
    sets["0,4" if nd >= 8 else "0,3"] = [0, 4 if nd >= 8 else 3]
if nd >= 8:
    sets["4-7"] = [4, 5, 6, 7]
    sets["even"] = [0, 2, 4, 6]
    sets["0,1,4,5"] = [0, 1, 4, 5]
for name, devs in sets.items():
    out["h2d_sets"][name] = run(devs)
    out["d2h_sets"][name] = run(devs, False)


def duplex(devs):
    """H2D on every device of the set and D2H on the same devices at the same time (second buffers)."""
    h2 = [torch.empty(GB // 16, dtype=torch.float64).pin_memory() for _ in devs]
    d2 = [torch.empty(GB // 16, dtype=torch.float64, device=f"cuda:{d}") for d in devs]
    s2 = [torch.cuda.Stream(device=d) for d in devs]
    best = 0.0
    for _ in range(3):
        for d in devs:
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for i, d in enumerate(devs):
            with torch.cuda.stream(streams[d]):
                devb[d].copy_(host[d], non_blocking=True)
            with torch.cuda.stream(s2[i]):
                h2[i].copy_(d2[i], non_blocking=True)
        for i, d in enumerate(devs):
            streams[d].synchronize(); s2[i].synchronize()
        dt = time.perf_counter() - t0
        best = max(best, len(devs) * (GB + GB // 2) / dt * 1e-9)
    return best


out["duplex_sets"]["0"] = duplex([0])
out["duplex_sets"]["all"] = duplex(list(range(nd)))

# host memory: memcpy of 1 GiB by 1 thread and by all threads (bytes copied per second, read + write = 2x that)
src = np.ones(GB // 8); dst = np.empty_like(src)
t0 = time.perf_counter(); np.copyto(dst, src); out["host_memcpy_1thread_GBps"] = GB / (time.perf_counter() - t0) * 1e-9
nt = os.cpu_count() or 1
parts = np.array_split(np.arange(src.size), nt)
def cp(ix):
    np.copyto(dst[ix[0]:ix[-1] + 1], src[ix[0]:ix[-1] + 1])
best = 0.0
for _ in range(3):
    th = [threading.Thread(target=cp, args=(ix,)) for ix in parts]
    t0 = time.perf_counter()
    [t.start() for t in th]; [t.join() for t in th]
    best = max(best, GB / (time.perf_counter() - t0) * 1e-9)
out["host_memcpy_all_threads_GBps"] = best
out["host_threads"] = nt
try:
    out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
except Exception as ex:
    out["topo"] = str(ex)
print(json.dumps(out))
