"""How fast is the host at the reference's batch-1 layer arithmetic, per thread count?  (The
GPU box reports 128 cores; the CPU baseline must not oversubscribe them.)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

h, inter = 4096, 11008
w1 = torch.randn(3 * h, h)
w2 = torch.randn(2 * inter, h)
w3 = torch.randn(h, inter)
x = torch.randn(1, h)
print("affinity", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count(), flush=True)
for path in ("/sys/fs/cgroup/cpu.max",):
    try:
        print(path, open(path).read().strip(), flush=True)
    except Exception as e:
        print(path, "n/a", flush=True)
for nt in (8, 16, 32, 64, 128):
    if nt > (os.cpu_count() or 1):
        break
    torch.set_num_threads(nt)
    for _ in range(2):
        a = x @ w1.T; b = x @ w2.T; c = b[:, :inter] @ w3.T
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        a = x @ w1.T; b = x @ w2.T; c = b[:, :inter] @ w3.T
    dt = (time.perf_counter() - t0) / n
    gb = (w1.numel() + w2.numel() + w3.numel()) * 4 / 1e9
    print(f"threads={nt:4d}  layer-sized GEMV set {dt * 1e3:8.2f} ms  -> {gb / dt:7.1f} GB/s", flush=True)
    if dt > 2.0:
        break
