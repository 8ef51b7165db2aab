#!/usr/bin/env python
"""Host-fabric probe: aggregate pinned host->device copy bandwidth with 1, 2, 4 .. N GPUs copying concurrently (one thread and
one pinned 1 GiB buffer per GPU, plain cudaMemcpyAsync through torch).  Explains the end-to-end (host-buffer) scaling of
bench.py: the e2e number is H2D-bound, and what G ranks share is this box's host memory / PCIe fabric, not the GPUs."""
import json
import sys
import threading
import time

import torch

N = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
N = min(N, torch.cuda.device_count())
GIB = 1 << 30
host = [torch.empty(GIB, dtype=torch.uint8).pin_memory() for _ in range(N)]
dev = [torch.empty(GIB, dtype=torch.uint8, device=f"cuda:{i}") for i in range(N)]
out = {"bytes_per_copy": GIB, "results": []}
g = 1
while g <= N:
    def worker(i, reps, res):
        torch.cuda.set_device(i)
        s = torch.cuda.Stream(device=i)
        with torch.cuda.stream(s):
            dev[i].copy_(host[i], non_blocking=True)
            s.synchronize()
            bar.wait()
            t0 = time.perf_counter()
            for _ in range(reps):
                dev[i].copy_(host[i], non_blocking=True)
            s.synchronize()
            res[i] = time.perf_counter() - t0
    bar = threading.Barrier(g)
    res = [0.0] * g
    th = [threading.Thread(target=worker, args=(i, 4, res)) for i in range(g)]
    [t.start() for t in th]
    [t.join() for t in th]
    out["results"].append({"gpus_copying": g, "aggregate_gbs": g * 4 * GIB / max(res) / 1e9, "per_gpu_gbs": [4 * GIB / r / 1e9 for r in res]})
    g *= 2
print(json.dumps(out))
