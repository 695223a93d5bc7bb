"""Host<->device copy bandwidth of the box: one direction at a time and both at once (two streams),
pinned host memory. Decides whether overlapping the upload and the download of the e2e path can pay."""
import sys
import time

import torch

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 30
h_up = torch.empty(n, dtype=torch.uint8).pin_memory()
h_dn = torch.empty(n, dtype=torch.uint8).pin_memory()
d_up = torch.empty(n, dtype=torch.uint8, device="cuda")
d_dn = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, dn, chunks=1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m = n // chunks
    for k in range(chunks):
        if up:
            with torch.cuda.stream(s1):
                d_up[k * m:(k + 1) * m].copy_(h_up[k * m:(k + 1) * m], non_blocking=True)
        if dn:
            with torch.cuda.stream(s2):
                h_dn[k * m:(k + 1) * m].copy_(d_dn[k * m:(k + 1) * m], non_blocking=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


for _ in range(2):
    run(True, True)
for name, up, dn, ch in (("h2d", True, False, 1), ("d2h", False, True, 1), ("both", True, True, 1), ("both/64 chunks", True, True, 64),
                         ("both/1024 chunks", True, True, 1024)):
    t = min(run(up, dn, ch) for _ in range(3))
    print(f"{name:18s} {n / 1e9:.2f} GB per direction in {t * 1e3:8.2f} ms = {n / t / 1e9:6.1f} GB/s per direction, {(int(up) + int(dn)) * n / t / 1e9:6.1f} GB/s total")
