#!/usr/bin/env python
"""How fast can a kernel read pinned host memory over PCIe (orc_probe_hostread), next to a cudaMemcpy of the
same buffer?  Decides whether emit_kernel should read the qualities in place (orc_params.qual_zero_copy)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nanopore-barcoding-orc_b200")]
import torch
from orcdemux import lib

L = lib.load()
n = 1 << 30
a = torch.empty(n + 4096, dtype=torch.uint8, pin_memory=True)
a.random_(0, 255)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(a[:n], non_blocking=True); torch.cuda.synchronize()
    h2d = n / (time.perf_counter() - t0) / 1e9
print("cudaMemcpy H2D %.1f GB/s" % h2d)
for chunk, stride in ((512, 512), (4096, 4096), (480, 608), (480, 1216), (256, 1216), (1024, 1216)):
    g = L.orc_probe_hostread(0, a.data_ptr(), n, chunk, stride)
    print("kernel reads %4d of every %4d bytes: %.1f GB/s useful" % (chunk, stride, g))
# the same while a cudaMemcpy H2D runs beside it on another stream
b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
s2 = torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.stream(s2):
    for _ in range(4):
        d.copy_(b, non_blocking=True)
g = L.orc_probe_hostread(0, a.data_ptr(), n, 480, 608)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("beside 4 GiB of cudaMemcpy H2D: kernel reads 480/608 at %.1f GB/s useful; both done in %.1f ms = %.1f GB/s total useful"
      % (g, dt * 1e3, (4 * n + 3 * (n // 608) * 480) / dt / 1e9))
# SM-driven host reads while the copy engine moves data the OTHER way (D2H) -- the traffic pattern of an e2e
# leg in which pack_kernel and emit_kernel read their inputs in place and only the FASTQ text is copied back
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.stream(s2):
    for _ in range(4):
        b.copy_(d, non_blocking=True)
g = L.orc_probe_hostread(0, a.data_ptr(), n, 4096, 4096)
t1 = time.perf_counter() - t0
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("beside 4 GiB of cudaMemcpy D2H: kernel reads contiguously at %.1f GB/s (probe returned after %.1f ms); the D2H copies took %.1f ms = %.1f GB/s"
      % (g, t1 * 1e3, dt * 1e3, 4 * n / dt / 1e9))
