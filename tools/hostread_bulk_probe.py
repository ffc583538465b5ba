#!/usr/bin/env python
"""SM loads against bulk asynchronous copies (cp.async.bulk into shared memory and on to device memory) for reading
pinned host memory over PCIe: alone, beside a cudaMemcpy H2D, beside a cudaMemcpy D2H, beside both -- the traffic
of the e2e leg, where emit_kernel reads the qualities in place (tools/e2e_diag.py)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nanopore-barcoding-orc_b200")]
import torch
from orcdemux import lib

L = lib.load()
n = 1 << 30
a = torch.empty(n + 65536, dtype=torch.uint8, pin_memory=True)
a.random_(0, 255)
b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
c = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()


def probe(mode, chunk, stride, h2d=False, d2h=False, reps=3):
    if mode:
        os.environ["ORC_PROBE_BULK"] = str(mode)
    else:
        os.environ.pop("ORC_PROBE_BULK", None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if h2d:
        with torch.cuda.stream(s_h2d):
            for _ in range(reps):
                d.copy_(b, non_blocking=True)
    if d2h:
        with torch.cuda.stream(s_d2h):
            for _ in range(reps):
                c.copy_(d2, non_blocking=True)
    g = L.orc_probe_hostread(0, a.data_ptr(), n, chunk, stride)
    t1 = time.perf_counter() - t0
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return g, t1, dt


print("mode 0 = 16-byte loads per lane (8 blocks of 256 threads per SM); mode k = bulk copies, k blocks of 8 issuing lanes per SM")
for chunk, stride in ((480, 608), (2048, 2048), (4096, 4096), (8192, 8192)):
    for mode in (0, 1, 2, 3):
        if mode and 16 * chunk * mode > 200 * 1024:
            continue
        g = probe(mode, chunk, stride)[0]
        line = "chunk %5d of %5d  mode %d: alone %5.1f GB/s" % (chunk, stride, mode, g)
        for h2d, d2h, what in ((True, False, "beside H2D"), (False, True, "beside D2H"), (True, True, "beside both")):
            g, t1, dt = probe(mode, chunk, stride, h2d, d2h)
            copied = (3 * n if h2d else 0), (3 * n if d2h else 0)
            line += " | %s: probe %5.1f GB/s, copies done after %6.1f ms (%5.1f GB/s per copy direction)" % (
                what, g, dt * 1e3, max(copied) / dt / 1e9)
        print(line, flush=True)
