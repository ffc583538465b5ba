import sys, time; sys.path.insert(0,'nanopore-barcoding-orc_b200')
import numpy as np, torch
from orcdemux import engine as E, synth
rs = E.pin_readset(synth.generate(1<<20, 300, 900, 1002, workers=16))
eng = E.Engine(E.m13_rounds(), max_reads=rs.n_reads, max_bytes=int(rs.seq.shape[0]), max_name_bytes=int(rs.names.shape[0])+64, n_slots=3)
for i in range(3):
    eng.submit(0, rs); r = eng.wait(0, copy=False); t = eng.timings(0)
    print("serial: h2d %.2f ms (%.1f GB/s)  kernels %.2f  d2h %.2f ms" % (t["h2d_ms"], 1.268/t["h2d_ms"]*1e3, t["total_ms"], t["d2h_ms"]), "fastq GB", r.fastq.nbytes/1e9)
# raw torch copy bandwidth
a = torch.empty(1<<30, dtype=torch.uint8, pin_memory=True); d = torch.empty(1<<30, dtype=torch.uint8, device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter(); d.copy_(a, non_blocking=True); torch.cuda.synchronize(); h2d=time.perf_counter()-t0
    t0=time.perf_counter(); a.copy_(d, non_blocking=True); torch.cuda.synchronize(); d2h=time.perf_counter()-t0
print("torch 1GiB h2d %.1f GB/s d2h %.1f GB/s" % (1.0737/h2d, 1.0737/d2h))
s1=torch.cuda.Stream(); s2=torch.cuda.Stream(); b = torch.empty(1<<30, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(1<<30, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0=time.perf_counter()
with torch.cuda.stream(s1): d.copy_(a, non_blocking=True)
with torch.cuda.stream(s2): b.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); print("bidirectional 1GiB each: %.1f ms" % ((time.perf_counter()-t0)*1e3))
# pipelined e2e with timing per iteration
K=8; S=3; inflight=[]; t0=time.perf_counter(); marks=[]
for i in range(K):
    s=i%S
    if len(inflight)==S:
        eng.wait(inflight.pop(0), copy=False); marks.append(time.perf_counter()-t0)
    eng.submit(s, rs); inflight.append(s)
while inflight: eng.wait(inflight.pop(0), copy=False); marks.append(time.perf_counter()-t0)
print("pipelined: per-step completion times ms", [round(1e3*(b-a),1) for a,b in zip([0]+marks[:-1], marks)])
