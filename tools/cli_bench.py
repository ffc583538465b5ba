#!/usr/bin/env python
"""Files-in -> files-out throughput (the user-facing end to end: gzip inflate -> FASTQ index -> H2D -> kernels
-> D2H -> gzip deflate into 96 bin files) on synthetic COI reads, one JSON line per variant:

  * the fused `two-round` command on a .fastq.gz of this library's own writer (members with a size field:
    inflated member-parallel), on a FOREIGN single-stream .fastq.gz (zlib, what `gzip` leaves: inflated
    chunk-parallel by csrc/orc_pgz.h; ORC_NO_PGZ=1 in the environment gives the single zlib stream) and on plain FASTQ;
  * the reference script's own flow through the `cutadapt` shim (02_cutadapt_loop.sh:64-72 once, :94-102
    twelve times on the round-1 bins, which are member-structured because this library wrote them).

Run on the GPU box:

    python tools/cli_bench.py [--reads 1048576] [--dir /dev/shm/orc_cli]
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "nanopore-barcoding-orc_b200")
sys.path[:0] = [ROOT, PKG]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=1 << 20)
    ap.add_argument("--dir", default="/dev/shm/orc_cli")
    ap.add_argument("-j", type=int, default=os.cpu_count() or 8)
    a = ap.parse_args()
    import numpy as np
    from orcdemux import fastq as F
    from orcdemux import m13, synth
    shutil.rmtree(a.dir, ignore_errors=True)
    os.makedirs(os.path.join(a.dir, "pychopped"))
    fwd, rev, _ = m13.write_tables(os.path.join(a.dir, "adapters"))
    t0 = time.time()
    rs = synth.generate(a.reads, 300, 900, seed=1002, workers=min(16, a.j))
    raw = np.frombuffer(rs.to_fastq_bytes(), dtype=np.uint8)
    gen_s = time.time() - t0
    inputs = {"gz": os.path.join(a.dir, "pychopped", "pychopped_bench.fastq.gz"),
              "plain": os.path.join(a.dir, "pychopped", "pychopped_benchp.fastq"),
              "foreign": os.path.join(a.dir, "pychopped", "pychopped_benchf.fastq.gz")}

    class R:
        fastq = raw
        bin_offsets = np.array([0, raw.size, raw.size], dtype=np.uint64)
    t0 = time.time()
    w = F.BinWriters([inputs["gz"], inputs["plain"]], 5, a.j)
    R2 = type("R2", (), {"fastq": raw, "bin_offsets": np.array([0, 0, raw.size], dtype=np.uint64)})
    w.wait(w.write_batch(R))
    w.wait(w.write_batch(R2))
    w.close()
    import zlib
    co = zlib.compressobj(1, zlib.DEFLATED, 31)         # one gzip member, one deflate stream, like `gzip -1`
    with open(inputs["foreign"], "wb") as fh:
        mv = memoryview(raw)
        for o in range(0, raw.size, 1 << 24):
            fh.write(co.compress(mv[o:o + (1 << 24)]))
        fh.write(co.flush())
    write_s = time.time() - t0
    sys.stderr.write("generated %d reads in %.1f s, wrote both inputs in %.1f s (%.0f MB text, %.0f MB gz)\n" %
                     (a.reads, gen_s, write_s, raw.size / 1e6, os.path.getsize(inputs["gz"]) / 1e6))
    for variant, inp, extra, tag in (("two-round: fastq.gz (own members) -> 96 x fastq.gz (members coded on the GPU)", inputs["gz"], [], "gz"),
                                     ("two-round: fastq.gz (own members) -> 96 x fastq.gz (--host-gzip: zlib on the host threads)", inputs["gz"], ["--host-gzip"], "gzh"),
                                     ("two-round: fastq.gz (foreign, one stream) -> 96 x fastq.gz", inputs["foreign"], [], "foreign"),
                                     ("two-round: fastq -> 96 x fastq (no gzip)", inputs["plain"], ["--no-gzip"], "plain")):
        out = os.path.join(a.dir, "demuxed_" + tag)
        t0 = time.time()
        r = subprocess.run([sys.executable, "-m", "orcdemux.cli", "two-round", inp, "--sp5", fwd, "--sp27", rev,
                            "--outdir", out, "-j", str(a.j)] + extra, capture_output=True, text=True,
                           env=dict(os.environ, PYTHONPATH=PKG))
        wall = time.time() - t0
        if r.returncode != 0:
            sys.stderr.write(r.stderr)
            return 1
        ds = {"gz": "bench", "gzh": "bench", "foreign": "benchf", "plain": "benchp"}[tag]
        rep = json.load(open(os.path.join(out, "SP27", "orcdemux_%s.json" % ds)))
        binned = sum(rep["bins"].values())
        print(json.dumps({"variant": variant, "reads": rep["reads"], "reads_in_valid_bins": binned,
                          "pipeline_s": rep["elapsed_seconds"], "process_wall_s": wall,
                          "reads_per_s": rep["reads"] / rep["elapsed_seconds"],
                          "input_text_MB_per_s": raw.size / 1e6 / rep["elapsed_seconds"],
                          "host_threads": a.j,
                          "output_MB": sum(os.path.getsize(os.path.join(out, "SP27", f)) for f in os.listdir(os.path.join(out, "SP27"))) / 1e6}), flush=True)
    # the script's flow: round 1 once, round 2 on each of the twelve SP5 bins (13 processes, 13 engine set-ups)
    shim = os.path.join(PKG, "bin", "cutadapt")
    out = os.path.join(a.dir, "demuxed_script")
    os.makedirs(os.path.join(out, "SP5"))
    os.makedirs(os.path.join(out, "SP27"))
    ds = "bench"
    t0 = time.time()
    cmds = [[shim, "--action=trim", "-e", "0.1", "-j", str(a.j), "--rc", "-g", "file:" + fwd,
             "-o", os.path.join(out, "SP5", "{name}_%s.fastq.gz" % ds), inputs["gz"],
             "--json=" + os.path.join(out, "SP5", "cutadapt_SP5_%s.json" % ds)]]
    tenv = dict(os.environ, ORCDEMUX_TIMING="1")
    r = subprocess.run(cmds[0], capture_output=True, text=True, env=tenv)
    if r.returncode != 0:
        sys.stderr.write(r.stderr)
        return 1
    r1_s = time.time() - t0
    phases = [l for l in r.stderr.splitlines() if l.startswith("orcdemux timing")][:1]
    ids = sorted(f[:-len("_%s.fastq.gz" % ds)] for f in os.listdir(os.path.join(out, "SP5")) if f.endswith(".fastq.gz"))
    ids = [i for i in ids if "unknown" not in i]
    for ident in ids:
        r = subprocess.run([shim, "--action=trim", "-e", "0.1", "-j", str(a.j), "--rc", "-a", "file:" + rev,
                            "-o", os.path.join(out, "SP27", "{name}_%s_%s.fastq.gz" % (ident, ds)),
                            os.path.join(out, "SP5", "%s_%s.fastq.gz" % (ident, ds)),
                            "--json=" + os.path.join(out, "SP27", "%s_%s.json" % (ident, ds))],
                           capture_output=True, text=True, env=tenv)
        if r.returncode != 0:
            sys.stderr.write(r.stderr)
            return 1
        if len(phases) < 2:
            phases += [l for l in r.stderr.splitlines() if l.startswith("orcdemux timing")][:1]
    wall = time.time() - t0
    print(json.dumps({"variant": "script flow through the shim: 1 x round 1 + %d x round 2 (02:64-72, 94-102), gz in, gz out" % len(ids),
                      "reads": a.reads, "process_wall_s": wall, "round1_wall_s": r1_s, "reads_per_s": a.reads / wall,
                      "host_threads": a.j, "phases_round1_and_first_round2": phases}), flush=True)
    shutil.rmtree(a.dir, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
