#!/usr/bin/env python
"""Files-in -> files-out throughput of the fused two-round command (the user-facing end to end:
gzip inflate -> FASTQ index -> H2D -> kernels -> D2H -> gzip deflate into 96 bin files), on
synthetic COI reads.  Prints one JSON line per variant.  Run on the GPU box:

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
              "plain": os.path.join(a.dir, "pychopped", "pychopped_benchp.fastq")}

    class R:
        fastq = raw
        bin_offsets = np.array([0, raw.size, raw.size], dtype=np.uint64)
    t0 = time.time()
    w = F.BinWriters([inputs["gz"], inputs["plain"]], 5, a.j)
    R2 = type("R2", (), {"fastq": raw, "bin_offsets": np.array([0, 0, raw.size], dtype=np.uint64)})
    w.wait(w.write_batch(R))
    w.wait(w.write_batch(R2))
    w.close()
    write_s = time.time() - t0
    sys.stderr.write("generated %d reads in %.1f s, wrote both inputs in %.1f s (%.0f MB text, %.0f MB gz)\n" %
                     (a.reads, gen_s, write_s, raw.size / 1e6, os.path.getsize(inputs["gz"]) / 1e6))
    for variant, inp, extra in (("fastq.gz -> 96 x fastq.gz", inputs["gz"], []),
                                ("fastq -> 96 x fastq (no gzip)", inputs["plain"], ["--no-gzip"])):
        out = os.path.join(a.dir, "demuxed_" + ("gz" if not extra else "plain"))
        t0 = time.time()
        r = subprocess.run([sys.executable, "-m", "orcdemux.cli", "two-round", inp, "--sp5", fwd, "--sp27", rev,
                            "--outdir", out, "-j", str(a.j)] + extra, capture_output=True, text=True,
                           env=dict(os.environ, PYTHONPATH=PKG))
        wall = time.time() - t0
        if r.returncode != 0:
            sys.stderr.write(r.stderr)
            return 1
        ds = "bench" if not extra else "benchp"
        rep = json.load(open(os.path.join(out, "SP27", "orcdemux_%s.json" % ds)))
        binned = sum(rep["bins"].values())
        print(json.dumps({"variant": variant, "reads": rep["reads"], "reads_in_valid_bins": binned,
                          "pipeline_s": rep["elapsed_seconds"], "process_wall_s": wall,
                          "reads_per_s": rep["reads"] / rep["elapsed_seconds"],
                          "input_text_MB_per_s": raw.size / 1e6 / rep["elapsed_seconds"],
                          "host_threads": a.j}), flush=True)
    shutil.rmtree(a.dir, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
