#!/usr/bin/env python
"""Files in -> files out from a FOREIGN single-stream .fastq.gz (what `gzip` leaves of pychopper's output, the
input of 02_cutadapt_loop.sh:64-72): the fused two-round command with the chunk-parallel inflate of
csrc/orc_pgz.h and, for comparison, with ORC_NO_PGZ=1 (one zlib stream).  The two output trees must hold the
same bytes after decompression.  One JSON line per variant.

    python tools/foreign_gz_bench.py [--reads 524288] [--dir /dev/shm/orc_fgz]
"""
import argparse
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "nanopore-barcoding-orc_b200")
sys.path[:0] = [ROOT, PKG]


def tree_digest(d):
    h = hashlib.sha256()
    for f in sorted(os.listdir(d)):
        if f.endswith(".fastq.gz"):
            h.update(f.encode())
            h.update(hashlib.sha256(gzip.open(os.path.join(d, f)).read()).digest())
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=1 << 19)
    ap.add_argument("--dir", default="/dev/shm/orc_fgz")
    ap.add_argument("-j", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--only-pgz", action="store_true", help="skip the zlib variant and the comparison of the trees")
    a = ap.parse_args()
    import numpy as np
    from orcdemux import m13, synth
    shutil.rmtree(a.dir, ignore_errors=True)
    os.makedirs(a.dir)
    fwd, rev, _ = m13.write_tables(os.path.join(a.dir, "adapters"))
    rs = synth.generate(a.reads, 300, 900, seed=1002, workers=min(16, a.j))
    raw = np.frombuffer(rs.to_fastq_bytes(), dtype=np.uint8)
    inp = os.path.join(a.dir, "pychopped_f.fastq.gz")
    co = zlib.compressobj(1, zlib.DEFLATED, 31)         # one gzip member, one deflate stream, like `gzip -1`
    with open(inp, "wb") as fh:
        mv = memoryview(raw)
        for o in range(0, raw.size, 1 << 24):
            fh.write(co.compress(mv[o:o + (1 << 24)]))
        fh.write(co.flush())
    digests = []
    variants = (("chunk-parallel inflate (csrc/orc_pgz.h)", {}), ("one zlib stream (ORC_NO_PGZ=1)", {"ORC_NO_PGZ": "1"}))
    for tag, env in variants[:1] if a.only_pgz else variants:
        out = os.path.join(a.dir, "demuxed_" + ("pgz" if not env else "zlib"))
        t0 = time.time()
        r = subprocess.run([sys.executable, "-m", "orcdemux.cli", "two-round", inp, "--sp5", fwd, "--sp27", rev,
                            "--outdir", out, "-j", str(a.j)], capture_output=True, text=True,
                           env=dict(os.environ, PYTHONPATH=PKG, ORC_IO_DEBUG="1", **env))
        wall = time.time() - t0
        if r.returncode != 0:
            sys.stderr.write(r.stderr)
            return 1
        rep = json.load(open(os.path.join(out, "SP27", "orcdemux_f.json")))
        digests.append("" if a.only_pgz else tree_digest(os.path.join(out, "SP27")))
        print(json.dumps({"variant": "two-round: foreign fastq.gz -> 96 x fastq.gz, " + tag, "reads": rep["reads"],
                          "pipeline_s": rep["elapsed_seconds"], "process_wall_s": wall,
                          "reads_per_s": rep["reads"] / rep["elapsed_seconds"],
                          "input_text_MB_per_s": raw.size / 1e6 / rep["elapsed_seconds"], "host_threads": a.j,
                          "gz_MB": os.path.getsize(inp) / 1e6, "tree_sha256": digests[-1][:16],
                          "reader_thread": [l for l in r.stderr.splitlines() if l.startswith("orc_reader:")][:1]}), flush=True)
    shutil.rmtree(a.dir, ignore_errors=True)
    if not a.only_pgz and digests[0] != digests[1]:
        sys.stderr.write("the two trees differ\n")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
