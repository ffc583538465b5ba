"""Throughput of the long-adapter path (a round with an adapter over 64 nt runs long_kernel): synthetic COI reads
with a 5' adapter of 100 nt in front.  Usage: python tools/long_probe.py [reads]"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nanopore-barcoding-orc_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402

from orcdemux import engine as E, synth  # noqa: E402
from orcdemux.lib import ORC_BACK, ORC_FRONT  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
rnd = random.Random(1)
f = ["".join(rnd.choice("ACGT") for _ in range(100)) for _ in range(4)]
b = ["".join(rnd.choice("ACGT") for _ in range(100)) for _ in range(4)]
rs = synth.generate(n, 300, 900, seed=7, workers=8)
rounds = [E.Round(["f%d" % i for i in range(4)], f, ORC_FRONT, 0.1, 3, True, True),
          E.Round(["b%d" % i for i in range(4)], b, ORC_BACK, 0.1, 3, True, True)]
with E.Engine(rounds, max_reads=n, max_bytes=int(rs.seq.shape[0]), n_slots=1, want_matches=False) as eng:
    eng.run(rs)
    eng.launch(0)
    eng.sync(0)
    t = eng.timings(0)
print(json.dumps({"reads": n, "adapters": "4 x 100 nt per round, --rc", "step_ms": t["total_ms"],
                  "reads_per_s": n / (t["total_ms"] * 1e-3), "cells": t["cells"],
                  "gcups": sum(t["cells"]) / (t["total_ms"] * 1e-3) / 1e9}))
