"""Device time of the gzip stage (orc_params.emit_gzip) beside the step it follows: one batch, drop-bins of the
reference script's final tree or every bin, printed as JSON.  Usage: python tools/gz_probe.py [reads] [all]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nanopore-barcoding-orc_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402

from orcdemux import engine as E, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
keep_all = len(sys.argv) > 2 and sys.argv[2] == "all"
rs = synth.generate(n, 300, 900, seed=1002, workers=8)
drop = np.zeros(169, dtype=np.uint8)
if not keep_all:
    drop[0::13] = 1
    drop[:13] = 1
    drop[9 * 13:] = 1            # SP27_009..012 (02:114-118)
eng = E.Engine(E.m13_rounds(), max_reads=n, max_bytes=int(rs.seq.shape[0]), n_slots=1, want_matches=False,
               drop_bins=drop, emit_gzip=True)
res = eng.run(rs)
for _ in range(3):
    eng.launch(0)
    eng.sync(0)
t = eng.timings(0)
text = t["emit_bytes"] // 2
print(json.dumps({"reads": n, "bins_kept": int((drop == 0).sum()), "text_bytes": text, "gzip_bytes": t["gzip_bytes"],
                  "ratio": t["gzip_bytes"] / max(text, 1), "gzip_ms": t["gzip_ms"], "emit_ms": t["emit_ms"],
                  "step_ms": t["total_ms"], "gzip_read_gbs": text / (t["gzip_ms"] * 1e-3) / 1e9}))
eng.close()
