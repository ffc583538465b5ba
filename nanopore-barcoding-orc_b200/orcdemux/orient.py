"""`orcdemux orient` -- the step in front of the demultiplexer: find the two amplicon primers in every read,
turn the read into the orientation the primer configuration calls "+", keep the primers, sort out what has no
valid pair.  This is what the reference asks of pychopper
(/root/reference/scripts/01_pychopper.sh:45-57:
     pychopper -b M13_seqs_for_pychopper.fa -c M13_config_for_pychopper.txt -k LSK114 -Q 10
               -w X_rescued.fastq -u X_unclass.fastq -l X_short.fastq -S X_stats.out -p -t 24 -m edlib IN > X_pass.fastq)
with the primers of adapters_primers/M13_seqs_for_pychopper.fa (the 17-nt index is written as N x 17) and the
configuration "+:SP5,-SP27|-:SP27,-SP5".

It is pychopper-STYLE, not a restatement of pychopper: the search is the demultiplexer's own (cutadapt
semantics: IUPAC wildcards in the primers, error rate relative to the non-N primer length, best match, one
segment per read) on the same kernels, as a two-round job with --rc and --action=retain:

    round 1   -g <5' primer of the "+" configuration>   --rc   (a "-" read is found on its reverse complement)
    round 2   -a <3' primer of the "+" configuration>   --rc

A read passes if both primers are found and round 2 did not turn the read again; it comes out in "+"
orientation from the first base of the 5' primer to the last base of the 3' primer (-p; without -p the
primers are cut off).  pychopper's hit chaining over several primer occurrences, its rescue of fused reads
(-w: the file is created empty) and its cutoff autotuning are not rebuilt.  Reads whose mean quality (mean
error probability, as a Phred value) is below -Q go nowhere, as in pychopper; reads shorter than -z after
orientation go to -l.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from typing import List, Optional, Tuple

import numpy as np

from . import engine as E
from . import fastq as F
from .lib import ORC_BACK, ORC_FRONT
from .primers import Unsupported

_COMP = str.maketrans("ACGTUMRWSYKVHDBNacgtumrwsykvhdbn", "TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn")


def revcomp(s: str) -> str:
    return s.translate(_COMP)[::-1]


def parse_config(text: str) -> List[Tuple[str, Tuple[str, bool], Tuple[str, bool]]]:
    """"+:SP5,-SP27|-:SP27,-SP5" -> [("+", ("SP5", False), ("SP27", True)), ("-", ("SP27", False), ("SP5", True))]
    (True: the primer's reverse complement)."""
    out = []
    for part in text.strip().split("|"):
        if not part.strip():
            continue
        strand, _, pair = part.partition(":")
        names = [x.strip() for x in pair.split(",")]
        if strand.strip() not in ("+", "-") or len(names) != 2 or not all(names):
            raise Unsupported("primer configuration %r (expected +:A,-B|-:B,-A)" % part)
        out.append((strand.strip(), *[(n.lstrip("-"), n.startswith("-")) for n in names]))
    return out


def primer_pair(primers: dict, config) -> Tuple[str, str]:
    """The (5' primer, 3' primer) of the "+" configuration as they appear on a "+" read.  The "-"
    configuration must be its mirror image (then a "-" read is a "+" read reverse-complemented, which is what
    --rc searches)."""
    plus = [c for c in config if c[0] == "+"]
    minus = [c for c in config if c[0] == "-"]
    if len(plus) != 1 or len(minus) > 1:
        raise Unsupported("exactly one + configuration (and at most one -) is handled")
    (_, (a, arc), (b, brc)) = plus[0]
    for n in (a, b):
        if n not in primers:
            raise Unsupported("primer %s of the configuration is not in the primer file" % n)
    five = revcomp(primers[a]) if arc else primers[a]
    three = revcomp(primers[b]) if brc else primers[b]
    if minus:
        (_, (c, crc), (d, drc)) = minus[0]
        m5 = revcomp(primers[c]) if crc else primers[c]
        m3 = revcomp(primers[d]) if drc else primers[d]
        if m5.upper() != revcomp(three).upper() or m3.upper() != revcomp(five).upper():
            raise Unsupported("the - configuration is not the reverse complement of the + configuration")
    return five, three


_QERR = 10.0 ** (-(np.arange(256, dtype=np.float64) - 33.0) / 10.0)


def mean_qscores(text: np.ndarray, qual_offsets: np.ndarray, lengths: np.ndarray, block: int = 4096) -> np.ndarray:
    """Phred value of the mean error probability of every read (0 for empty reads); reads in file order."""
    n = lengths.shape[0]
    out = np.zeros(n, dtype=np.float64)
    L = lengths.astype(np.int64)
    q0 = qual_offsets.astype(np.int64)
    for a in range(0, n, block):                    # a block of reads at a time keeps the temporaries small
        b = min(n, a + block)
        lo, hi = int(q0[a]), int((q0[a:b] + L[a:b]).max())
        csum = np.concatenate([[0.0], np.cumsum(_QERR[text[lo:hi]])])
        s = q0[a:b] - lo
        tot = csum[s + L[a:b]] - csum[s]
        nz = L[a:b] > 0
        o = np.zeros(b - a)
        o[nz] = -10.0 * np.log10(np.maximum(tot[nz] / L[a:b][nz], 1e-30))
        out[a:b] = o
    return out


class _Sub:
    """A TextBatch restricted to some of its reads (the raw-text layout takes any subset)."""

    def __init__(self, tb, keep: np.ndarray):
        self.text, self.n_bytes = tb.text, tb.n_bytes
        self.n_reads = int(keep.shape[0])
        mk = lambda a: np.ascontiguousarray(a[:tb.n_reads][keep])
        self.offsets, self.lengths = mk(tb.offsets), mk(tb.lengths)
        self.qual_offsets, self.name_offsets, self.name_lengths = mk(tb.qual_offsets), mk(tb.name_offsets), mk(tb.name_lengths)


def run(input_path: str, out_path: str, primers: dict, config, min_qual: float = 7.0, keep_primers: bool = False,
        unclassified: Optional[str] = None, short: Optional[str] = None, rescued: Optional[str] = None,
        stats: Optional[str] = None, error_rate: float = 0.15, min_overlap: int = 10, min_len: int = 50,
        threads: int = 8, device: int = 0, level: int = 1) -> dict:
    five, three = primer_pair(primers, config)
    action = "retain" if keep_primers else "trim"
    rounds = [E.Round(["five_prime"], [five], ORC_FRONT, error_rate, min_overlap, True, True, action),
              E.Round(["three_prime"], [three], ORC_BACK, error_rate, min_overlap, True, True, action)]
    # bins: (a1 + 1) + 2 * (a2 + 1): 0 none, 1 only the 5' primer, 2 (cannot happen), 3 both
    t0 = time.time()
    from .cli import _batch_shape
    (max_reads, max_bytes), slots = _batch_shape(), 3
    reader = F.FastqReader(input_path, max_reads, max_bytes, keep=3, ahead=2, threads=max(1, threads // 2))
    c = dict(total=0, lowq=0, passed=0, plus=0, minus=0, unclassified=0, short=0, turned_twice=0)
    if rescued:
        open(rescued, "wb").close()                 # fused reads are not split: nothing is ever rescued
    out_fh = sys.stdout.buffer if out_path == "-" else open(out_path, "wb")
    unc_fh = open(unclassified, "wb") if unclassified else None
    short_fh = open(short, "wb") if short else None

    def handle(res, tb):
        m0, m1 = res.matches
        both = (m0["adapter"] >= 0) & (m1["adapter"] >= 0)
        ok = both & (m1["is_rc"] == 0)                  # round 2 turning the read again: no valid pair
        c["turned_twice"] += int((both & ~ok).sum())
        long_enough = res.out_len >= min_len
        # a bin holds its reads in input order, so its records line up with the reads that went there
        for b in (0, 1, 2):                             # no primer / only the 5' primer: unclassified
            text = res.fastq[int(res.bin_offsets[b]):int(res.bin_offsets[b + 1])]
            c["unclassified"] += int(res.bin_counts[b])
            if unc_fh and text.size:
                unc_fh.write(text.tobytes())
        text = res.fastq[int(res.bin_offsets[3]):int(res.bin_offsets[4])]
        idx = np.flatnonzero(res.bin == 3)
        if idx.size == 0:
            return
        good = ok[idx] & long_enough[idx]
        c["passed"] += int(good.sum())
        c["minus"] += int((good & (m0["is_rc"][idx] != 0)).sum())
        c["plus"] += int((good & (m0["is_rc"][idx] == 0)).sum())
        c["short"] += int((ok[idx] & ~long_enough[idx]).sum())
        c["unclassified"] += int((~ok[idx]).sum())
        if good.all():
            out_fh.write(text.tobytes())
            return
        nl = np.flatnonzero(text == 10)                 # four lines per record
        ends = nl[3::4] + 1
        starts = np.concatenate([[0], ends[:-1]])
        raw = text.tobytes()
        for j in range(idx.size):
            piece = raw[int(starts[j]):int(ends[j])]
            if good[j]:
                out_fh.write(piece)
            elif ok[idx[j]]:
                if short_fh:
                    short_fh.write(piece)
            elif unc_fh:
                unc_fh.write(piece)

    try:
        with E.Engine(rounds, device=device, max_reads=max_reads, max_bytes=max_bytes, n_slots=slots,
                      emit_fastq=True, want_matches=True) as eng:
            pending = []
            k = 0
            for tb in reader:
                c["total"] += tb.n_reads
                q = mean_qscores(tb.text, tb.qual_offsets[:tb.n_reads], tb.lengths[:tb.n_reads])
                keep = np.flatnonzero(q >= min_qual)
                c["lowq"] += tb.n_reads - int(keep.size)
                if keep.size == 0:
                    continue
                sub = _Sub(tb, keep)
                slot = k % slots
                if len(pending) == slots - 1:
                    s_, tb_ = pending.pop(0)
                    handle(eng.wait(s_), tb_)
                eng.submit(slot, sub)
                pending.append((slot, sub))
                k += 1
            while pending:
                s_, tb_ = pending.pop(0)
                handle(eng.wait(s_), tb_)
    finally:
        reader.close()
        for fh in (unc_fh, short_fh):
            if fh:
                fh.close()
        if out_fh is not sys.stdout.buffer:
            out_fh.close()
    c["elapsed_seconds"] = time.time() - t0
    if stats:
        with open(stats, "w") as fh:
            fh.write("Category\tName\tValue\n")
            for kx in ("total", "lowq", "passed", "plus", "minus", "unclassified", "short", "turned_twice"):
                fh.write("Classification\t%s\t%d\n" % (kx, c[kx]))
    return c


def main(argv: List[str]) -> int:
    ap = argparse.ArgumentParser(prog="orcdemux orient", description=__doc__.split("\n\n")[0])
    ap.add_argument("-b", required=True, help="primer FASTA (pychopper -b)")
    ap.add_argument("-c", required=True, help="primer configuration file or string (pychopper -c)")
    ap.add_argument("-k", default=None, help="sequencing kit: accepted and ignored (the primers come from -b)")
    ap.add_argument("-Q", type=float, default=7.0, help="minimum mean base quality")
    ap.add_argument("-z", type=int, default=50, help="minimum length of an oriented read")
    ap.add_argument("-w", default=None, help="rescued reads (created empty: fused reads are not split)")
    ap.add_argument("-u", default=None, help="unclassified reads")
    ap.add_argument("-l", default=None, help="reads shorter than -z after orientation")
    ap.add_argument("-S", default=None, help="statistics, tab separated")
    ap.add_argument("-p", action="store_true", help="keep the primers")
    ap.add_argument("-t", type=int, default=8, help="host threads")
    ap.add_argument("-m", default="edlib", help="pychopper's backend choice: accepted and ignored")
    ap.add_argument("-e", type=float, default=0.15, help="error rate of the primer search (relative to the non-N length)")
    ap.add_argument("-O", type=int, default=10, help="minimum overlap of a primer with the read")
    ap.add_argument("input")
    ap.add_argument("output", nargs="?", default="-")
    a = ap.parse_args(argv)
    names, seqs = F.read_adapters_fasta(a.b)
    text = open(a.c).read() if os.path.exists(a.c) else a.c
    c = run(a.input, a.output, dict(zip(names, seqs)), parse_config(text), a.Q, a.p, a.u, a.l, a.w, a.S, a.e, a.O, a.z, a.t)
    sys.stderr.write("orcdemux orient: %d reads, %d passed (%d + / %d -), %d unclassified, %d below Q %.1f, %d short\n"
                     % (c["total"], c["passed"], c["plus"], c["minus"], c["unclassified"], c["lowq"], a.Q, c["short"]))
    return 0
