"""cutadapt-compatible command line for the demultiplexing step (drop-in for the two call
shapes of /root/reference/scripts/02_cutadapt_loop.sh):

  round 1 (02:64-72)   cutadapt --action=trim -e 0.1 -j 24 --rc -g file:FWD.fa
                                -o DIR/SP5/{name}_DS.fastq.gz IN.fastq.gz --json=REPORT
  round 2 (02:94-102)  cutadapt --action=trim -e 0.1 -j 24 --rc -a file:REV_RC.fa
                                -o DIR/SP27/{name}_SP5id_DS.fastq.gz DIR/SP5/SP5id_DS.fastq.gz --json=REPORT

Kept surface: -g/-a (file:PATH, SEQ, name=SEQ; several allowed), -e, -O, --action=trim, --rc,
-j (accepted), --json, --no-indels (regular adapters: Hamming distance along diagonals; anchored
^file: / file$: adapters: the indexed Hamming fast path; anchored adapters WITH indels are refused),
-o with {name} (one file per adapter name plus "unknown", all created even if empty).
Adapters may hold IUPAC wildcards.  Without {name} in -o: one output file, linked -g FWD...REV pairs or
one plain adapter type, --untrimmed-output / --discard-untrimmed, FASTA or FASTQ (primers.py; the
call shapes of 04_cleaning_primers.sh:371-388, 468-507).
Anything else exits with status 2 and an "unsupported" message: there is no CPU fallback.

`python -m orcdemux.cli two-round ...` runs both rounds fused on the GPU and leaves the file
tree of the whole script (02:107-119: unknown and SP27_009..012 removed).
"""
from __future__ import annotations

import json
import os
import sys
import time
from typing import List, Optional

import numpy as np

from . import engine as E
from . import fastq as F
from .lib import ORC_BACK, ORC_FRONT, ORC_PREFIX, ORC_SUFFIX


from .primers import Unsupported        # one exception type for everything outside the surface (exit status 2)


def _parse_adapter_specs(specs: List[str], kind: int):
    """-> (names, sequences, anchored).  Syntax kept: SEQ, name=SEQ, file:PATH, and the anchored
    forms ^SEQ / ^file:PATH (5') and SEQ$ / file$:PATH (3')."""
    names, seqs, flags = [], [], []
    for spec in specs:
        anchored = False
        s = spec
        if kind == ORC_FRONT and s.startswith("^"):
            anchored, s = True, s[1:]
        if s.startswith("file$:") and kind == ORC_BACK:
            anchored, s = True, "file:" + s[6:]
        if s.startswith("file:"):
            n, q = F.read_adapters_fasta(s[5:])
        else:
            name = None
            if "=" in s:
                name, s = s.split("=", 1)
            if kind == ORC_BACK and s.endswith("$"):
                anchored, s = True, s[:-1]
            if kind == ORC_FRONT and s.startswith("^"):
                anchored, s = True, s[1:]
            n, q = [name], [s.upper().replace("U", "T")]
        for a, b in zip(n, q):
            if any(c in b for c in "X.;{}[]$^"):
                raise Unsupported("adapter syntax beyond plain sequences: %r" % b)
            names.append(a)
            seqs.append(b)
            flags.append(anchored)
    if any(flags) and not all(flags):
        raise Unsupported("a mix of anchored and regular adapters in one invocation")
    # cutadapt names unnamed adapters "1", "2", ... in order
    names = [nm if nm else str(i + 1) for i, nm in enumerate(names)]
    return names, seqs, bool(flags and flags[0])


def parse_cutadapt_argv(argv: List[str]):
    opt = dict(e=0.1, O=3, rc=False, indels=True, action="trim", json=None, out=None, cores=1,
               g=[], a=[], level=1, inputs=[], quiet=False,      # cutadapt 4.x: gzip level 1 unless --compression-level
               untrimmed_output=None, discard_untrimmed=False, order=[])      # order: "g"/"a" as given
    i = 0

    def need(flag):
        nonlocal i
        i += 1
        if i >= len(argv):
            raise Unsupported("option %s needs a value" % flag)
        return argv[i]

    while i < len(argv):
        a = argv[i]
        if a.startswith("--") and "=" in a:
            key, val = a.split("=", 1)
        else:
            key, val = a, None
        if key in ("-e", "--error-rate", "--errors"):
            opt["e"] = float(val if val is not None else need(key))
        elif a.startswith("-e") and len(a) > 2 and not a.startswith("--"):
            opt["e"] = float(a[2:])
        elif key in ("-O", "--overlap"):
            opt["O"] = int(val if val is not None else need(key))
        elif a.startswith("-O") and len(a) > 2 and not a.startswith("--"):
            opt["O"] = int(a[2:])
        elif key in ("-j", "--cores"):
            opt["cores"] = int(val if val is not None else need(key))
        elif a.startswith("-j") and len(a) > 2 and not a.startswith("--"):
            opt["cores"] = int(a[2:])
        elif key in ("--rc", "--revcomp"):
            opt["rc"] = True
        elif key == "--no-indels":
            opt["indels"] = False
        elif key == "--action":
            opt["action"] = val if val is not None else need(key)
        elif key in ("-g", "--front"):
            opt["g"].append(val if val is not None else need(key))
            opt["order"].append("g")
        elif key in ("-a", "--adapter"):
            opt["a"].append(val if val is not None else need(key))
            opt["order"].append("a")
        elif key in ("-o", "--output"):
            opt["out"] = val if val is not None else need(key)
        elif key == "--untrimmed-output":
            opt["untrimmed_output"] = val if val is not None else need(key)
        elif key in ("--discard-untrimmed", "--trimmed-only"):
            opt["discard_untrimmed"] = True
        elif key == "--json":
            opt["json"] = val if val is not None else need(key)
        elif key in ("-Z",):
            opt["level"] = 1
        elif key == "--compression-level":
            opt["level"] = int(val if val is not None else need(key))
        elif key in ("--quiet", "--report"):
            if key == "--report" and val is None:
                need(key)
            opt["quiet"] = opt["quiet"] or key == "--quiet"
        elif a.startswith("-") and a != "-":
            raise Unsupported("cutadapt option %s is outside the demultiplexing surface this build replaces" % a)
        else:
            opt["inputs"].append(a)
        i += 1
    if opt["action"] not in ("trim", "retain"):
        raise Unsupported("--action=%s (trim and retain are built)" % opt["action"])
    if len(opt["inputs"]) != 1:
        raise Unsupported("exactly one (single-end) input file is expected, got %d" % len(opt["inputs"]))
    if not opt["out"]:
        raise Unsupported("-o is required")
    if "{name}" not in opt["out"]:
        # one output file: the primer-trimming call shapes of 04_cleaning_primers.sh (orcdemux/primers.py)
        if not (opt["g"] or opt["a"]):
            raise Unsupported("no adapters given")
        if opt["action"] != "trim":
            raise Unsupported("--action=%s with a single output file" % opt["action"])
        return opt
    if opt["untrimmed_output"] or opt["discard_untrimmed"]:
        raise Unsupported("--untrimmed-output / --discard-untrimmed together with a {name} template")
    if any("..." in x for x in opt["g"] + opt["a"]):
        raise Unsupported("linked adapters together with a {name} template")
    if bool(opt["g"]) == bool(opt["a"]):
        raise Unsupported("give either -g or -a adapters (one adapter type per invocation)")
    return opt


def run_primer_trim(opt, argv, device=0) -> int:
    """One output file (no {name}): linked -g FWD...REV pairs or one plain adapter type, optional
    --untrimmed-output / --discard-untrimmed, FASTA or FASTQ (04_cleaning_primers.sh:371-388, 468-507)."""
    from . import primers
    t0 = time.time()
    c = primers.run(opt, device=device)
    rep = _report(c["n_in"], c["bp_in"], c["bp_out"], c["n_with"], 0, list(c["per_adapter"]),
                  np.array(list(c["per_adapter"].values()), dtype=np.int64), time.time() - t0, argv)
    rep["input"]["path1"] = opt["inputs"][0]
    rep["read_counts"]["output"] = c["n_written"]
    if opt["json"]:
        with open(opt["json"], "w") as fh:
            json.dump(rep, fh, indent=2)
    if not opt["quiet"]:
        print("This is orcdemux (cutadapt 4.9-compatible primer trimming on B200)")
        print("Command line parameters: " + " ".join(argv))
        print("=== Summary ===\n")
        print("Total reads processed:           %15s" % format(c["n_in"], ","))
        print("Reads with adapters:             %15s (%.1f%%)" % (format(c["n_with"], ","), 100.0 * c["n_with"] / max(c["n_in"], 1)))
        print("Reads written (passing filters): %15s (%.1f%%)\n" % (format(c["n_written"], ","), 100.0 * c["n_written"] / max(c["n_in"], 1)))
        for nm, k in c["per_adapter"].items():
            print("=== Adapter %s ===\n\nTrimmed: %d times\n" % (nm, k))
    return 0


def _report(n_in, bp_in, bp_out, n_with, n_rc, names, per_adapter, elapsed, argv):
    return {
        "tag": "Cutadapt report", "schema_version": [0, 3],
        "cutadapt_version": "4.9-compatible (orcdemux, B200)",
        "command_line_arguments": argv, "cores": 1,
        "input": {"path1": None, "path2": None, "paired": False},
        "read_counts": {"input": n_in, "filtered": {"too_short": None, "too_long": None, "too_many_n": None,
                                                    "too_many_expected_errors": None, "casava_filtered": None,
                                                    "discard_trimmed": None, "discard_untrimmed": None},
                        "output": n_in, "reverse_complemented": n_rc, "read1_with_adapter": n_with,
                        "read2_with_adapter": None},
        "basepair_counts": {"input": bp_in, "input_read1": bp_in, "input_read2": None, "quality_trimmed": None,
                            "quality_trimmed_read1": None, "quality_trimmed_read2": None, "poly_a_trimmed": None,
                            "output": bp_out, "output_read1": bp_out, "output_read2": None},
        "adapters_read1": [{"name": nm, "total_matches": int(c)} for nm, c in zip(names, per_adapter)],
        "adapters_read2": None, "elapsed_seconds": elapsed,
    }


class EndStats:
    """Per-adapter statistics of cutadapt's report (report.py / adapters.py EndStatistics as far as
    they derive from the match records): histogram of removed lengths split by error count, the
    expected number of random matches per length (gc_content 0.5), matches found on the reverse
    complement, and for 3' adapters the base preceding the match."""

    def __init__(self, names, seqs, kind, max_error_rate, indels, revcomp):
        self.names, self.seqs, self.kind = names, seqs, kind
        self.rate, self.indels, self.revcomp = max_error_rate, indels, revcomp
        self.front = kind in (ORC_FRONT, ORC_PREFIX)
        self.hist = [dict() for _ in names]            # removed length -> {errors: count}
        self.on_rc = np.zeros(len(names), dtype=np.int64)
        # 3' adapters: the base preceding the match, columns A C G T none/other (RemoveAfterMatch.adjacent_base)
        self.adjacent = np.zeros((len(names), 5), dtype=np.int64)

    _COMP = np.arange(256, dtype=np.uint8)
    for _a, _b in zip(b"ACGT", b"TGCA"):
        _COMP[_a] = _b
    _BASE_COL = np.full(256, 4, dtype=np.int64)
    for _i, _a in enumerate(b"ACGT"):
        _BASE_COL[_a] = _i

    def add(self, m, in_len, tb=None):
        """m: orc_match records of the round; in_len: read lengths before the round; tb: the batch
        (its text gives the bases next to 3' matches)."""
        has = m["adapter"] >= 0
        if not has.any():
            return
        a = m["adapter"][has].astype(np.int64)
        err = m["errors"][has].astype(np.int64)
        L = in_len[has].astype(np.int64)
        removed = m["query_stop"][has].astype(np.int64) if self.front else L - m["query_start"][has]
        self.on_rc += np.bincount(a[m["is_rc"][has] != 0], minlength=len(self.names))
        if not self.front and tb is not None:
            qs = m["query_start"][has].astype(np.int64)
            rc = m["is_rc"][has] != 0
            off = tb.offsets[:m.shape[0]][has].astype(np.int64)
            # read[qs - 1] in the orientation that matched: on the reverse complement that is the
            # complement of the stored base at n - qs
            pos = np.where(rc, off + (L - qs), off + qs - 1)
            ok = qs > 0
            base = tb.text[np.where(ok, pos, off)]
            base = np.where(rc, self._COMP[base], base)
            col = np.where(ok, self._BASE_COL[base], 4)
            np.add.at(self.adjacent, (a, col), 1)
        key = (a << 40) | (removed << 8) | err
        ks, cs = np.unique(key, return_counts=True)
        for k, c in zip(ks.tolist(), cs.tolist()):
            h = self.hist[k >> 40].setdefault((k >> 8) & 0xFFFFFFFF, {})
            h[k & 0xFF] = h.get(k & 0xFF, 0) + c

    def _end(self, i, n_reads):
        seq = self.seqs[i]
        walk = seq[::-1] if self.front else seq
        p, probs = 1.0, [1.0]
        for c in walk:
            p *= 0.25                                  # gc_content 0.5: every base 0.25
            probs.append(p)
        kmax = int(len(seq) * self.rate)
        lengths = []
        for ln in sorted(self.hist[i]):
            h = self.hist[i][ln]
            lengths.append({"len": ln, "expect": n_reads * probs[min(len(seq), ln)],
                            "counts": [h.get(e, 0) for e in range(max(kmax, max(h)) + 1)]})
        typ = {ORC_FRONT: "regular_five_prime", ORC_BACK: "regular_three_prime",
               ORC_PREFIX: "anchored_five_prime", ORC_SUFFIX: "anchored_three_prime"}[self.kind]
        return {"type": typ, "sequence": seq, "error_rate": self.rate, "indels": bool(self.indels),
                "error_lengths": [int(e / self.rate) - 1 if e else 0 for e in range(kmax + 1)][1:] + [len(seq)]
                if self.rate > 0 else [len(seq)],
                "matches": int(sum(sum(h.values()) for h in self.hist[i].values())),
                "adjacent_bases": None if self.front else
                {k: int(v) for k, v in zip(("A", "C", "G", "T", ""), self.adjacent[i])},
                "dominant_adjacent_base": None if self.front else self._dominant(i), "trimmed_lengths": lengths}

    def _dominant(self, i):
        """report.py: a base that precedes > 80 % of at least 20 matches is flagged."""
        total = int(self.adjacent[i].sum())
        if total < 20:
            return None
        j = int(np.argmax(self.adjacent[i][:4]))
        return "ACGT"[j] if self.adjacent[i][j] / total > 0.8 else None

    def text(self, n_reads, min_overlap):
        """The per-adapter sections of cutadapt's text report (report.py full_report)."""
        out = []
        typ = {ORC_FRONT: "regular 5'", ORC_BACK: "regular 3'", ORC_PREFIX: "anchored 5'",
               ORC_SUFFIX: "anchored 3'"}[self.kind]
        for i, nm in enumerate(self.names):
            end = self._end(i, n_reads)
            seq = self.seqs[i]
            out.append("=== Adapter %s ===\n" % nm)
            line = "Sequence: %s; Type: %s; Length: %d; Trimmed: %d times" % (seq, typ, len(seq), end["matches"])
            if self.revcomp:
                line += "; Reverse-complemented: %d times" % int(self.on_rc[i])
            out.append(line + "\n")
            out.append("Minimum overlap: %d" % min(min_overlap, len(seq)))
            out.append("No. of allowed errors:")
            lo, parts = 1, []
            for e, hi in enumerate(end["error_lengths"]):
                parts.append(("%d-%d bp: %d" % (lo, hi, e)) if hi > lo else ("%d bp: %d" % (hi, e)))
                lo = hi + 1
            out.append("; ".join(parts) + "\n")
            if end["matches"] == 0:
                continue
            if not self.front:
                tot = max(int(self.adjacent[i].sum()), 1)
                out.append("Bases preceding removed adapters:")
                for k, lab in enumerate(("A", "C", "G", "T", "none/other")):
                    out.append("  %s: %.1f%%" % (lab, 100.0 * self.adjacent[i][k] / tot))
                dom = self._dominant(i)
                if dom:
                    out.append("WARNING:\n    The adapter is preceded by '%s' extremely often.\n"
                               "    The provided adapter sequence could be incomplete at its 5' end." % dom)
                out.append("")
            out.append("Overview of removed sequences")
            out.append("length\tcount\texpect\tmax.err\terror counts")
            for t in end["trimmed_lengths"]:
                ln = t["len"]
                out.append("%d\t%d\t%.1f\t%d\t%s" % (ln, sum(t["counts"]), t["expect"],
                                                       int(self.rate * min(ln, len(seq))),
                                                       " ".join(str(c) for c in t["counts"])))
            out.append("")
        return "\n".join(out)

    def as_json(self, n_reads):
        out = []
        for i, nm in enumerate(self.names):
            end = self._end(i, n_reads)
            out.append({"name": nm, "total_matches": end["matches"],
                        "on_reverse_complement": int(self.on_rc[i]) if self.revcomp else None, "linked": False,
                        "five_prime_end": end if self.front else None,
                        "three_prime_end": None if self.front else end})
        return out


def _batch_shape(path: Optional[str] = None):
    """(reads, bytes) of one batch; ORCDEMUX_BATCH_READS shrinks it (tests, small-memory hosts).

    A small input gets small batches: the engine's device arenas and the page-locked buffers of reader and
    engine are sized by the batch, and page-locking memory costs about a second per gigabyte here (measured
    with ORCDEMUX_TIMING=1: of the 4.3 s of one round-2 call on a 100 MB bin, 3.8 s went into creating and
    releasing 1.3 GB of reader buffers) -- which the twelve round-2 calls of the script (02:94-102) pay twelve
    times.  So a batch is about a quarter of the input (at least 8 MB, at most 256 MB of text): a few batches
    in flight over the three slots, buffers in proportion to the work.  The estimate of the text size (four
    times a .gz) only has to be roughly right: the reader cuts batches by what fits."""
    env = os.environ.get("ORCDEMUX_BATCH_READS")
    reads = int(env) if env else 1 << 18
    if reads < 1:
        raise Unsupported("ORCDEMUX_BATCH_READS must be positive")
    nbytes = max(1 << 20, min(1 << 28, reads * 4096))
    if not env and path:
        try:
            size = os.path.getsize(path)
        except OSError:
            size = None
        if size is not None:
            text = size * (4 if path.endswith(".gz") else 1)
            want = max(1 << 23, min(1 << 28, (text // 4 + (1 << 20)) & ~((1 << 20) - 1)))
            if want < nbytes:
                nbytes = want
                reads = max(1 << 12, nbytes >> 10)
    return reads, nbytes


class _Phases:
    """ORCDEMUX_TIMING=1: wall time of the phases of one invocation on stderr (where do the seconds of a small
    call go: interpreter and library start-up, engine and buffers, streaming, closing the writers)."""
    def __init__(self):
        self.on = os.environ.get("ORCDEMUX_TIMING") == "1"
        self.t = time.time()
        self.parts = []

    def mark(self, name):
        if self.on:
            now = time.time()
            self.parts.append("%s %.3f s" % (name, now - self.t))
            self.t = now

    def done(self):
        if self.on:
            sys.stderr.write("orcdemux timing: " + ", ".join(self.parts) + "\n")


def _stream(reader, eng, writers, slots, on_result, rank=0, world=1):
    """reader -> GPU -> writers with `slots` batches between the stages: while batch k is on the
    GPU the bins of batches k-1 and k-2 are being deflated and the reader inflates k+1, k+2.
    A slot is submitted again only after the writer has let go of its result buffers.
    With world > 1 this rank only takes the batches dealt to it (shard.owner_of_batch): batch ids count
    the reader's batches, which are the same on every rank."""
    tickets = [None] * slots
    pending = []

    def drain(slot, tb):
        res = eng.wait(slot, copy=False)
        on_result(res, tb)
        tickets[slot] = writers.write_batch(res, members=eng.emit_gzip)

    k = 0
    for batch_id, tb in enumerate(reader):
        if world > 1 and batch_id % world != rank:
            continue
        slot = k % slots
        if tickets[slot] is not None:
            writers.wait(tickets[slot])
            tickets[slot] = None
        eng.submit(slot, tb)
        pending.append((slot, tb))
        if len(pending) > 1:
            drain(*pending.pop(0))
        k += 1
    while pending:
        drain(*pending.pop(0))


def run_single_round(opt, argv, device=0) -> int:
    kind = ORC_FRONT if opt["g"] else ORC_BACK
    names, seqs, anchored = _parse_adapter_specs(opt["g"] or opt["a"], kind)
    if anchored and opt["indels"]:
        raise Unsupported("anchored adapters need --no-indels (the indel variant is not built)")
    if anchored:
        kind = ORC_PREFIX if kind == ORC_FRONT else ORC_SUFFIX
    rnd = E.Round(names, seqs, kind, opt["e"], opt["O"], opt["indels"], opt["rc"], opt["action"])
    t0 = time.time()
    ph = _Phases()
    (max_reads, max_bytes), slots = _batch_shape(opt["inputs"][0]), 3
    threads = max(2, min(os.cpu_count() or 2, opt["cores"] if opt["cores"] > 0 else (os.cpu_count() or 2)))
    reader = F.FastqReader(opt["inputs"][0], max_reads, max_bytes, keep=3, ahead=2, threads=threads)
    paths = [opt["out"].replace("{name}", "unknown")] + [opt["out"].replace("{name}", n) for n in names]
    writers = F.BinWriters(paths, opt["level"], threads=threads)
    n_in = bp_in = bp_out = n_with = n_rc = 0
    per = np.zeros(len(names), dtype=np.int64)
    stats = EndStats(names, seqs, kind, opt["e"], opt["indels"], opt["rc"])

    def on_result(res, tb):
        nonlocal n_in, bp_in, bp_out, n_with, n_rc, per
        m = res.matches[0]
        n_in += res.n_reads
        bp_in += tb.total_bases()
        bp_out += int(res.out_len.sum(dtype=np.uint64))
        has = m["adapter"] >= 0
        n_with += int(has.sum())
        n_rc += int((m["is_rc"] != 0).sum())
        per += np.bincount(m["adapter"][has], minlength=len(names))
        stats.add(m, tb.lengths[:res.n_reads], tb)

    ph.mark("reader + writers")
    try:
        with E.Engine([rnd], device=device, max_reads=max_reads, max_bytes=max_bytes, n_slots=slots,
                      emit_fastq=True, want_matches=True) as eng:
            ph.mark("engine (%d reads, %d MB per batch)" % (max_reads, max_bytes >> 20))
            try:
                _stream(reader, eng, writers, slots, on_result)
                ph.mark("stream")
            finally:
                writers.close()         # drains: the result buffers belong to the engine
                ph.mark("writers closed")
    finally:
        writers.close()
        reader.close()
    ph.mark("engine closed")
    ph.done()
    rep = _report(n_in, bp_in, bp_out, n_with, n_rc, names, per, time.time() - t0, argv)
    rep["input"]["path1"] = opt["inputs"][0]
    rep["adapters_read1"] = stats.as_json(n_in)
    if opt["json"]:
        with open(opt["json"], "w") as fh:
            json.dump(rep, fh, indent=2)
    if not opt["quiet"]:
        el = rep["elapsed_seconds"]
        print("This is orcdemux (cutadapt 4.9-compatible demultiplexing on B200)")
        print("Command line parameters: " + " ".join(argv))
        print("Processing single-end reads on 1 GPU, %d host threads ..." % threads)
        print("Finished in %.3f s (%.3f us/read; %.2f M reads/minute).\n" %
              (el, 1e6 * el / max(n_in, 1), n_in / max(el, 1e-9) * 60e-6))
        print("=== Summary ===\n")
        print("Total reads processed:           %15s" % format(n_in, ","))
        print("Reads with adapters:             %15s (%.1f%%)" % (format(n_with, ","), 100.0 * n_with / max(n_in, 1)))
        if opt["rc"]:
            print("Reverse-complemented:            %15s (%.1f%%)" % (format(n_rc, ","), 100.0 * n_rc / max(n_in, 1)))
        print("Reads written (passing filters): %15s (100.0%%)\n" % format(n_in, ","))
        print("Total basepairs processed: %15s bp" % format(bp_in, ","))
        print("Total written (filtered):  %15s bp (%.1f%%)\n" % (format(bp_out, ","), 100.0 * bp_out / max(bp_in, 1)))
        print(stats.text(n_in, opt["O"]))
    return 0


def dataset_name(infile: str) -> str:
    """02_cutadapt_loop.sh:25-35."""
    ds = os.path.basename(infile)
    if ds.startswith("pychopped_"):
        ds = ds[len("pychopped_"):]
    for suf in (".fastq.gz", ".fastq", ".fq.gz", ".fq", ".gz", "_pass"):
        if ds.endswith(suf):
            ds = ds[:-len(suf)]
    return ds


def run_two_round(args: List[str], device=0) -> int:
    """Both rounds fused: the file tree 02_cutadapt_loop.sh leaves in demuxed/SP27/."""
    import argparse
    ap = argparse.ArgumentParser(prog="orcdemux two-round")
    ap.add_argument("input")
    ap.add_argument("--sp5", required=True, help="FASTA of the 5' SP5 adapters (-g file:)")
    ap.add_argument("--sp27", required=True, help="FASTA of the 3' SP27 rc adapters (-a file:)")
    ap.add_argument("-e", type=float, default=0.1)
    ap.add_argument("-O", type=int, default=3)
    ap.add_argument("--outdir", default=None)
    ap.add_argument("--keep-unknown", action="store_true")
    ap.add_argument("--keep-invalid", action="store_true", help="keep SP27_009..012 combinations")
    ap.add_argument("--no-gzip", action="store_true")
    ap.add_argument("-j", type=int, default=8)
    ap.add_argument("--compression-level", type=int, default=1, help="gzip level of the bin files with --host-gzip (cutadapt 4.x default: 1); the members coded on the GPU "
                         "are Huffman-only whatever the level")
    ap.add_argument("--host-gzip", action="store_true",
                    help="deflate the bin files with zlib on the host threads; default: the GPU codes every bin of a "
                         "batch as a gzip member (dynamic Huffman, literals only) and only those bytes come back")
    ap.add_argument("--gpus", type=int, default=1,
                    help="GPUs of this node to shard the batches over, one process each (the same happens under "
                         "torchrun, whose RANK / WORLD_SIZE are honoured)")
    a = ap.parse_args(args)
    from . import shard
    rank, world, local_rank = shard.dist_env()
    if world == 1 and a.gpus > 1:
        # no torchrun around us: be the launcher; the ranks run this same command
        return shard.spawn_ranks(a.gpus, ["two-round"] + list(args))
    share = False
    if world > 1:
        # one GPU per rank; with fewer GPUs than ranks (a test box) ranks share devices and the count gather
        # goes over gloo, because NCCL wants a device per rank
        import torch
        n_dev = max(1, torch.cuda.device_count())
        share = world > n_dev
        device = local_rank % n_dev
    n5, s5 = F.read_adapters_fasta(a.sp5)
    n27, s27 = F.read_adapters_fasta(a.sp27)
    ds = dataset_name(a.input)
    outdir = a.outdir or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(a.input))), "demuxed")
    os.makedirs(os.path.join(outdir, "SP5"), exist_ok=True)
    os.makedirs(os.path.join(outdir, "SP27"), exist_ok=True)
    rounds = [E.Round(n5, s5, ORC_FRONT, a.e, a.O, True, True), E.Round(n27, s27, ORC_BACK, a.e, a.O, True, True)]
    ext = ".fastq" if a.no_gzip else ".fastq.gz"
    n_bins = (len(n5) + 1) * (len(n27) + 1)
    drop = np.zeros(n_bins, dtype=np.uint8)
    paths: List[Optional[str]] = [None] * n_bins
    invalid = {"SP27_009", "SP27_010", "SP27_011", "SP27_012"}      # 02:114-118
    for b in range(n_bins):
        i5, i27 = b % (len(n5) + 1) - 1, b // (len(n5) + 1) - 1
        nm5 = n5[i5] if i5 >= 0 else "unknown"
        nm27 = n27[i27] if i27 >= 0 else "unknown"
        unknown = i5 < 0 or i27 < 0
        if (unknown and not a.keep_unknown) or (nm27 in invalid and not a.keep_invalid) or i5 < 0:
            drop[b] = 1
            continue
        paths[b] = os.path.join(outdir, "SP27", "%s_%s_%s%s" % (nm27, nm5, ds, ext))
    (max_reads, max_bytes), slots = _batch_shape(a.input), 3
    threads = max(2, min(os.cpu_count() or 2, a.j) // max(world, 1))
    reader = F.FastqReader(a.input, max_reads, max_bytes, keep=3, ahead=2, threads=threads)
    # with several ranks every rank writes part files (+ an index of their chunks); rank 0 stitches them
    my_paths = [shard.part_path(p, rank) if (p and world > 1) else p for p in paths]
    writers = F.BinWriters(my_paths, a.compression_level, threads=threads, index=world > 1)
    t0 = time.time()
    n_in = 0

    def on_result(res, tb):
        nonlocal n_in
        n_in += res.n_reads

    try:
        with E.Engine(rounds, device=device, max_reads=max_reads, max_bytes=max_bytes, n_slots=slots,
                      emit_fastq=True, want_matches=False, drop_bins=drop,
                      emit_gzip=not (a.no_gzip or a.host_gzip)) as eng:
            try:
                _stream(reader, eng, writers, slots, on_result, rank, world)
            finally:
                writers.close()         # drains: the result buffers belong to the engine
            counts = eng.counts().astype(np.int64)
    finally:
        writers.close()
        reader.close()
    if world > 1:
        # the only exchanges: the per-bin count gather, and a barrier before the part files are stitched
        import torch.distributed as dist
        shard.init_process_group(not share, local_rank)
        tot = shard.gather_counts(np.concatenate([counts, [n_in]]))
        counts, n_in = tot[:-1], int(tot[-1])
        dist.barrier()
        if rank == 0:
            shard.merge_part_files(paths, world, a.compression_level)
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        with open(os.path.join(outdir, "SP27", "orcdemux_%s.json" % ds), "w") as fh:
            json.dump({"dataset": ds, "reads": n_in, "elapsed_seconds": time.time() - t0, "gpus": world,
                       "bins": {os.path.basename(p): int(counts[b]) for b, p in enumerate(paths) if p}}, fh, indent=1)
        print("Demultiplexing complete! %d reads on %d GPU%s, results in: %s" % (n_in, world, "s" if world > 1 else "", outdir))
    return 0


def main(argv: Optional[List[str]] = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    try:
        if argv and argv[0] == "two-round":
            return run_two_round(argv[1:])
        if argv and argv[0] == "orient":
            from . import orient
            return orient.main(argv[1:])
        if argv and argv[0] in ("--version",):
            print("4.9 (orcdemux)")
            return 0
        opt = parse_cutadapt_argv(argv)
        if "{name}" not in opt["out"]:
            return run_primer_trim(opt, argv)
        return run_single_round(opt, argv)
    except Unsupported as e:
        sys.stderr.write("orcdemux: unsupported: %s\n" % e)
        return 2
    except (E.OrcError, ValueError, OSError) as e:
        sys.stderr.write("orcdemux: error: %s\n" % e)
        return 1


if __name__ == "__main__":
    sys.exit(main())
