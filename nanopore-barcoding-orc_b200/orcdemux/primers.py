"""Primer trimming with one output file: the cutadapt call shapes of the reference's
scripts/04_cleaning_primers.sh.

  round 1 (04:371-388)  cutadapt -j N -g FWD...REV [-g FWD2...REV2 ...]
                                 --untrimmed-output=UNTRIMMED.fasta -o TRIMMED.fasta CONSENSUS.fasta
  round 2 (04:468-507)  cutadapt -j N -g FWD [-g ...] -a REV [-a ...] -o OUT.fasta UNTRIMMED.fasta

A linked adapter given with -g is cutadapt's LinkedAdapter with both parts required and neither
anchored (parser.py _parse_linked): the 5' part is located in the read, the 3' part in what is left
behind it, and the read counts as trimmed only when both are found; score and errors of the pair are
the sums (LinkedMatch), and of several pairs the best wins by MultipleAdapters' rule (score, then
fewer errors, then order).  Each pair is one two-round pass of the GPU engine (5' round, 3' round on
the remainder) -- the same kernels as the demultiplexer, with IUPAC wildcards in the primers; the
selection over the pairs and the FASTA/FASTQ files are handled here.  These inputs are consensus
sequences (thousands of records), so the records are read and written in Python; the streaming
native reader/writers are the demultiplexer's.

Plain adapters, also 5' and 3' ones side by side (round 2 of the script), keep cutadapt's one match
per read: the best over all adapters in command-line order, found by comparing the winners of the GPU
passes (one per adapter type and wildcard class, select_best).

Not supported here (exit 2): --rc, anchored parts, per-adapter parameters (;e=...),
`required`/`optional`, -n / --times above 1."""
from __future__ import annotations

import gzip
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import engine as E
from . import synth
from .lib import MATCH_DTYPE, ORC_BACK, ORC_FRONT

MAX_PER_PASS = 32                                # adapters per GPU pass (csrc/orc_core.cuh MAX_AD)
MAX_PER_PASS_LONG = 16                           # of adapters over 64 nt (MAX_AD_LONG)
Record = Tuple[str, str, Optional[str]]          # name (header without @ or >), sequence, qualities or None


class Unsupported(Exception):
    pass


def parse_linked_specs(specs: Sequence[str]):
    """-g values -> [(name, fwd, rev)] if every one is linked, [] if none is."""
    out = []
    n_linked = 0
    for i, spec in enumerate(specs):
        name, s = None, spec
        if "=" in s and "..." in s.split("=", 1)[1]:
            name, s = s.split("=", 1)
        if "..." not in s:
            continue
        n_linked += 1
        fwd, rev = s.split("...", 1)
        for part in (fwd, rev):
            if not part or any(c in part for c in "^$;{}[]. ") or part.startswith("file:"):
                raise Unsupported("linked adapter syntax beyond SEQ1...SEQ2: %r" % spec)
        out.append((name or str(i + 1), fwd.upper().replace("U", "T"), rev.upper().replace("U", "T")))
    if n_linked and n_linked != len(specs):
        raise Unsupported("linked and plain adapters in one invocation")
    return out


def _open(path: str, mode: str):
    return gzip.open(path, mode) if path.endswith(".gz") else open(path, mode)


def read_sequences(path: str) -> Tuple[List[Record], str]:
    """FASTA (multi-line allowed) or FASTQ (four-line), optionally gzipped -> (records, format)."""
    with _open(path, "rt") as fh:
        text = fh.read()
    if not text.strip():
        return [], "fasta"
    first = text.lstrip()[0]
    recs: List[Record] = []
    if first == ">":
        name, parts = None, []
        for line in text.splitlines():
            if line.startswith(">"):
                if name is not None:
                    recs.append((name, "".join(parts), None))
                name, parts = line[1:], []
            elif name is not None:
                parts.append(line.strip())
        if name is not None:
            recs.append((name, "".join(parts), None))
        return recs, "fasta"
    if first != "@":
        raise ValueError("%s is neither FASTA nor FASTQ" % path)
    lines = text.splitlines()
    if len(lines) % 4:
        raise ValueError("%s: FASTQ record cut short" % path)
    for i in range(0, len(lines), 4):
        if not lines[i].startswith("@") or not lines[i + 2].startswith("+") or len(lines[i + 1]) != len(lines[i + 3]):
            raise ValueError("%s: malformed FASTQ record at line %d" % (path, i + 1))
        recs.append((lines[i][1:], lines[i + 1], lines[i + 3]))
    return recs, "fastq"


def output_format(path: str, input_format: str) -> str:
    base = path[:-3] if path.endswith(".gz") else path
    ext = os.path.splitext(base)[1].lower()
    if ext in (".fasta", ".fa", ".fna"):
        return "fasta"
    if ext in (".fastq", ".fq"):
        return "fastq"
    return input_format


def write_sequences(path: str, recs: Sequence[Record], fmt: str) -> None:
    with _open(path, "wt") as fh:
        for name, seq, qual in recs:
            if fmt == "fasta":
                fh.write(">%s\n%s\n" % (name, seq))          # dnaio FastaWriter: one line per sequence
            else:
                if qual is None:
                    raise Unsupported("FASTQ output from FASTA input (no qualities)")
                fh.write("@%s\n%s\n+\n%s\n" % (name, seq, qual))


def select_linked(m0s: Sequence[np.ndarray], m1s: Sequence[np.ndarray]) -> np.ndarray:
    """Per read the index of the winning linked pair, -1 if no pair has both parts.
    MultipleAdapters.match_to over LinkedMatch objects: higher summed score, then fewer summed
    errors, then the first in command-line order."""
    n = m0s[0].shape[0]
    best = np.full(n, -1, dtype=np.int32)
    bscore = np.zeros(n, dtype=np.int64)
    berr = np.zeros(n, dtype=np.int64)
    for i, (m0, m1) in enumerate(zip(m0s, m1s)):
        ok = (m0["adapter"] >= 0) & (m1["adapter"] >= 0)
        score = m0["score"].astype(np.int64) + m1["score"]
        err = m0["errors"].astype(np.int64) + m1["errors"]
        take = ok & ((best < 0) | (score > bscore) | ((score == bscore) & (err < berr)))
        best[take] = i
        bscore[take] = score[take]
        berr[take] = err[take]
    return best


def trim_linked(recs: Sequence[Record], m0s, m1s, best: np.ndarray):
    """-> (trimmed-or-unchanged records in input order, boolean mask of the trimmed ones)."""
    out: List[Record] = []
    for r, (name, seq, qual) in enumerate(recs):
        i = int(best[r])
        if i < 0:
            out.append((name, seq, qual))
            continue
        a = int(m0s[i]["query_stop"][r])                    # RemoveBeforeMatch
        b = a + int(m1s[i]["query_start"][r])               # RemoveAfterMatch on the remainder
        out.append((name, seq[a:b], qual[a:b] if qual is not None else None))
    return out, best >= 0


def trim_single(recs: Sequence[Record], m: np.ndarray, front: bool):
    out: List[Record] = []
    for r, (name, seq, qual) in enumerate(recs):
        if m["adapter"][r] < 0:
            out.append((name, seq, qual))
        elif front:
            a = int(m["query_stop"][r])
            out.append((name, seq[a:], qual[a:] if qual is not None else None))
        else:
            b = int(m["query_start"][r])
            out.append((name, seq[:b], qual[:b] if qual is not None else None))
    return out, m["adapter"] >= 0


def select_best(ms: Sequence[np.ndarray], poss: Sequence[np.ndarray]):
    """Plain adapters of both types side by side (04:468-507 gives the forward primers with -g and the
    reverse ones with -a): cutadapt keeps ONE match per read, the best over all adapters in
    command-line order (MultipleAdapters.match_to: higher score, then fewer errors, then the first).
    The adapters are matched in groups on the GPU (one pass per adapter type, and per pass either all or
    none with IUPAC wildcards); the overall winner is the winner of its group, so the group results are
    compared by the same rule.  poss[g] maps an adapter index of group g to its place on the command
    line.  -> (winning group per read or -1, its command-line position or -1)."""
    n = ms[0].shape[0]
    best = np.full(n, -1, dtype=np.int32)
    bpos = np.full(n, -1, dtype=np.int64)
    bs = np.zeros(n, dtype=np.int64)
    be = np.zeros(n, dtype=np.int64)
    for g, (m, pos) in enumerate(zip(ms, poss)):
        ok = m["adapter"] >= 0
        p = pos[np.where(ok, m["adapter"], 0)]
        s_, e_ = m["score"].astype(np.int64), m["errors"].astype(np.int64)
        take = ok & ((best < 0) | (s_ > bs) | ((s_ == bs) & (e_ < be)) | ((s_ == bs) & (e_ == be) & (p < bpos)))
        best[take] = g
        bpos[take] = p[take]
        bs[take] = s_[take]
        be[take] = e_[take]
    return best, bpos


def _match_batches(rounds, recs: Sequence[Record], device: int, batch: int = 1 << 16):
    """Run `rounds` over the records on the GPU -> one match-record array per round."""
    sets = []
    for lo in range(0, len(recs), batch):
        part = recs[lo:lo + batch]
        sets.append(synth.from_records([(nm, sq, q if q is not None else "I" * len(sq)) for nm, sq, q in part]))
    outs: List[List[np.ndarray]] = [[] for _ in rounds]
    with E.Engine(rounds, device=device, max_reads=max(rs.n_reads for rs in sets),
                  max_bytes=max(int(rs.seq.shape[0]) for rs in sets) + 64,
                  max_name_bytes=max(int(rs.names.shape[0]) for rs in sets) + 64, n_slots=1,
                  emit_fastq=False, want_matches=True) as eng:
        for rs in sets:
            res = eng.run(rs)
            for k in range(len(rounds)):
                outs[k].append(np.array(res.matches[k], copy=True))
    return [np.concatenate(x) for x in outs]


def run(opt, device: int = 0):
    """opt: the dict of cli.parse_cutadapt_argv.  Returns the counters for the report."""
    if opt["rc"]:
        raise Unsupported("--rc together with a single output file")
    recs, fmt = read_sequences(opt["inputs"][0])
    pairs = parse_linked_specs(opt["g"])
    e, ov, indels = opt["e"], opt["O"], opt["indels"]
    if pairs:
        if opt["a"]:
            raise Unsupported("-a next to linked -g adapters")
        m0s, m1s = [], []
        for name, fwd, rev in pairs:
            rounds = [E.Round([name], [fwd], ORC_FRONT, e, ov, indels, False),
                      E.Round([name], [rev], ORC_BACK, e, ov, indels, False)]
            m0, m1 = _match_batches(rounds, recs, device) if recs else (np.zeros(0, MATCH_DTYPE),) * 2
            m0s.append(m0)
            m1s.append(m1)
        best = select_linked(m0s, m1s) if recs else np.zeros(0, np.int32)
        out, trimmed = trim_linked(recs, m0s, m1s, best)
        per_adapter = {p[0]: int((best == i).sum()) for i, p in enumerate(pairs)}
    else:
        from .cli import _parse_adapter_specs
        order = opt["order"] or ["g"] * len(opt["g"]) + ["a"] * len(opt["a"])
        it = {"g": iter(opt["g"]), "a": iter(opt["a"])}
        groups = {}                         # (adapter type, has wildcards, over 64 nt) -> names, sequences, positions
        all_names = []
        for t in order:
            kind = ORC_FRONT if t == "g" else ORC_BACK
            nm, sq, anchored = _parse_adapter_specs([next(it[t])], kind)
            if anchored:
                raise Unsupported("anchored adapters together with a single output file")
            for a_, b_ in zip(nm, sq):
                pos = len(all_names)
                all_names.append(str(pos + 1) if (a_ == "1" and len(nm) == 1) else a_)
                g = groups.setdefault((kind, any(c not in "ACGT" for c in b_), len(b_) > 64), ([], [], []))
                g[0].append(all_names[-1]); g[1].append(b_); g[2].append(pos)
        kinds, ms, poss = [], [], []
        for (kind, _, long_), (nm, sq, ps) in groups.items():
            # the kernels take up to 32 adapters per round, 16 of those over 64 nt (the cell-by-cell path)
            per = MAX_PER_PASS_LONG if long_ else MAX_PER_PASS
            for lo in range(0, len(nm), per):
                hi = lo + per
                kinds.append(kind)
                poss.append(np.array(ps[lo:hi], dtype=np.int64))
                ms.append(_match_batches([E.Round(nm[lo:hi], sq[lo:hi], kind, e, ov, indels, False)], recs, device)[0]
                          if recs else np.zeros(0, MATCH_DTYPE))
        best, bpos = select_best(ms, poss)
        out = list(recs)
        for g, kind in enumerate(kinds):
            mg = ms[g].copy()
            mg["adapter"][best != g] = -1
            out, _ = trim_single(out, mg, kind == ORC_FRONT)
        trimmed = best >= 0
        per_adapter = {nm: int((bpos == i).sum()) for i, nm in enumerate(all_names)}
    ofmt = output_format(opt["out"], fmt)
    if opt["untrimmed_output"]:
        write_sequences(opt["out"], [x for x, t in zip(out, trimmed) if t], ofmt)
        write_sequences(opt["untrimmed_output"], [x for x, t in zip(out, trimmed) if not t],
                        output_format(opt["untrimmed_output"], fmt))
        written = int(trimmed.sum())
    elif opt["discard_untrimmed"]:
        write_sequences(opt["out"], [x for x, t in zip(out, trimmed) if t], ofmt)
        written = int(trimmed.sum())
    else:
        write_sequences(opt["out"], out, ofmt)
        written = len(out)
    return {"n_in": len(recs), "n_with": int(trimmed.sum()), "n_written": written,
            "bp_in": sum(len(x[1]) for x in recs), "bp_out": sum(len(x[1]) for x, t in zip(out, trimmed)
                                                                 if t or not (opt["untrimmed_output"] or opt["discard_untrimmed"])),
            "per_adapter": per_adapter}
