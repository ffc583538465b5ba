"""Pure-Python restatement of cutadapt 4.9's matching rules -- TEST INFRASTRUCTURE ONLY.

*** PARITY UNPINNED *** (see oracle/cutadapt_oracle.c header): cutadapt is not vendored
under /root/reference (it is called as a subprocess at scripts/02_cutadapt_loop.sh:64-72
and :94-102) and cannot be installed here.  This module restates SURVEY.md section 8(c)
R0..R10 a second time, with a different structure from the C oracle (the full DP matrix
is materialised, no Ukkonen band, the path is kept per cell), so that the two
restatements check each other.  Pure-Python loops: small cases only.

Only tests/ may import this module.
"""
from __future__ import annotations

REF_START, QUERY_START, REF_END, QUERY_STOP = 1, 2, 4, 8
FRONT = REF_START | QUERY_START | QUERY_STOP      # cutadapt align.py Where.FRONT  (-g)
BACK = QUERY_START | QUERY_STOP | REF_END         # Where.BACK   (-a)
PREFIX = QUERY_STOP                               # Where.PREFIX (-g ^)
SUFFIX = QUERY_START                              # Where.SUFFIX (-a ...$)

_COMP = {}
for _a, _b in zip("ACGTUMRWSYKVHDBN", "TGCAAKYWSRMBDHVN"):
    _COMP[_a] = _b
    _COMP[_a.lower()] = _b.lower()


def reverse_complement(seq: str) -> str:
    """dnaio SequenceRecord.reverse_complement(): IUPAC complement, case kept."""
    return "".join(_COMP.get(c, c) for c in reversed(seq))


_IUPAC = {"X": 0, "A": 1, "C": 2, "M": 3, "G": 4, "R": 5, "S": 6, "V": 7, "T": 8, "U": 8, "W": 9, "Y": 10,
          "H": 11, "K": 12, "D": 13, "B": 14, "N": 15}
_ACGT = {"A": 1, "C": 2, "G": 4, "T": 8, "U": 8}


def locate(ref: str, query: str, max_error_rate: float, flags: int, min_overlap: int = 1,
           indel_cost: int = 1, wildcard_ref: bool = False):
    """Aligner.locate for ASCII comparison (adapter made of ACGT only, no read wildcards), or with
    wildcard_ref for an adapter with IUPAC characters: adapter through _iupac_table(), read through
    _acgt_table(), characters match when the masks intersect, and the N of the adapter do not count
    towards the length the error rate applies to (_set_reference n_counts, R5/R6).

    Returns (ref_start, ref_stop, query_start, query_stop, score, errors) or None.
    Follows SURVEY 8(c) R1-R7 with the full matrix.
    """
    m, n = len(ref), len(query)
    if wildcard_ref:
        n_before = [sum(1 for c in ref[:i] if c in "nN") for i in range(m + 1)]     # N in ref[:i]
        ref = [_IUPAC.get(c.upper(), 0) for c in ref]
        query = [_ACGT.get(c.upper(), 0) for c in query]
        same = lambda a, b: (a & b) != 0
        eff_row = lambda length: (length - n_before[length]) if length < m else m - n_before[m]
        eff_col = lambda length, i: (length - (n_before[i] - n_before[i - length])) if length < m else m - n_before[m]
    else:
        same = lambda a, b: a == b
        eff_row = lambda length: length
        eff_col = lambda length, i: length
    k = int(max_error_rate * m)
    max_n = n if flags & QUERY_START else min(n, m + k)
    min_n = 0 if flags & QUERY_STOP else max(0, n - m - k)
    # cell = (cost, score, origin)
    col = []
    for i in range(m + 1):
        if (flags & REF_START) and (flags & QUERY_START):
            col.append((min(i, min_n) * indel_cost, 0, min_n - i))
        elif flags & REF_START:
            col.append((min_n * indel_cost, 0, min(0, min_n - i)))
        elif flags & QUERY_START:
            col.append((i * indel_cost, 0, max(0, min_n - i)))
        else:
            col.append((max(i, min_n) * indel_cost, 0, 0))
    best = None  # (score, cost, origin, ref_stop, query_stop)
    broke = False
    for j in range(min_n + 1, max_n + 1):
        new = [None] * (m + 1)
        if flags & QUERY_START:
            new[0] = (col[0][0], col[0][1], j)
        else:
            new[0] = (j * indel_cost, col[0][1], col[0][2])
        for i in range(1, m + 1):
            diag, left, up = col[i - 1], col[i], new[i - 1]
            if same(ref[i - 1], query[j - 1]):
                new[i] = (diag[0], diag[1] + 1, diag[2])
            else:
                c_diag, c_del, c_ins = diag[0] + 1, left[0] + indel_cost, up[0] + indel_cost
                if c_diag <= c_del and c_diag <= c_ins:
                    new[i] = (c_diag, diag[1] - 1, diag[2])
                elif c_ins <= c_del:
                    new[i] = (c_ins, up[1] - 2, up[2])
                else:
                    new[i] = (c_del, left[1] - 2, left[2])
        col = new
        cost, score, origin = col[m]
        if cost <= k and (flags & QUERY_STOP):
            length = m + min(origin, 0)
            ok = length >= min_overlap and cost <= eff_row(length) * max_error_rate
            if ok:
                if best is None:
                    upd = True
                else:
                    best_length = m + min(best[2], 0)
                    upd = (origin <= best[2] + m // 2 and score > best[0]) or \
                          (length > best_length and score > best[0])
                if upd:
                    best = (score, cost, origin, m, j)
                    if cost == 0 and origin >= 0:
                        broke = True
                        break
    if max_n == n and not broke:
        first_i = 0 if flags & REF_END else m
        for i in range(m, first_i - 1, -1):
            cost, score, origin = col[i]
            length = i + min(origin, 0)
            ok = length >= min_overlap and cost <= eff_col(length, i) * max_error_rate
            # R6 has no "best unset" clause: an unset best stands there with score 0 and cost m + n + 1, so a
            # candidate with a negative score (error rates of 0.5 and more) does not replace it
            b_score, b_cost = (0, m + n + 1) if best is None else (best[0], best[1])
            if ok and (score > b_score or (score == b_score and cost < b_cost)):
                best = (score, cost, origin, i, n)
    if best is None:
        return None
    score, cost, origin, ref_stop, query_stop = best
    if origin >= 0:
        return (0, ref_stop, origin, query_stop, score, cost)
    return (-origin, ref_stop, 0, query_stop, score, cost)


class Adapter:
    """adapters.py SingleAdapter subset: ACGT-only sequences, types front/back."""

    def __init__(self, name, sequence, where, max_errors=0.1, min_overlap=3, indels=True):
        self.name = name
        self.sequence = sequence.upper().replace("U", "T")
        self.where = where
        if max_errors >= 1:
            max_errors /= len(self.sequence)
        self.max_error_rate = max_errors
        self.min_overlap = min(min_overlap, len(self.sequence))
        self.indel_cost = 1 if indels else 100000

    def match_to(self, sequence: str):
        return locate(self.sequence, sequence.upper(), self.max_error_rate, self.where,
                      self.min_overlap, self.indel_cost)


def best_of(adapters, sequence):
    """MultipleAdapters.match_to (R8)."""
    best = None
    for idx, ad in enumerate(adapters):
        t = ad.match_to(sequence)
        if t is None:
            continue
        if best is None or t[4] > best[1][4] or (t[4] == best[1][4] and t[5] < best[1][5]):
            best = (idx, t)
    return best


def round_read(adapters, revcomp, name, seq, qual):
    """ReverseComplementer(AdapterCutter(times=1, action='trim')) (R9, R10).

    Returns (adapter_index_or_-1, is_rc, tuple_or_None, name, seq, qual) of the output read.
    """
    fwd = best_of(adapters, seq)
    rev = None
    rc_seq = rc_qual = None
    if revcomp:
        rc_seq, rc_qual = reverse_complement(seq), qual[::-1]
        rev = best_of(adapters, rc_seq)
    fs = fwd[1][4] if fwd else 0
    rs = rev[1][4] if rev else 0
    use_rc = bool(revcomp) and rs > fs
    if use_rc:
        chosen, s, q, name = rev, rc_seq, rc_qual, name + " rc"
    else:
        chosen, s, q = fwd, seq, qual
    if chosen is None:
        return (-1, use_rc, None, name, s, q)
    idx, t = chosen
    if adapters[idx].where in (FRONT, PREFIX):
        s, q = s[t[3]:], q[t[3]:]
    else:
        s, q = s[:t[2]], q[:t[2]]
    return (idx, use_rc, t, name, s, q)
