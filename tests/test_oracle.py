"""The oracle against the known-answer vectors the reference's domain offers.

The reference (a shell pipeline around un-vendored cutadapt 4.9) ships no tests, so parity
with real cutadapt is UNPINNED; these vectors are (i) the examples of cutadapt's user guide
for regular 3'/5' adapters, (ii) cases derivable by hand from the published algorithm
(SURVEY.md 8c), committed as tests/golden/kat.json by tests/golden/make_golden.py.  Three
independent restatements (C banded, C unpruned, pure-Python full matrix) must agree."""
import json
import os
import random

import numpy as np
import pytest

import helpers as H
import oracle
from oracle import pyoracle as P
from orcdemux import m13, synth

FRONT, BACK, PREFIX, SUFFIX = 11, 14, 8, 2
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "kat.json")


def test_golden_locate_vectors():
    with open(GOLDEN) as fh:
        kat = json.load(fh)
    assert len(kat["locate"]) >= 30
    for v in kat["locate"]:
        exp = tuple(v["expect"]) if v["expect"] is not None else None
        args = (v["ref"], v["query"], v["rate"], v["flags"], v["min_overlap"], v.get("indel_cost", 1))
        assert oracle.locate(*args) == exp, v
        assert oracle.locate(*args, unpruned=True) == exp, v
        assert P.locate(*args) == exp, v


def test_golden_wildcard_vectors():
    """Adapters with IUPAC wildcards: cutadapt's own known-answer tests for N (recalled) and
    hand-derived mask cases, on the C oracle (banded and unpruned) and the Python restatement."""
    with open(GOLDEN) as fh:
        kat = json.load(fh)
    assert len(kat["locate_wildcard"]) >= 16
    for v in kat["locate_wildcard"]:
        exp = tuple(v["expect"]) if v["expect"] is not None else None
        args = (v["ref"], v["query"], v["rate"], v["flags"], v["min_overlap"], 1)
        assert oracle.locate(*args, wildcard_ref=True) == exp, v
        assert oracle.locate(*args, wildcard_ref=True, unpruned=True) == exp, v
        assert P.locate(*args, wildcard_ref=True) == exp, v


def test_wildcard_c_vs_python_random():
    import random
    rnd = random.Random(515)
    n_match = 0
    for _ in range(1500):
        m = rnd.randint(3, 24)
        ref = "".join(rnd.choice("ACGTACGTACGTRYSWKMBDHVNNNX") for _ in range(m))
        inst = "".join(rnd.choice({"R": "AG", "Y": "CT", "S": "CG", "W": "AT", "K": "GT", "M": "AC", "B": "CGT",
                                   "D": "AGT", "H": "ACT", "V": "ACG", "N": "ACGT", "X": "ACGT"}.get(c, c)) for c in ref)
        cut = rnd.choice([inst, inst[rnd.randint(0, m - 1):], inst[:rnd.randint(1, m)]])
        if rnd.random() < 0.5 and cut:
            p = rnd.randrange(len(cut))
            cut = cut[:p] + rnd.choice(["", rnd.choice("ACGTNU"), cut[p] + rnd.choice("ACGT")]) + cut[p + 1:]
        q = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 12))) + cut + \
            "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 12)))
        if rnd.random() < 0.3:
            q = cut
        rate = rnd.choice([0.0, 0.1, 0.2, 0.34])
        flags = rnd.choice([FRONT, BACK, FRONT, BACK, PREFIX, SUFFIX, 15])
        mo, ic = rnd.randint(1, 5), rnd.choice([1, 1, 100000])
        a = oracle.locate(ref, q, rate, flags, mo, ic, wildcard_ref=True)
        assert a == oracle.locate(ref, q, rate, flags, mo, ic, wildcard_ref=True, unpruned=True) \
            == P.locate(ref, q, rate, flags, mo, ic, wildcard_ref=True), (ref, q, rate, flags, mo, ic)
        n_match += a is not None
    assert n_match > 300


def test_golden_read_vectors():
    with open(GOLDEN) as fh:
        kat = json.load(fh)
    rounds = H.m13_rounds()
    names5 = [n for n, _ in m13.sp5_forward()]
    names27 = [n for n, _ in m13.sp27_reverse_rc()]
    recs = [(v["name"], v["seq"], "I" * len(v["seq"])) for v in kat["reads"]]
    rs = synth.from_records(recs)
    rec0, rec1, oseq, oqual, olen = H.run_oracle(rounds, rs, n_threads=2)
    for i, v in enumerate(kat["reads"]):
        got5 = names5[rec0["adapter"][i]] if rec0["adapter"][i] >= 0 else "unknown"
        got27 = names27[rec1["adapter"][i]] if rec1["adapter"][i] >= 0 else "unknown"
        assert got5 == v["sp5"], (v["name"], got5)
        assert got27 == v["sp27"], (v["name"], got27)
        if "r1" in v:
            assert [int(rec0[f][i]) for f in H.FIELDS[1:]] == v["r1"], v["name"]
        if "r2" in v:
            assert [int(rec1[f][i]) for f in H.FIELDS[1:]] == v["r2"], v["name"]
        o = int(rs.offsets[i])
        assert oseq[o:o + int(olen[i])].tobytes().decode() == v["trimmed"], v["name"]


def test_user_guide_examples():
    for q, kept in [("MYSEQUENCEADAPTER", "MYSEQUENCE"), ("MYSEQUENCEADAP", "MYSEQUENCE"),
                    ("MYSEQUENCEADAPTERSOMETHINGELSE", "MYSEQUENCE"), ("MYSEQUENCEAD", "MYSEQUENCEAD")]:
        t = oracle.locate("ADAPTER", q, 0.1, BACK, 3)
        assert (q[:t[2]] if t else q) == kept
    for q in ["ADAPTERMYSEQUENCE", "DAPTERMYSEQUENCE", "TERMYSEQUENCE", "SOMETHINGADAPTERMYSEQUENCE"]:
        t = oracle.locate("ADAPTER", q, 0.1, FRONT, 3)
        assert q[t[3]:] == "MYSEQUENCE"


def test_allowed_errors_by_length():
    # e = 0.1: 0 errors for aligned length 0-9, 1 for 10-19, ... (cutadapt guide, "Error tolerance")
    ad = "ACGTTGCAAGCTTAGGCATCGATCCGATTAGC"
    for L in (9, 10, 19, 20, 29, 30):
        read = "T" * 50 + ad[:L]
        mut = read[:52] + ("A" if read[52] != "A" else "C") + read[53:]    # one substitution inside
        t = oracle.locate(ad, mut, 0.1, BACK, 3)
        if L >= 10:
            assert t is not None and t[5] == 1 and t[1] == L
        else:
            assert t is None or t[5] == 0


def test_restatements_agree_random():
    rnd = random.Random(7)
    for _ in range(4000):
        m, n = rnd.randint(3, 30), rnd.randint(0, 70)
        ref = "".join(rnd.choice("ACGT") for _ in range(m))
        q = [rnd.choice("ACGT") for _ in range(n)]
        if n > 5 and rnd.random() < 0.75:
            p = rnd.randint(-m // 2, n - 1)
            for i, c in enumerate(ref):
                if 0 <= p + i < n and rnd.random() < 0.9:
                    q[p + i] = c
        q = "".join(q)
        # (0.5 and more: negative scores, where the last-column pass must not take what an unset best -- score 0 --
        # stands against, R6)
        rate = rnd.choice([0.0, 0.1, 0.2, 0.3, 0.34, 0.5, 0.7])
        flags = rnd.choice([FRONT, BACK, PREFIX, SUFFIX, 15, 10, 6, 9])
        mo, ic = rnd.randint(1, 5), rnd.choice([1, 1, 1, 100000])
        a = oracle.locate(ref, q, rate, flags, mo, ic)
        assert a == oracle.locate(ref, q, rate, flags, mo, ic, unpruned=True) == P.locate(ref, q, rate, flags, mo, ic)


def test_round_read_python_vs_c():
    rs = synth.generate(40, 300, 500, seed=3)
    rounds = H.m13_rounds()
    rec0, rec1, oseq, oqual, olen = H.run_oracle(rounds, rs, n_threads=1)
    ads = [P.Adapter(n, s, P.FRONT) for n, s in m13.sp5_forward()]
    for r in range(12):
        name, seq, qual = rs.read(r)
        idx, is_rc, t, nm, s, q = P.round_read(ads, True, name, seq, qual)
        assert idx == rec0["adapter"][r] and int(is_rc) == rec0["is_rc"][r]
        if t is not None:
            assert list(t) == [int(rec0[f][r]) for f in H.FIELDS[2:]]


def test_affix_comparers():
    assert oracle.affix_compare("AAXAA", "AAAAATTTTTTTTT", 0.9) == (0, 5, 0, 5, 3, 1)
    assert oracle.affix_compare("AANAA", "AACAATTTTTTTTT", 0.9, wildcard_ref=True) == (0, 5, 0, 5, 5, 0)
    assert oracle.affix_compare("AAXAA", "TTTTTTTAAAAA", 0.9, suffix=True) == (0, 5, 7, 12, 3, 1)
    assert oracle.affix_compare("GAGCGTCTAATCGTAAT", "GAGCGTCTTATCGTAATACGT", 0.1, min_overlap=17) == (0, 17, 0, 17, 15, 1)
    assert oracle.affix_compare("GAGCGTCTAATCGTAAT", "GAGCGTCTTTTCGTAATACGT", 0.1, min_overlap=17) is None


def test_synthetic_truth_agreement():
    rs = synth.generate(3000, 300, 900, seed=11)
    rec0, rec1, *_ = H.run_oracle(H.m13_rounds(), rs)
    t = rs.truth
    has5 = t["sp5"] > 0
    assert ((rec0["adapter"] + 1 == t["sp5"]) | ~has5).mean() > 0.97
    both = (rec0["adapter"] + 1 == t["sp5"]) & (t["sp27"] > 0) & has5
    assert (rec1["adapter"][both] + 1 == t["sp27"][both]).mean() > 0.95
    assert abs(rec0["is_rc"].mean() - 0.10) < 0.03


def test_indexed_anchored_lookup_equals_a_literal_dict():
    """SURVEY R11 read literally: a dict over every string within k mismatches of every anchored adapter (more
    matches win a key, the later adapter wins equal matches), looked up with read[:L]; an N in the affix goes to the
    comparer loop (R8 over PrefixComparer).  The C oracle decides the same from Hamming distances without building
    the dict -- here the dict is built."""
    import itertools
    rnd = random.Random(811)

    def environment(s, k):
        out = {s: 0}
        for d in range(1, k + 1):
            for pos in itertools.combinations(range(len(s)), d):
                for sub in itertools.product("ACGT", repeat=d):
                    if all(s[p] != c for p, c in zip(pos, sub)):
                        t = list(s)
                        for p, c in zip(pos, sub):
                            t[p] = c
                        out["".join(t)] = d
        return out

    n_hit = n_fallback = 0
    for trial in range(14):
        L = rnd.choice([6, 9, 12, 17])
        rate = rnd.choice([0.0, 0.1, 0.12, 0.2])
        k = int(rate * L)
        indexed = k <= 2                      # (with more errors cutadapt builds no index: the comparer loop decides)
        base = "".join(rnd.choice("ACGT") for _ in range(L))
        seqs = []
        for _ in range(rnd.randint(2, 10)):
            s = list(base) if rnd.random() < 0.6 else [rnd.choice("ACGT") for _ in range(L)]
            for _k in range(rnd.randint(0, 3)):
                s[rnd.randrange(L)] = rnd.choice("ACGT")
            seqs.append("".join(s))
        index = {}
        for a, s in enumerate(seqs if indexed else []):
            for key, e in environment(s, k).items():
                m = L - e
                if key in index and m < index[key][2]:
                    continue
                index[key] = (a, e, m)
        suffix = trial % 2 == 1
        recs = []
        for i in range(300):
            a = list(rnd.choice(seqs))
            for _k in range(rnd.choice([0, 0, 1, 1, 2, 3])):
                a[rnd.randrange(L)] = rnd.choice("ACGTN" if rnd.random() < 0.2 else "ACGT")
            body = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 30)))
            s = (body + "".join(a)) if suffix else ("".join(a) + body)
            if rnd.random() < 0.05:
                s = s[:rnd.randint(0, L - 1)]           # shorter than the adapters: no match
            recs.append(("r%d" % i, s, "I" * len(s)))
        rs = synth.from_records(recs)
        sets = [(oracle.AdapterSet(seqs, oracle.SUFFIX if suffix else oracle.PREFIX, rate, 3, indels=False), 0)]
        rec, *_ = oracle.demux_batch(sets, rs.seq, rs.qual, rs.offsets, rs.lengths, n_threads=1)
        for i, (_, s, _) in enumerate(recs):
            affix = (s[len(s) - L:] if suffix else s[:L]) if len(s) >= L else None
            if affix is None:
                want = None
            elif "N" in affix or not indexed:
                n_fallback += 1
                want = None
                for a, ad in enumerate(seqs):
                    matches = sum(x == y for x, y in zip(ad, affix))
                    e = L - matches
                    if e > int(rate * L):
                        continue
                    if want is None or matches - e > want[1] or (matches - e == want[1] and e < want[2]):
                        want = (a, matches - e, e)
            else:
                hit = index.get(affix)
                want = None if hit is None else (hit[0], hit[2], hit[1])
            got = None if rec["adapter"][i] < 0 else (int(rec["adapter"][i]), int(rec["score"][i]), int(rec["errors"][i]))
            assert got == want, (trial, seqs, rate, s, got, want)
            n_hit += want is not None
    assert n_hit > 1500 and n_fallback > 50
