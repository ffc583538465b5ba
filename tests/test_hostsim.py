"""The device arithmetic (csrc/orc_core.cuh: bit-parallel scan -> candidate hull -> banded
exact DP -> selection), compiled for the host by tests/hostsim.cpp, against the oracle.
This is how the exactness of the kernels' algorithm is checked without a GPU; the `-m gpu`
tests repeat the comparison through the real kernels and the C ABI."""
import random

import numpy as np
import pytest

import helpers as H
import oracle
from orcdemux import m13, synth


def _compare(rounds, rs, threads=8):
    rec0, rec1, oseq, oqual, olen = H.run_oracle(rounds, rs, n_threads=threads)
    m0, m1, lo, ln, rc, nt = H.run_hostsim(rounds, rs)
    idx, nbad = H.diff_matches(rec0, m0)
    assert nbad == 0, ("round 1", idx, [rs.read(int(i))[1] for i in idx[:1]])
    if len(rounds) > 1:
        idx, nbad = H.diff_matches(rec1, m1)
        assert nbad == 0, ("round 2", idx)
    assert np.array_equal(olen, ln)
    vb = H.view_bytes(rs, lo, ln, rc)
    for i in range(rs.n_reads):
        o, L = int(rs.offsets[i]), int(olen[i])
        assert vb[i][0] == oseq[o:o + L].tobytes() and vb[i][1] == oqual[o:o + L].tobytes(), i
    return rec0, rec1


def test_synthetic_coi():
    _compare(H.m13_rounds(), synth.generate(6000, 300, 900, seed=1002))


def test_synthetic_rrna():
    _compare(H.m13_rounds(), synth.generate(500, 1000, 3500, seed=1003))


def test_seed_segments_of_long_reads():
    """Reads longer than one seed segment (orc_core.cuh SEED_SEG = 1024 columns, at most four segments, the last
    one taking the rest): adapters -- intact, with errors, reverse-complemented -- at and across every segment
    boundary and the overlap behind it, also two of them close together on either side of a boundary, in reads
    of two to six segments.  The segmented probe must open the same windows as the whole-read probe did."""
    rnd = random.Random(4242)
    f = [s for _, s in m13.sp5_forward()]
    b = [s for _, s in m13.sp27_reverse_rc()]
    comp = str.maketrans("ACGT", "TGCA")

    def mutate(s, rate):
        out = []
        for c in s:
            u = rnd.random()
            if u < rate:
                out.append(rnd.choice("ACGT"))
            elif u < 1.5 * rate:
                continue
            elif u < 2 * rate:
                out.append(c); out.append(rnd.choice("ACGT"))
            else:
                out.append(c)
        return "".join(out)

    recs = []
    for i in range(700):
        n = rnd.choice([1100, 1500, 2047, 2048, 2049, 2500, 3071, 3200, 4096, 4200, 5000, 6200])
        body = [rnd.choice("ACGT") for _ in range(n)]
        for _ in range(rnd.randint(1, 3)):
            ad = rnd.choice(f + b)
            if rnd.random() < 0.5:
                ad = mutate(ad, rnd.choice([0.0, 0.03, 0.06, 0.1]))
            if rnd.random() < 0.4:
                ad = ad.translate(comp)[::-1]
            edge = rnd.choice([1024, 2048, 3072, 4096]) + rnd.choice([0, 0, 72, 73, -8, 8])
            pos = min(max(0, edge - rnd.randint(0, len(ad) + 12)), max(0, n - len(ad)))
            body[pos:pos + len(ad)] = list(ad)
        s = "".join(body)
        if rnd.random() < 0.5:
            s = rnd.choice(f) + s
        if rnd.random() < 0.5:
            s = s + rnd.choice(b)
        recs.append(("s%d" % i, s, "I" * len(s)))
    # the lemma's worst case: five substitutions leave exactly two of the seven 8-row pieces intact, and the
    # segment boundary falls between those two -- only the overlap behind a segment keeps the pair together
    for i in range(300):
        ad = list(rnd.choice(f))
        keep = sorted(rnd.sample(range(7), 2))
        for pc in range(7):
            if pc not in keep:
                q = 8 * pc + rnd.randrange(8)
                ad[q] = rnd.choice([c for c in "ACGT" if c != ad[q]])
        ad = "".join(ad)
        edge = rnd.choice([1024, 2048, 3072])
        start = edge - 8 * keep[0] - rnd.randint(1, 8 * (keep[1] - keep[0]))
        n = rnd.choice([edge + 200, edge + 1500, 4500])
        body = [rnd.choice("ACGT") for _ in range(n)]
        body[start:start + len(ad)] = list(ad)
        sq = "".join(body)
        if rnd.random() < 0.5:
            sq = sq.translate(comp)[::-1]
        recs.append(("w%d" % i, sq, "I" * len(sq)))
    rs = synth.from_records(recs)
    rec0, _ = _compare(H.m13_rounds(), rs)
    assert int((rec0["adapter"][700:] >= 0).sum()) >= 290       # the constructed occurrences are found
    # the flank scan instead of the seeds must agree as well (same windows, by a different route)
    m0a, m1a, *_ = H.run_hostsim(H.m13_rounds(), rs)
    m0b, m1b, *_ = H.run_hostsim(H.m13_rounds(), rs, filter_mode=1 | 4)
    assert H.diff_matches(m0a, m0b)[1] == 0 and H.diff_matches(m1a, m1b)[1] == 0


def test_no_rc_and_other_thresholds():
    rs = synth.generate(1500, 300, 600, seed=5)
    for e, ov, rc in [(0.1, 3, 0), (0.2, 5, 1), (0.0, 3, 1), (0.05, 10, 1), (0.15, 1, 1), (3, 3, 1)]:
        _compare(H.m13_rounds(e, ov, rc), rs)


def _adversarial_reads(rnd, adapters_f, adapters_b, n):
    recs = []
    comp = str.maketrans("ACGT", "TGCA")
    for i in range(n):
        kind = rnd.randrange(14)
        L = rnd.randint(0, 260)
        body = "".join(rnd.choice("ACGT") for _ in range(L))
        a = rnd.choice(adapters_f)
        b = rnd.choice(adapters_b)

        def mutate(s, rate):
            out = []
            for c in s:
                u = rnd.random()
                if u < rate:
                    out.append(rnd.choice("ACGT"))
                elif u < 1.5 * rate:
                    continue
                elif u < 2 * rate:
                    out.append(c); out.append(rnd.choice("ACGT"))
                else:
                    out.append(c)
            return "".join(out)
        if kind == 0:
            s = a + body + b
        elif kind == 1:
            s = mutate(a, 0.06) + body + mutate(b, 0.06)
        elif kind == 2:
            s = a[rnd.randint(0, len(a)):] + body + b[:rnd.randint(0, len(b))]
        elif kind == 3:
            s = body[:L // 2] + mutate(a, 0.04) + body[L // 2:] + a + body[:7] + mutate(b, 0.05) + b
        elif kind == 4:
            s = rnd.choice("ACGT") * L + a[-rnd.randint(1, 12):]
        elif kind == 5:
            s = (a + body + b).translate(comp)[::-1]
        elif kind == 6:
            s = mutate(a, 0.12) + body + mutate(b, 0.12)
        elif kind == 7:
            s = a[:rnd.randint(1, len(a))] * 2 + body + b[-rnd.randint(1, len(b)):] * 2
        elif kind == 8:
            s = body + b[:rnd.randint(0, 14)]
        elif kind == 9:
            s = a[-rnd.randint(0, 14):] + body
        elif kind == 10:
            s = mutate(a + b, 0.03) * rnd.randint(1, 3)
        elif kind == 12:      # 3' adapter in the middle of the read, near the error limit
            tail = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(40, 200)))
            s = body + mutate(b, rnd.choice([0.03, 0.08, 0.12])) + tail
        elif kind == 13:      # 5' adapter in the middle of the read
            tail = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(40, 200)))
            s = body + mutate(a, rnd.choice([0.03, 0.08, 0.12])) + tail
        else:
            s = "".join(rnd.choice("ACGTN") for _ in range(L))
        if rnd.random() < 0.1:
            s = s.lower()
        recs.append(("a%d" % i, s, "".join(chr(33 + rnd.randint(0, 40)) for _ in s)))
    return synth.from_records(recs)


def test_adversarial_m13():
    rnd = random.Random(99)
    f = [s for _, s in m13.sp5_forward()]
    b = [s for _, s in m13.sp27_reverse_rc()]
    _compare(H.m13_rounds(), _adversarial_reads(rnd, f, b, 4000))


def test_random_adapter_sets():
    rnd = random.Random(123)
    for trial in range(12):
        nf, nb = rnd.randint(1, 16), rnd.randint(1, 16)
        mk = lambda: "".join(rnd.choice("ACGT") for _ in range(rnd.choice([3, 8, 17, 20, 33, 57, 64])))
        shared = mk()[:10]
        f = [(shared if rnd.random() < 0.5 else "") + mk() for _ in range(nf)]
        f = [x[:64] for x in f]
        b = [(mk() + (shared if rnd.random() < 0.5 else ""))[:64] for _ in range(nb)]
        e = rnd.choice([0.0, 0.1, 0.1, 0.2, 0.3])
        ov = rnd.choice([1, 3, 3, 5, 8])
        rc = rnd.choice([0, 1, 1])
        rounds = [(f, oracle.FRONT, e, ov, rc), (b, oracle.BACK, e, ov, rc)]
        _compare(rounds, _adversarial_reads(rnd, f, b, 400), threads=4)


def _seed_friendly_sets(rnd):
    """Adapter sets stage 1 can seed: every adapter at least 8 (k + 1) long, a long shared prefix."""
    nf, nb = rnd.randint(1, 16), rnd.randint(1, 16)
    mk = lambda: "".join(rnd.choice("ACGT") for _ in range(rnd.choice([40, 48, 57, 59, 64])))
    shared = mk()[:rnd.choice([17, 25, 32])]
    sfx = mk()[:rnd.choice([0, 8, 17, 23])]
    f = [(shared + mk())[:64 - len(sfx)] + sfx for _ in range(nf)]
    b = [(shared + mk())[:64 - len(sfx)] + sfx for _ in range(nb)]
    e = rnd.choice([0.0, 0.03, 0.05, 0.08, 0.1, 0.1])       # floor(m / 8) - k is 1 or more, often exactly 1 or 2
    return f, b, e, rnd.choice([1, 3, 3, 8]), rnd.choice([0, 1, 1])


def _seeded_passes():
    import ctypes
    sd = (ctypes.c_uint64 * 2)()
    H.hostsim().hostsim_seeded(sd)
    return int(sd[0]) + int(sd[1])


def test_seeded_stage1_m13_and_random_sets():
    """Stage 1s (exact 8-mer seeds): the seeded path, the flank scan
    it replaces (filter_mode bit 2) and the oracle agree; the seeded path is really taken."""
    rnd = random.Random(4242)
    f = [s for _, s in m13.sp5_forward()]
    b = [s for _, s in m13.sp27_reverse_rc()]
    rs = _adversarial_reads(rnd, f, b, 3000)
    before = _seeded_passes()
    rec0, rec1 = _compare(H.m13_rounds(), rs)
    assert _seeded_passes() - before >= rs.n_reads          # round 1 of every read, round 2 of the assigned ones
    mid = _seeded_passes()
    m0, m1, lo, ln, rc, nt = H.run_hostsim(H.m13_rounds(), rs, filter_mode=1 | 4)
    assert _seeded_passes() == mid                          # bit 2: no seeds
    assert H.diff_matches(rec0, m0)[1] == 0 and H.diff_matches(rec1, m1)[1] == 0
    # tens of kilobases: more chance key matches than the seed scan keeps apart -> flank scan
    body = "".join(rnd.choice("ACGT") for _ in range(40000))
    long_recs = [("l0", f[3] + body + b[4]), ("l1", m13.revcomp(f[5] + body + b[1])), ("l2", body[:20000] + f[2] + body[20000:])]
    _compare(H.m13_rounds(), synth.from_records([(nm, sq, "I" * len(sq)) for nm, sq in long_recs]))
    seeded = 0
    for trial in range(10):
        f, b, e, ov, rc_ = _seed_friendly_sets(rnd)
        rounds = [(f, oracle.FRONT, e, ov, rc_), (b, oracle.BACK, e, ov, rc_)]
        if trial % 3 == 2:
            rounds = rounds[::-1]
        before = _seeded_passes()
        _compare(rounds, _adversarial_reads(rnd, f, b, 400), threads=4)
        seeded += _seeded_passes() - before
    assert seeded > 2000


def test_no_indels_unanchored():
    """--no-indels on regular adapters: Hamming distance along diagonals, settled in the scan."""
    rnd = random.Random(17)
    f = [s for _, s in m13.sp5_forward()]
    b = [s for _, s in m13.sp27_reverse_rc()]
    for rs in (synth.generate(2500, 300, 600, seed=9), _adversarial_reads(rnd, f, b, 2500)):
        rounds = H.m13_rounds()
        rec0, rec1, oseq, oqual, olen = H.run_oracle(rounds, rs, indels=False)
        for fm in (0, 1):
            m0, m1, lo, ln, rc, nt = H.run_hostsim(rounds, rs, fm, indels=0)
            assert H.diff_matches(rec0, m0)[1] == 0 and H.diff_matches(rec1, m1)[1] == 0
            assert np.array_equal(olen, ln) and int(nt.sum()) == 0


def test_single_round_back_only():
    rnd = random.Random(5)
    b = [s for _, s in m13.sp27_reverse_rc()]
    rs = _adversarial_reads(rnd, b, b, 600)
    _compare([(b, oracle.BACK, 0.1, 3, 1)], rs)


def _anchored_reads(rnd, seqs, n, suffix):
    recs = []
    comp = str.maketrans("ACGT", "TGCA")
    for i in range(n):
        a = list(rnd.choice(seqs))
        kind = rnd.randrange(9)
        if kind == 1:
            a[rnd.randrange(len(a))] = rnd.choice("ACGT")
        elif kind == 2:
            for _ in range(2):
                a[rnd.randrange(len(a))] = rnd.choice("ACGT")
        elif kind == 3:
            a[rnd.randrange(len(a))] = "N"
        elif kind == 4:
            a[rnd.randrange(len(a))] = rnd.choice("RYKMSW-")
        elif kind == 5:
            a = a[:rnd.randrange(len(a))]
        elif kind == 6:
            a[rnd.randrange(len(a))] = "N"; a[rnd.randrange(len(a))] = rnd.choice("ACGT")
        a = "".join(a)
        body = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(0, 60)))
        s = body + a if suffix else a + body
        if kind == 7:
            s = s.translate(comp)[::-1]
        if kind == 8:
            s = s.lower()
        recs.append(("h%d" % i, s, "I" * len(s)))
    return synth.from_records(recs)


def test_anchored_no_indel_path():
    """BASELINE config 4 shape: -g ^file: / -a file$: with --no-indels (indexed Hamming lookup)."""
    rnd = random.Random(4)
    var = [s for _, s in m13.variable_all()]
    close = ["ACGTACGTACGTACGTA", "ACGTACGTACGTACGTC", "ACGAACGTACGTACGTA", "TTGTACGTACGTACGTA"]   # ties / collisions
    for seqs in (var, close, var[:1], ["ACGTAC", "ACGTTC", "TCGTAC"]):
        for typ in (oracle.PREFIX, oracle.SUFFIX):
            for e, rc in ((0.1, 1), (0.12, 0), (0.34, 1), (0.0, 1)):
                rs = _anchored_reads(rnd, seqs, 500, typ == oracle.SUFFIX)
                sets = [(oracle.AdapterSet(seqs, typ, e, 3, indels=False), rc)]
                rec0, _, oseq, oqual, olen = oracle.demux_batch(sets, rs.seq, rs.qual, rs.offsets, rs.lengths, n_threads=2)
                m0, m1, lo, ln, rcv, nt = H.run_hostsim([(seqs, typ, e, 3, rc)], rs)
                idx, nbad = H.diff_matches(rec0, m0)
                assert nbad == 0, (seqs[:2], typ, e, rc, idx, rs.read(int(idx[0]))[1], rec0[idx[0]], m0[idx[0]])
                assert np.array_equal(olen, ln)
                vb = H.view_bytes(rs, lo, ln, rcv)
                for i in range(rs.n_reads):
                    o, L = int(rs.offsets[i]), int(olen[i])
                    assert vb[i][0] == oseq[o:o + L].tobytes()
    rs = synth.generate(3000, 300, 600, seed=1004, anchored=True)
    sets = [(oracle.AdapterSet(var, oracle.PREFIX, 0.1, 3, indels=False), 1)]
    rec0, *_ = oracle.demux_batch(sets, rs.seq, rs.qual, rs.offsets, rs.lengths, n_threads=4)
    m0, *_ = H.run_hostsim([(var, oracle.PREFIX, 0.1, 3, 1)], rs)
    assert H.diff_matches(rec0, m0)[1] == 0
    assert (rec0["adapter"] >= 0).mean() > 0.5


IUPAC = {"R": "AG", "Y": "CT", "S": "CG", "W": "AT", "K": "GT", "M": "AC", "B": "CGT", "D": "AGT",
         "H": "ACT", "V": "ACG", "N": "ACGT", "I": "ACGT", "X": "ACGT"}


def _iupac_sets(rnd, n_in_front):
    """Adapter sets in which every adapter holds IUPAC wildcards (cutadapt then compares through
    bit masks, takes N out of the effective length and reads U as T)."""
    def mk(allow_n):
        m = rnd.choice([8, 17, 20, 26, 33, 57, 64])
        s = [rnd.choice("ACGT") for _ in range(m)]
        for _ in range(rnd.randint(1, max(1, m // 5))):
            s[rnd.randrange(m)] = rnd.choice("RYSWKMBDHV" + ("NNNI" if allow_n else "") + ("X" if rnd.random() < 0.05 else ""))
        if rnd.random() < 0.3 and allow_n and m >= 33:        # a run of N like the M13 index placeholder
            a = rnd.randrange(m - 17)
            s[a:a + 17] = "N" * 17
        if rnd.random() < 0.3:
            s = [c.lower() for c in s]
        return "".join(s)
    shared = "".join(rnd.choice("ACGTRY") for _ in range(14))
    nf, nb = rnd.randint(1, 12), rnd.randint(1, 12)
    f = [((shared if rnd.random() < 0.6 else "") + mk(n_in_front))[:64] for _ in range(nf)]
    b = [(mk(True) + (shared if rnd.random() < 0.6 else ""))[:64] for _ in range(nb)]
    return f, b


def _instances(rnd, adapters, per=3):
    out = []
    for a in adapters:
        for _ in range(per):
            out.append("".join(rnd.choice(IUPAC[c]) if c in IUPAC else c for c in a.upper().replace("U", "T")))
    return out


def test_iupac_adapter_sets():
    """Next row N4: adapters with IUPAC wildcards (the primer shape of 04_cleaning_primers.sh)."""
    rnd = random.Random(404)
    done = refused = 0
    for trial in range(16):
        f, b = _iupac_sets(rnd, n_in_front=(trial % 2 == 1))
        e = rnd.choice([0.0, 0.1, 0.1, 0.15, 0.2, 2.0])
        ov = rnd.choice([1, 3, 3, 5, 8])
        rc = rnd.choice([0, 1, 1])
        rounds = [(f, oracle.FRONT, e, ov, rc), (b, oracle.BACK, e, ov, rc)]
        if trial % 5 == 4:
            rounds = rounds[1:]
        rs = _adversarial_reads(rnd, _instances(rnd, f), _instances(rnd, b), 300)
        recs = [rs.read(i) for i in range(rs.n_reads)]
        recs = [(nm, sq.replace("T", "U", 1) if i % 9 == 0 else sq, q) for i, (nm, sq, q) in enumerate(recs)]
        # reads that lie inside a 5' adapter: the only place where the last-column test sees a partial
        # overlap of a 5' adapter, with its own N count
        for a in _instances(rnd, f, per=1):
            for _ in range(6):
                cut = a[-rnd.randint(1, len(a)):]
                if rnd.random() < 0.5 and len(cut) > 2:
                    p = rnd.randrange(len(cut))
                    cut = cut[:p] + rnd.choice("ACGT") + cut[p + 1:]
                recs.append(("in%d" % len(recs), cut, "I" * len(cut)))
        rs = synth.from_records(recs)
        try:
            rec0, rec1 = _compare(rounds, rs, threads=4)
        except RuntimeError as ex:
            # an absolute error count that is too large for some adapter
            assert "unsupported" in str(ex) and "error rate" in str(ex), str(ex)
            refused += 1
            continue
        done += 1
        assert (rec0["adapter"] >= 0).mean() > 0.3
    assert done >= 10, (done, refused)
    # the M13 primers with their 17 N placeholder (adapters_primers/M13_seqs_for_pychopper.fa shape) as 3' adapters
    sp = ["CATGTAATGCACGTACTTTCAGGGTNNNNNNNNNNNNNNNNNTGTAAAACGACGGCCA", "GATCAGGTGAGGCTGCGACGACTNNNNNNNNNNNNNNNNNCAGGAAACAGCTATGAC"]
    rs = _adversarial_reads(rnd, _instances(rnd, sp), _instances(rnd, sp), 600)
    _compare([(sp, oracle.BACK, 0.1, 3, 1)], rs)


def test_plain_and_iupac_adapters_side_by_side():
    """cutadapt decides per adapter: a plain ACGT adapter is compared as ASCII, one with wildcards through the IUPAC
    masks.  Mixed in one round, or one kind per round, all of them run through the masks here (a plain adapter is
    its own mask set without N) -- the same results except for a read with U (T only through the masks), and a
    batch that holds one is refused."""
    rnd = random.Random(808)
    hits = 0
    for trial in range(10):
        f, b = _iupac_sets(rnd, n_in_front=(trial % 2 == 1))
        pf = ["".join(rnd.choice("ACGT") for _ in range(rnd.choice([8, 17, 25, 40, 64]))) for _ in range(rnd.randint(1, 5))]
        pb = ["".join(rnd.choice("ACGT") for _ in range(rnd.choice([8, 17, 25, 40, 64]))) for _ in range(rnd.randint(1, 5))]
        if trial % 3 == 0:          # mixed inside both rounds
            f, b = (f + pf)[:16], (pb + b)[:16]
            rnd.shuffle(f); rnd.shuffle(b)
        elif trial % 3 == 1:        # a plain round in front of an IUPAC one
            f = pf
        else:                       # an IUPAC round in front of a plain one
            b = pb
        e = rnd.choice([0.0, 0.1, 0.1, 0.2])
        ov = rnd.choice([1, 3, 5])
        rounds = [(f, oracle.FRONT, e, ov, 1), (b, oracle.BACK, e, ov, 1)]
        rs = _adversarial_reads(rnd, _instances(rnd, f), _instances(rnd, b), 500)
        rec0, rec1 = _compare(rounds, rs, threads=4)
        hits += int((rec0["adapter"] >= 0).sum()) + int((rec1["adapter"] >= 0).sum())
    assert hits > 2000
    # a read with U: refused when plain adapters stand beside IUPAC ones, fine when all are of one kind
    u = synth.from_records([("x", "ACGUACGUACGT", "I" * 12)])
    with pytest.raises(RuntimeError, match="unsupported.*U"):
        H.run_hostsim([(["ACGTACGT", "ACGNACGT"], oracle.BACK, 0.1, 3, 1)], u)
    with pytest.raises(RuntimeError, match="unsupported.*U"):
        H.run_hostsim([(["ACGTACGT"], oracle.FRONT, 0.1, 3, 1), (["ACGNACGT"], oracle.BACK, 0.1, 3, 1)], u)
    H.run_hostsim([(["ACGNACGT", "ACGRACGT"], oracle.BACK, 0.1, 3, 1)], u)
    H.run_hostsim([(["ACGTACGT"], oracle.BACK, 0.1, 3, 1)], u)


def test_iupac_degenerate_adapters():
    """Adapters made of N only (effective length 0: no errors allowed, everything matches), of X only
    (nothing matches), 64 N, wildcards at either end."""
    rnd = random.Random(1)
    recs = []
    for i in range(300):
        s = "".join(rnd.choice("ACGTN") for _ in range(rnd.randint(0, 60)))
        recs.append(("r%d" % i, s, "I" * len(s)))
    rs = synth.from_records(recs)
    for f, b in ((["NNNN"], ["NNNNNN"]), (["X"], ["XX"]), (["N" * 24, "ACGN"], ["NACGT", "N" * 64]),
                 (["XACGT", "ACGTX"], ["NX", "XN"])):
        for e, ov, rc in ((0.0, 1, 1), (0.1, 3, 0), (0.5, 1, 1)):
            rounds = [(f, oracle.FRONT, e, ov, rc), (b, oracle.BACK, e, ov, rc)]
            rec0, rec1, oseq, oqual, olen = H.run_oracle(rounds, rs, n_threads=4)
            for fm in (0, 2):
                m0, m1, lo, ln, rcv, nt = H.run_hostsim(rounds, rs, fm)
                assert H.diff_matches(rec0, m0)[1] == 0 and H.diff_matches(rec1, m1)[1] == 0, (f, b, e, ov, rc, fm)
                assert np.array_equal(olen, ln)


def test_unsupported_is_refused():
    rs = synth.from_records([("x", "ACGT", "IIII")])
    with pytest.raises(RuntimeError, match="unsupported"):
        H.run_hostsim([(["ACGZ"], oracle.FRONT, 0.1, 3, 1)], rs)
    with pytest.raises(RuntimeError, match="unsupported"):
        H.run_hostsim([(["A" * 257], oracle.FRONT, 0.1, 3, 1)], rs)


def test_up_to_32_adapters_and_config4_unanchored_arm():
    """More than 16 adapters per round (two banks of 32 lanes in the 64-bit match table), and BASELINE
    configs[3]'s second arm: `-g file:M13_variable_indices_all.fa` unanchored WITH indels (24 x 17-mers,
    k = 1, no shared flank -> no stage-1 filter) on the reads of the anchored arm."""
    rnd = random.Random(3232)
    var = [s for _, s in m13.variable_all()]
    assert len(var) == 24
    rs = synth.generate(3000, 300, 600, seed=1004, anchored=True)
    for rc in (1, 0):
        rec0, _ = _compare([(var, oracle.FRONT, 0.1, 3, rc)], rs)
        assert (rec0["adapter"] >= 16).sum() > 100 and (rec0["adapter"] >= 0).mean() > 0.5
    _compare([(var, oracle.BACK, 0.1, 3, 1)], _adversarial_reads(rnd, var, var, 1500))
    for trial in range(6):
        nf, nb = rnd.randint(17, 32), rnd.randint(17, 32)
        mk = lambda: "".join(rnd.choice("ACGT") for _ in range(rnd.choice([8, 17, 20, 33, 57, 64])))
        shared = mk()[:rnd.choice([4, 14, 25])]
        f = [((shared if rnd.random() < 0.8 else "") + mk())[:64] for _ in range(nf)]
        b = [((shared if rnd.random() < 0.8 else "") + mk())[:64] for _ in range(nb)]
        e = rnd.choice([0.0, 0.1, 0.1, 0.2])
        rounds = [(f, oracle.FRONT, e, 3, 1), (b, oracle.BACK, e, 3, 1)]
        if trial % 2:
            rounds = rounds[::-1]
        rec0, rec1 = _compare(rounds, _adversarial_reads(rnd, f, b, 400), threads=4)
        assert (rec0["adapter"] >= 16).sum() > 0
    one = synth.from_records([("x", "ACGT", "IIII")])
    with pytest.raises(RuntimeError, match="unsupported"):
        H.run_hostsim([(["ACGTACGT"] * 33, oracle.FRONT, 0.1, 3, 1)], one)


def _long_sets(rnd, wild=False):
    alphabet = "ACGT" if not wild else "ACGTACGTACGTRYKMSWBDHVN"
    mk = lambda lo, hi: "".join(rnd.choice(alphabet) for _ in range(rnd.randint(lo, hi)))
    nf, nb = rnd.randint(1, 5), rnd.randint(1, 5)
    # every round holds at least one adapter over 64 nt; short ones ride along
    f = [mk(65, 200)] + [mk(rnd.choice([12, 40, 65, 90]), 130) for _ in range(nf - 1)]
    b = [mk(65, 256)] + [mk(rnd.choice([12, 40, 65, 90]), 130) for _ in range(nb - 1)]
    rnd.shuffle(f)
    rnd.shuffle(b)
    return f, b


def test_adapters_over_64_nt():
    """SURVEY 8f N4 "adapters > 64 nt": a round that holds one runs cutadapt's recurrence cell by cell
    (orc_core.cuh long_locate / long_match, long_kernel on the device).  Against the oracle: random sets of 1-5
    adapters of up to 256 nt per round (plain, and with IUPAC wildcards), both rounds long or a long round beside
    a bit-parallel one, with and without indels, --rc on and off, error rates up to 0.3 and as absolute counts."""
    rnd = random.Random(6401)
    hits = 0
    for trial in range(10):
        wild = trial % 3 == 2
        f, b = _long_sets(rnd, wild)
        e = rnd.choice([0.0, 0.1, 0.1, 0.2, 0.3, 5])
        ov = rnd.choice([1, 3, 3, 10, 70])
        rc = rnd.choice([0, 1, 1])
        indels = trial % 4 != 3
        plain = lambda x: "".join(c if c in "ACGT" else rnd.choice("ACGT") for c in x)
        if trial % 5 == 1:          # a long 5' round in front of the M13 3' round
            b = [s for _, s in m13.sp27_reverse_rc()] if not wild else ["ACGTNRYACGTTGCAAC", "TTGACNNRGATTACAGG"]
        if trial % 5 == 4:          # the M13 5' round in front of a long 3' round
            f = [s for _, s in m13.sp5_forward()] if not wild else ["ACGTNRYACGTTGCAAC", "TTGACNNRGATTACAGG"]
        rs = _adversarial_reads(rnd, [plain(x) for x in f], [plain(x) for x in b], 250)
        rounds = [(f, oracle.FRONT, e, ov, rc), (b, oracle.BACK, e, ov, rc)]
        rec0, rec1, oseq, oqual, olen = H.run_oracle(rounds, rs, n_threads=4, indels=indels)
        m0, m1, lo, ln, rcs, nt = H.run_hostsim(rounds, rs, indels=int(indels))
        assert H.diff_matches(rec0, m0)[1] == 0, ("round 1", trial)
        assert H.diff_matches(rec1, m1)[1] == 0, ("round 2", trial)
        assert np.array_equal(olen, ln)
        hits += int((rec0["adapter"] >= 0).sum()) + int((rec1["adapter"] >= 0).sum())
    assert hits > 1500
    # limits: 257 nt, 17 adapters in a long round
    with pytest.raises(RuntimeError, match="unsupported"):
        H.run_hostsim([(["A" * 257], oracle.FRONT, 0.1, 3, 1)], synth.from_records([("r", "ACGT", "IIII")]))
    with pytest.raises(RuntimeError, match="unsupported"):
        H.run_hostsim([(["ACGT" * 20] * 17, oracle.FRONT, 0.1, 3, 1)], synth.from_records([("r", "ACGT", "IIII")]))
