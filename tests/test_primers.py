"""Primer-trimming mode of the command line (one output file; 04_cleaning_primers.sh call shapes):
host logic on the CPU, the whole path on the GPU against the oracle's per-adapter matches."""
import random

import numpy as np
import pytest

import helpers  # noqa: F401  (puts the repo and the package on sys.path)
import oracle
from orcdemux import cli, primers
from orcdemux.lib import MATCH_DTYPE

FWD, REV = "GGWACWGGWTGAACWGTWTAYCCYCC", "TGRTTYTTYGGNCAYCCNGARGTNTA"      # degenerate COI primers (REV already rc)
FWD2, REV2 = "ACGTRYACGTSWACGTKMACGT", "TTGACCBDHVTTGACCAAGG"


def _expected(recs, pairs, e=0.1, ov=3):
    """LinkedAdapter / MultipleAdapters semantics straight from the oracle's per-adapter matches."""
    sets = [(oracle.AdapterSet([f], oracle.FRONT, e, ov), oracle.AdapterSet([r], oracle.BACK, e, ov)) for _, f, r in pairs]
    out = []
    for name, seq, qual in recs:
        best = None
        for i, (fs, bs) in enumerate(sets):
            m0 = fs.match(0, seq.upper())
            if m0 is None:
                continue
            rest = seq[m0[3]:]
            m1 = bs.match(0, rest.upper())
            if m1 is None:
                continue
            score, err = m0[4] + m1[4], m0[5] + m1[5]
            if best is None or score > best[0] or (score == best[0] and err < best[1]):
                best = (score, err, m0[3], m0[3] + m1[2])
        out.append(None if best is None else (name, seq[best[2]:best[3]], qual[best[2]:best[3]] if qual else None))
    return out


def _consensus(rnd, n, pairs):
    inst = lambda s: "".join(rnd.choice(primers_iupac[c]) if c in primers_iupac else c for c in s)
    recs = []
    for i in range(n):
        _, f, r = rnd.choice(pairs)
        body = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(20, 400)))
        kind = rnd.randrange(8)
        f1, r1 = inst(f), inst(r)
        if kind == 1:
            f1 = f1[:5] + rnd.choice("ACGT") + f1[6:]
        if kind == 2:
            r1 = r1[:7] + r1[8:]
        s = {3: body + r1, 4: f1 + body, 5: body}.get(kind, "ACGTTGCA"[:rnd.randint(0, 8)] + f1 + body + r1 + "TTGACA"[:rnd.randint(0, 6)])
        if kind == 6:
            s = s.lower()
        if kind == 7:
            _, f2, r2 = pairs[-1]
            s = inst(f2)[3:] + body + inst(r2)
        recs.append(("consensus_%d size=%d" % (i, rnd.randint(1, 99)), s, None))
    return recs


primers_iupac = {"R": "AG", "Y": "CT", "S": "CG", "W": "AT", "K": "GT", "M": "AC", "B": "CGT", "D": "AGT",
                 "H": "ACT", "V": "ACG", "N": "ACGT"}


def test_parse_and_argv():
    assert primers.parse_linked_specs(["ACGT...TTGA", "p2=AAC...GGT"]) == [("1", "ACGT", "TTGA"), ("p2", "AAC", "GGT")]
    assert primers.parse_linked_specs(["ACGT"]) == []
    with pytest.raises(primers.Unsupported):
        primers.parse_linked_specs(["ACGT...TTGA", "ACGT"])
    with pytest.raises(primers.Unsupported):
        primers.parse_linked_specs(["^ACGT...TTGA"])
    opt = cli.parse_cutadapt_argv(["-j", "4", "-g", "ACGT...TTGA", "--untrimmed-output=u.fasta", "-o", "t.fasta", "in.fasta"])
    assert opt["untrimmed_output"] == "u.fasta" and opt["out"] == "t.fasta" and opt["g"] == ["ACGT...TTGA"]
    with pytest.raises(cli.Unsupported):
        cli.parse_cutadapt_argv(["-g", "ACGT...TTGA", "-o", "{name}.fastq", "in.fastq"])
    with pytest.raises(cli.Unsupported):
        cli.parse_cutadapt_argv(["-g", "ACGT", "--untrimmed-output", "u.fq", "-o", "{name}.fastq", "in.fastq"])


def test_files_roundtrip(tmp_path):
    fa = tmp_path / "c.fasta"
    fa.write_text(">a desc\nACGT\nAC\n>b\n\n>c\nGG\n")
    recs, fmt = primers.read_sequences(str(fa))
    assert fmt == "fasta" and recs == [("a desc", "ACGTAC", None), ("b", "", None), ("c", "GG", None)]
    out = tmp_path / "o.fa.gz"
    primers.write_sequences(str(out), recs, primers.output_format(str(out), fmt))
    assert primers.read_sequences(str(out))[0] == recs
    fq = tmp_path / "r.fastq"
    fq.write_text("@r1 x\nACGT\n+\nIIII\n@r2\n\n+\n\n")
    recs, fmt = primers.read_sequences(str(fq))
    assert fmt == "fastq" and recs == [("r1 x", "ACGT", "IIII"), ("r2", "", "")]
    assert primers.output_format("x.out", "fastq") == "fastq" and primers.output_format("x.fna.gz", "fastq") == "fasta"
    with pytest.raises(primers.Unsupported):
        primers.write_sequences(str(tmp_path / "q.fastq"), [("a", "ACGT", None)], "fastq")


def test_select_and_trim():
    def rec(rows):
        m = np.zeros(len(rows), dtype=MATCH_DTYPE)
        for i, (ad, qs, qe, sc, er) in enumerate(rows):
            m[i]["adapter"], m[i]["query_start"], m[i]["query_stop"], m[i]["score"], m[i]["errors"] = ad, qs, qe, sc, er
        return m
    # read 0: only pair 0 complete; 1: pair 1 scores higher; 2: equal score, pair 1 fewer errors;
    # 3: full tie -> first; 4: nothing complete
    m0a = rec([(0, 0, 4, 4, 0), (0, 0, 4, 4, 0), (0, 0, 4, 2, 1), (0, 0, 4, 4, 0), (0, 0, 4, 4, 0)])
    m1a = rec([(0, 6, 9, 3, 0), (0, 6, 9, 3, 0), (0, 6, 9, 3, 0), (0, 6, 9, 3, 0), (-1, 0, 0, 0, 0)])
    m0b = rec([(0, 0, 5, 5, 0), (0, 1, 5, 4, 0), (0, 0, 5, 3, 0), (0, 0, 5, 4, 0), (-1, 0, 0, 0, 0)])
    m1b = rec([(-1, 0, 0, 0, 0), (0, 3, 9, 6, 0), (0, 2, 9, 2, 0), (0, 1, 9, 3, 0), (0, 2, 9, 5, 0)])
    best = primers.select_linked([m0a, m0b], [m1a, m1b])
    assert best.tolist() == [0, 1, 1, 0, -1]
    recs = [("r%d" % i, "ABCDEFGHIJKLMNOP", "0123456789abcdef") for i in range(5)]
    out, trimmed = primers.trim_linked(recs, [m0a, m0b], [m1a, m1b], best)
    assert trimmed.tolist() == [True, True, True, True, False]
    assert out[0][1:] == ("EFGHIJ", "456789") and out[1][1:] == ("FGH", "567") and out[2][1] == "FG" and out[4][1] == "ABCDEFGHIJKLMNOP"
    # plain adapters in groups: group 0 holds command-line positions 0 and 2, group 1 position 1
    ga = rec([(0, 0, 4, 4, 0), (1, 0, 4, 4, 0), (0, 0, 4, 4, 1), (-1, 0, 0, 0, 0), (-1, 0, 0, 0, 0)])
    gb = rec([(0, 6, 9, 4, 0), (0, 6, 9, 4, 0), (0, 6, 9, 4, 0), (0, 6, 9, 1, 0), (-1, 0, 0, 0, 0)])
    best, bpos = primers.select_best([ga, gb], [np.array([0, 2]), np.array([1])])
    assert best.tolist() == [0, 1, 1, 1, -1] and bpos.tolist() == [0, 1, 1, 1, -1]
    out, trimmed = primers.trim_single(recs, m0b, True)
    assert out[0][1] == "FGHIJKLMNOP" and out[4][1] == "ABCDEFGHIJKLMNOP" and not trimmed[4]
    out, _ = primers.trim_single(recs, m1b, False)
    assert out[1][1] == "ABC" and out[0][1] == "ABCDEFGHIJKLMNOP"


def _oracle_match_batches(rounds, recs, device, batch=1 << 16):
    """Stand-in for primers._match_batches in the CPU tests: the same per-round match records, from
    the oracle instead of the GPU (round 2 sees what round 1 left, like the engine)."""
    outs = [np.zeros(len(recs), dtype=MATCH_DTYPE) for _ in rounds]
    for o in outs:
        o["adapter"] = -1
    for r, (_, seq, _q) in enumerate(recs):
        rest = seq
        for k, rd in enumerate(rounds):
            front = rd.type == 0
            st = oracle.AdapterSet(rd.sequences, oracle.FRONT if front else oracle.BACK, rd.max_error_rate, rd.min_overlap)
            m = st.best_of(rest.upper())
            if m is None:
                break
            a, t = m
            o = outs[k][r]
            o["adapter"], o["ref_start"], o["ref_stop"], o["query_start"], o["query_stop"], o["score"], o["errors"] = (a,) + t
            rest = rest[t[3]:] if front else rest[:t[2]]
    return outs


def test_run_routing_on_cpu(tmp_path, monkeypatch):
    """primers.run end to end with the oracle standing in for the GPU passes: output routing
    (--untrimmed-output, --discard-untrimmed, everything in one file), groups by adapter type and
    wildcard class, more than 16 adapters of one kind."""
    monkeypatch.setattr(primers, "_match_batches", _oracle_match_batches)
    rnd = random.Random(5)
    pairs = [("1", FWD, REV), ("2", FWD2, REV2)]
    recs = _consensus(rnd, 120, pairs)
    src = tmp_path / "in.fasta"
    primers.write_sequences(str(src), recs, "fasta")
    exp = _expected(recs, pairs)
    base = ["-g", "%s...%s" % (FWD, REV), "-g", "%s...%s" % (FWD2, REV2)]
    t, u = tmp_path / "t.fasta", tmp_path / "u.fasta"
    c = primers.run(cli.parse_cutadapt_argv(base + ["--untrimmed-output", str(u), "-o", str(t), str(src)]))
    assert primers.read_sequences(str(t))[0] == [x for x in exp if x is not None]
    assert primers.read_sequences(str(u))[0] == [r for r, x in zip(recs, exp) if x is None]
    assert c["n_in"] == len(recs) and c["n_with"] == c["n_written"] == sum(x is not None for x in exp)
    assert sum(c["per_adapter"].values()) == c["n_with"]
    c = primers.run(cli.parse_cutadapt_argv(base + ["--discard-untrimmed", "-o", str(t), str(src)]))
    assert primers.read_sequences(str(t))[0] == [x for x in exp if x is not None]
    c = primers.run(cli.parse_cutadapt_argv(base + ["-o", str(t), str(src)]))
    assert primers.read_sequences(str(t))[0] == [x if x is not None else r for r, x in zip(recs, exp)]
    assert c["n_written"] == len(recs)
    # plain adapters of both kinds and wildcard classes, 18 of one kind: best over all in command-line order
    decoys = ["".join(rnd.choice("ACGT") for _ in range(18)) for _ in range(17)]
    ads = [(FWD, oracle.FRONT), (REV, oracle.BACK)] + [(d, oracle.FRONT) for d in decoys] + [(FWD2, oracle.FRONT), (REV2, oracle.BACK)]
    argv = [x for s_, w in ads for x in ("-g" if w == oracle.FRONT else "-a", s_)] + ["-o", str(t), str(src)]
    primers.run(cli.parse_cutadapt_argv(argv))
    sets = [oracle.AdapterSet([s_], w, 0.1, 3) for s_, w in ads]
    exp2 = []
    for name, seq, _ in recs:
        best = None
        for (s_, w), st in zip(ads, sets):
            m = st.match(0, seq.upper())
            if m is not None and (best is None or m[4] > best[0][4] or (m[4] == best[0][4] and m[5] < best[0][5])):
                best = (m, w)
        exp2.append((name, seq if best is None else (seq[best[0][3]:] if best[1] == oracle.FRONT else seq[:best[0][2]]), None))
    assert primers.read_sequences(str(t))[0] == exp2
    # FASTQ in, FASTQ out, qualities trimmed with the bases
    fq = tmp_path / "in.fastq"
    primers.write_sequences(str(fq), [(n, s_, "".join(chr(33 + (i % 40)) for i in range(len(s_)))) for n, s_, _ in recs[:20]], "fastq")
    primers.run(cli.parse_cutadapt_argv(["-g", FWD, "-o", str(tmp_path / "o.fastq"), str(fq)]))
    got = primers.read_sequences(str(tmp_path / "o.fastq"))[0]
    src_fq = primers.read_sequences(str(fq))[0]
    for (n, s_, q), (n0, s0, q0) in zip(got, src_fq):
        assert n == n0 and len(s_) == len(q) and s0.endswith(s_) and q0.endswith(q)
    with pytest.raises(primers.Unsupported):
        primers.run(cli.parse_cutadapt_argv(["--rc", "-g", FWD, "-o", str(t), str(src)]))
    # empty input: empty outputs, no GPU pass
    empty = tmp_path / "empty.fasta"
    empty.write_text("")
    c = primers.run(cli.parse_cutadapt_argv(base + ["--untrimmed-output", str(u), "-o", str(t), str(empty)]))
    assert c["n_in"] == 0 and t.read_text() == "" and u.read_text() == ""


@pytest.mark.gpu
def test_linked_primers_cli_gpu(tmp_path, capsys):
    """04_cleaning_primers.sh round 1 and round 2 through the command line on the GPU."""
    rnd = random.Random(77)
    pairs = [("1", FWD, REV), ("2", FWD2, REV2)]
    recs = _consensus(rnd, 700, pairs)
    src = tmp_path / "consensus.fasta"
    primers.write_sequences(str(src), recs, "fasta")
    trimmed_p, untrimmed_p = tmp_path / "primerless.fasta", tmp_path / "untrimmed.fasta"
    rc = cli.main(["-j", "8", "-g", "%s...%s" % (FWD, REV), "-g", "%s...%s" % (FWD2, REV2),
                   "--untrimmed-output=%s" % untrimmed_p, "-o", str(trimmed_p), str(src)])
    assert rc == 0
    exp = _expected(recs, pairs)
    got_t = primers.read_sequences(str(trimmed_p))[0]
    got_u = primers.read_sequences(str(untrimmed_p))[0]
    assert got_t == [x for x in exp if x is not None]
    assert got_u == [r for r, x in zip(recs, exp) if x is None]
    assert 100 < len(got_t) < 650 and len(got_u) > 50
    # round 2: the untrimmed ones against the forward primers alone, one output
    out2 = tmp_path / "round2.fasta"
    assert cli.main(["-j", "8", "-g", FWD, "-g", FWD2, "-o", str(out2), str(untrimmed_p)]) == 0
    fs = oracle.AdapterSet([FWD, FWD2], oracle.FRONT, 0.1, 3)
    exp2 = []
    for name, seq, _ in got_u:
        m = fs.best_of(seq.upper())
        exp2.append((name, seq[m[1][3]:] if m else seq, None))
    assert primers.read_sequences(str(out2))[0] == exp2
    # round 2 as the script writes it: forward primers with -g, reverse ones with -a, interleaved, one plain
    # adapter among the IUPAC ones; one match per read, the best over all of them in command-line order
    out3 = tmp_path / "round2_mixed.fasta"
    plain = "ACGTTGCAACGT"
    argv = ["-j", "8", "-g", FWD, "-a", REV, "-g", "p=" + plain, "-g", FWD2, "-a", REV2, "-o", str(out3), str(src)]
    assert cli.main(argv) == 0
    ads = [(FWD, oracle.FRONT), (REV, oracle.BACK), (plain, oracle.FRONT), (FWD2, oracle.FRONT), (REV2, oracle.BACK)]
    sets = [oracle.AdapterSet([s_], w, 0.1, 3) for s_, w in ads]
    exp3 = []
    for name, seq, _ in recs:
        best = None
        for (s_, w), st in zip(ads, sets):
            m = st.match(0, seq.upper())
            if m is not None and (best is None or m[4] > best[0][4] or (m[4] == best[0][4] and m[5] < best[0][5])):
                best = (m, w)
        if best is None:
            exp3.append((name, seq, None))
        else:
            exp3.append((name, seq[best[0][3]:] if best[1] == oracle.FRONT else seq[:best[0][2]], None))
    assert primers.read_sequences(str(out3))[0] == exp3
    assert sum(1 for a, b in zip(exp3, recs) if a[1] != b[1]) > 300
    # more adapters than one GPU pass takes: 20 plain 5' adapters, the real one last
    out4 = tmp_path / "many.fasta"
    decoys = ["".join(rnd.choice("ACGT") for _ in range(20)) for _ in range(19)]
    argv = ["-j", "8"] + [x for d in decoys for x in ("-g", d)] + ["-g", plain, "-o", str(out4), str(src)]
    assert cli.main(argv) == 0
    fs = oracle.AdapterSet(decoys + [plain], oracle.FRONT, 0.1, 3)
    exp4 = []
    for name, seq, _ in recs:
        m = fs.best_of(seq.upper())
        exp4.append((name, seq[m[1][3]:] if m else seq, None))
    assert primers.read_sequences(str(out4))[0] == exp4
    # refused shapes exit 2
    assert cli.main(["--rc", "-g", "%s...%s" % (FWD, REV), "-o", str(out2), str(src)]) == 2
    assert cli.main(["-g", "^" + FWD, "-o", str(out2), str(src)]) == 2
