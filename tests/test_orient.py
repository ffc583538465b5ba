"""`orcdemux orient`, the pychopper-style step in front of the demultiplexer
(/root/reference/scripts/01_pychopper.sh:45-57): host logic on the CPU, the whole pass on the GPU against the
same decisions made from the oracle's matches."""
import os
import random
import subprocess
import sys

import numpy as np
import pytest

import helpers as H
import oracle
from orcdemux import orient, synth

SP5 = "CATGTAATGCACGTACTTTCAGGGTNNNNNNNNNNNNNNNNNTGTAAAACGACGGCCA"
SP27 = "GATCAGGTGAGGCTGCGACGACTNNNNNNNNNNNNNNNNNCAGGAAACAGCTATGAC"
CONFIG = "+:SP5,-SP27|-:SP27,-SP5"


def test_config_and_primer_pair():
    cfg = orient.parse_config(CONFIG + "\n")
    assert cfg == [("+", ("SP5", False), ("SP27", True)), ("-", ("SP27", False), ("SP5", True))]
    five, three = orient.primer_pair({"SP5": SP5, "SP27": SP27}, cfg)
    assert five == SP5 and three == orient.revcomp(SP27) and three.startswith("GTCATAGCTGTTTCCTG")
    with pytest.raises(orient.Unsupported):
        orient.parse_config("+SP5,SP27")
    with pytest.raises(orient.Unsupported):       # the - configuration must mirror the + configuration
        orient.primer_pair({"SP5": SP5, "SP27": SP27}, orient.parse_config("+:SP5,-SP27|-:SP5,-SP27"))
    with pytest.raises(orient.Unsupported):
        orient.primer_pair({"SP5": SP5}, cfg)


def test_mean_qscores():
    recs = [("a", "ACGT", "IIII"), ("b", "AC", "+5"), ("c", "", ""), ("d", "ACGTAC", "!!!!!I")]
    text = np.frombuffer("".join("@%s\n%s\n+\n%s\n" % r for r in recs).encode(), dtype=np.uint8)
    qoff, ln, pos = [], [], 0
    for nm, s, q in recs:
        pos += 1 + len(nm) + 1 + len(s) + 1 + 2
        qoff.append(pos)
        ln.append(len(q))
        pos += len(q) + 1
    got = orient.mean_qscores(text, np.array(qoff, np.uint64), np.array(ln, np.uint32), block=2)
    exp = []
    for _, _, q in recs:
        exp.append(0.0 if not q else -10 * np.log10(np.mean([10 ** (-(ord(c) - 33) / 10) for c in q])))
    assert np.allclose(got, exp)


def _reads(n, seed):
    rnd = random.Random(seed)
    fill = lambda p: "".join(rnd.choice("ACGT") if c == "N" else c for c in p)
    recs, truth = [], []
    for i in range(n):
        body = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(200, 700)))
        kind = rnd.random()
        five, three = fill(SP5), fill(orient.revcomp(SP27))
        if kind < 0.08:
            s, t = body, "none"
        elif kind < 0.16:
            s, t = five + body, "five_only"
        elif kind < 0.22:
            s, t = body + three, "three_only"
        else:
            s, t = five + body + three, "+"
        # sequencing errors
        out = []
        for ch in s:
            u = rnd.random()
            if u < 0.01:
                out.append(rnd.choice("ACGT"))
            elif u < 0.02:
                continue
            elif u < 0.03:
                out.append(ch); out.append(rnd.choice("ACGT"))
            else:
                out.append(ch)
        s = "".join(out)
        if rnd.random() < 0.4:
            s = orient.revcomp(s)
            t = "-" if t == "+" else t
        lowq = rnd.random() < 0.1
        q = "".join(chr(33 + (rnd.randint(2, 6) if lowq else rnd.randint(12, 40))) for _ in s)
        recs.append(("q%d" % i, s, q))
        truth.append(t)
    return recs, truth


@pytest.mark.gpu
@pytest.mark.parametrize("keep", [True, False])
def test_orient_pass_equals_oracle_composition(tmp_path, keep):
    """Reads with the pychopper primers (random 17-mers where the primers have N) in both orientations, with
    and without both primers: the pass file, its orientation and its cut points equal what the same two
    searches by the CPU oracle give, composed in Python."""
    recs, truth = _reads(3000, 5 + keep)
    inp = tmp_path / "in.fastq"
    inp.write_bytes("".join("@%s\n%s\n+\n%s\n" % r for r in recs).encode())
    prim = tmp_path / "primers.fa"
    prim.write_text(">SP5\n%s\n>SP27\n%s\n" % (SP5, SP27))
    cfgf = tmp_path / "config.txt"
    cfgf.write_text(CONFIG + "\n")
    shim = os.path.join(H.PKG, "bin", "pychopper")
    out = tmp_path / "pass.fastq"
    args = [shim, "-b", str(prim), "-c", str(cfgf), "-k", "LSK114", "-Q", "10", "-w", str(tmp_path / "resc.fastq"),
            "-u", str(tmp_path / "unc.fastq"), "-l", str(tmp_path / "short.fastq"), "-S", str(tmp_path / "stats.out"),
            "-t", "4", "-m", "edlib", str(inp)] + (["-p"] if keep else [])
    with open(out, "wb") as fh:
        r = subprocess.run(args, stdout=fh, stderr=subprocess.PIPE, text=False)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    # ---- expected, from the oracle
    five, three = SP5, orient.revcomp(SP27)
    mq = [-10 * np.log10(np.mean([10 ** (-(ord(c) - 33) / 10) for c in q])) for _, _, q in recs]
    kept = [r_ for r_, m in zip(recs, mq) if m >= 10]
    rs = synth.from_records(kept)
    s1 = [(oracle.AdapterSet([five], oracle.FRONT, 0.15, 10), 1)]
    rec0, *_ = oracle.demux_batch(s1, rs.seq, rs.qual, rs.offsets, rs.lengths, n_threads=4)
    stage2, who = [], []
    for i, (nm, s, q) in enumerate(kept):
        m = rec0[i]
        if m["adapter"] < 0:
            continue
        if m["is_rc"]:
            s, q, nm = orient.revcomp(s), q[::-1], nm + " rc"
        cut = int(m["query_start"]) if keep else int(m["query_stop"])
        stage2.append((nm, s[cut:], q[cut:]))
        who.append(i)
    rs2 = synth.from_records(stage2)
    s2 = [(oracle.AdapterSet([three], oracle.BACK, 0.15, 10), 1)]
    rec1, *_ = oracle.demux_batch(s2, rs2.seq, rs2.qual, rs2.offsets, rs2.lengths, n_threads=4)
    exp = []
    for j, (nm, s, q) in enumerate(stage2):
        m = rec1[j]
        if m["adapter"] < 0 or m["is_rc"]:
            continue
        cut = int(m["query_stop"]) if keep else int(m["query_start"])
        if cut >= 50:
            exp.append("@%s\n%s\n+\n%s\n" % (nm, s[:cut], q[:cut]))
    assert out.read_bytes() == "".join(exp).encode()
    assert len(exp) > 1200
    stats = dict(l.split("\t")[1:] for l in open(tmp_path / "stats.out").read().splitlines()[1:])
    assert int(stats["total"]) == len(recs) and int(stats["lowq"]) == len(recs) - len(kept)
    assert int(stats["passed"]) == len(exp) == int(stats["plus"]) + int(stats["minus"]) and int(stats["minus"]) > 200
    assert int(stats["passed"]) + int(stats["unclassified"]) + int(stats["short"]) == len(kept)
    assert (tmp_path / "resc.fastq").read_bytes() == b""
    if keep:        # the oriented reads start with the 5' primer and end with the 3' primer, ready for step 02
        starts = sum(e.split("\n")[1].startswith("CATGTAATGC") for e in exp)
        ends = sum(e.split("\n")[1].endswith("CCTGATC") for e in exp)
        assert starts > 0.7 * len(exp) and ends > 0.7 * len(exp)
