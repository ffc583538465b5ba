"""The drop-in boundary on the GPU: the cutadapt-compatible command line replays
/root/reference/scripts/02_cutadapt_loop.sh:64-118 (round 1, the round-2 loop over the SP5
bins, the clean-up of `unknown` and SP27_009..012) on FASTQ(.gz) files, and the file tree and
decompressed bytes must equal what the oracle predicts.  The fused `two-round` command must
leave the same demuxed/SP27 tree."""
import glob
import gzip
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H
from orcdemux import cli, m13, synth

pytestmark = pytest.mark.gpu


def _expected_tree(rs, ds):
    """{relative path: bytes} after the reference script, from the oracle."""
    rec0, rec1, oseq, oqual, olen = H.run_oracle(H.m13_rounds(), rs)
    n5 = [n for n, _ in m13.sp5_forward()]
    n27 = [n for n, _ in m13.sp27_reverse_rc()]
    sp27 = {}
    for a in n5:
        for b in n27[:8]:
            sp27["SP27/%s_%s_%s.fastq.gz" % (b, a, ds)] = []
    for r in range(rs.n_reads):
        a0, a1 = int(rec0["adapter"][r]), int(rec1["adapter"][r])
        if a0 < 0 or a1 < 0 or a1 >= 8:
            continue
        name = rs.read(r)[0] + (" rc" if rec0["is_rc"][r] else "") + (" rc" if rec1["is_rc"][r] else "")
        o, L = int(rs.offsets[r]), int(olen[r])
        sp27["SP27/%s_%s_%s.fastq.gz" % (n27[a1], n5[a0], ds)].append(
            b"@" + name.encode() + b"\n" + oseq[o:o + L].tobytes() + b"\n+\n" + oqual[o:o + L].tobytes() + b"\n")
    return {k: b"".join(v) for k, v in sp27.items()}, rec0


def _read_gz(path):
    with gzip.open(path, "rb") as fh:
        return fh.read()


def test_reference_script_flow_and_fused(tmp_path):
    rs = synth.generate(6000, 300, 900, seed=77)
    ds = "sample1"
    work = tmp_path
    (work / "pychopped").mkdir()
    infile = work / "pychopped" / ("pychopped_%s_pass.fastq.gz" % ds)
    with gzip.open(infile, "wb", compresslevel=1) as fh:
        fh.write(rs.to_fastq_bytes())
    assert cli.dataset_name(str(infile)) == ds
    fwd, rev, _ = m13.write_tables(str(work / "adapters"))
    out = work / "demuxed"
    (out / "SP5").mkdir(parents=True)
    (out / "SP27").mkdir(parents=True)
    shim = os.path.join(H.PKG, "bin", "cutadapt")

    # round 1 (02:64-72), options after the positional input like the reference
    r = subprocess.run([shim, "--action=trim", "-e", "0.1", "-j", "24", "--rc", "-g", "file:" + fwd,
                        "-o", str(out / "SP5" / ("{name}_%s.fastq.gz" % ds)), str(infile),
                        "--json=" + str(out / "SP5" / ("cutadapt_SP5_%s.json" % ds))],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    suffix = "_%s.fastq.gz" % ds
    ids = sorted(os.path.basename(p)[:-len(suffix)] for p in glob.glob(str(out / "SP5" / ("*" + suffix))))
    assert len(ids) == 13 and "unknown" in ids          # every bin file exists, even if empty
    ids = [i for i in ids if "unknown" not in i]         # 02:75-80
    rep = json.load(open(out / "SP5" / ("cutadapt_SP5_%s.json" % ds)))
    assert rep["read_counts"]["input"] == rs.n_reads

    # round 2 (02:91-103)
    for ident in ids:
        r = subprocess.run([shim, "--action=trim", "-e", "0.1", "-j", "24", "--rc", "-a", "file:" + rev,
                            "-o", str(out / "SP27" / ("{name}_%s_%s.fastq.gz" % (ident, ds))),
                            str(out / "SP5" / ("%s_%s.fastq.gz" % (ident, ds))),
                            "--json=" + str(out / "SP27" / ("%s_%s.json" % (ident, ds)))],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    # clean-up (02:110-118)
    for p in glob.glob(str(out / "**" / "*unknown*"), recursive=True):
        os.remove(p)
    for k in ("009", "010", "011", "012"):
        for p in glob.glob(str(out / "**" / ("*SP27_%s*" % k)), recursive=True):
            os.remove(p)

    expected, rec0 = _expected_tree(rs, ds)
    got = sorted(os.path.relpath(p, out) for p in glob.glob(str(out / "SP27" / "*.fastq.gz")))
    assert got == sorted(expected)                       # 96 files
    assert len(got) == 96
    for rel, data in expected.items():
        assert _read_gz(out / rel) == data, rel
    assert rep["read_counts"]["read1_with_adapter"] == int((rec0["adapter"] >= 0).sum())
    assert rep["read_counts"]["reverse_complemented"] == int(rec0["is_rc"].sum())

    # the fused two-round command leaves the same SP27 tree
    out2 = work / "demuxed_fused"
    r = subprocess.run([sys.executable, "-m", "orcdemux.cli", "two-round", str(infile), "--sp5", fwd, "--sp27", rev,
                        "--outdir", str(out2)], capture_output=True, text=True,
                       env=dict(os.environ, PYTHONPATH=H.PKG))
    assert r.returncode == 0, r.stderr
    got2 = sorted(os.path.relpath(p, out2) for p in glob.glob(str(out2 / "SP27" / "*.fastq.gz")))
    assert got2 == sorted(expected)
    for rel, data in expected.items():
        assert _read_gz(out2 / rel) == data, rel

    # the same in many small batches: reader thread -> 3 slots -> writer pool, order kept per bin
    out3 = work / "demuxed_fused_small"
    r = subprocess.run([sys.executable, "-m", "orcdemux.cli", "two-round", str(infile), "--sp5", fwd, "--sp27", rev,
                        "--outdir", str(out3), "-j", "3"], capture_output=True, text=True,
                       env=dict(os.environ, PYTHONPATH=H.PKG, ORCDEMUX_BATCH_READS="700"))
    assert r.returncode == 0, r.stderr
    for rel, data in expected.items():
        assert _read_gz(out3 / rel) == data, rel
    assert json.load(open(out3 / "SP27" / ("orcdemux_%s.json" % ds)))["reads"] == rs.n_reads


def test_two_round_on_two_ranks_equals_one(tmp_path):
    """`two-round --gpus 2`: the batches of the input are dealt to two ranks (two GPUs when the box has them,
    else both on the one GPU), every rank writes part files, rank 0 stitches them in batch order.  The tree
    must equal the one-GPU tree byte for byte after decompression, and so must the counts."""
    rs = synth.generate(9000, 300, 900, seed=78)
    ds = "s2"
    (tmp_path / "pychopped").mkdir()
    infile = tmp_path / "pychopped" / ("pychopped_%s.fastq" % ds)
    infile.write_bytes(rs.to_fastq_bytes())
    fwd, rev, _ = m13.write_tables(str(tmp_path / "adapters"))
    env = dict(os.environ, PYTHONPATH=H.PKG, ORCDEMUX_BATCH_READS="800")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    trees = {}
    for gpus in (1, 2, 3):
        out = tmp_path / ("demuxed%d" % gpus)
        r = subprocess.run([sys.executable, "-m", "orcdemux.cli", "two-round", str(infile), "--sp5", fwd, "--sp27", rev,
                            "--outdir", str(out), "--gpus", str(gpus), "-j", "4"], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        files = sorted(os.listdir(out / "SP27"))
        assert len([f for f in files if f.endswith(".fastq.gz")]) == 96 and not [f for f in files if ".part" in f]
        trees[gpus] = {f: _read_gz(out / "SP27" / f) for f in files if f.endswith(".fastq.gz")}
        rep = json.load(open(out / "SP27" / ("orcdemux_%s.json" % ds)))
        assert rep["reads"] == rs.n_reads and rep["gpus"] == gpus
        trees[(gpus, "bins")] = rep["bins"]
    expected, _ = _expected_tree(rs, ds)
    assert {"SP27/" + k: v for k, v in trees[1].items()} == expected
    assert trees[2] == trees[1] and trees[3] == trees[1]
    assert trees[(2, "bins")] == trees[(1, "bins")] == trees[(3, "bins")]


def test_json_report_equals_report_built_from_oracle_records(tmp_path):
    """N1: the --json of a GPU run (read counts, per-adapter totals, matches on the reverse complement,
    trimmed-length histograms by error count, bases preceding 3' matches) against the same statistics built
    here from the ORACLE's match records, for both invocation shapes of the script."""
    shim = os.path.join(H.PKG, "bin", "cutadapt")
    fwd, rev, _ = m13.write_tables(str(tmp_path / "adapters"))
    rs = synth.generate(8000, 300, 900, seed=31)
    rec0, rec1, oseq, oqual, olen = H.run_oracle(H.m13_rounds(), rs)
    infile = tmp_path / "in.fastq"
    infile.write_bytes(rs.to_fastq_bytes())
    (tmp_path / "SP5").mkdir()
    r = subprocess.run([shim, "--action=trim", "-e", "0.1", "-j", "4", "--rc", "-g", "file:" + fwd,
                        "-o", str(tmp_path / "SP5" / "{name}.fastq"), str(infile), "--json=" + str(tmp_path / "r1.json")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rep = json.load(open(tmp_path / "r1.json"))
    L = rs.lengths.astype(np.int64)

    def check(rep, rec, in_len, names, seqs, front, adjacent=None):
        has = rec["adapter"] >= 0
        assert rep["read_counts"]["input"] == rec.shape[0] and rep["read_counts"]["output"] == rec.shape[0]
        assert rep["read_counts"]["read1_with_adapter"] == int(has.sum())
        assert rep["read_counts"]["reverse_complemented"] == int((rec["is_rc"] != 0).sum())
        assert rep["basepair_counts"]["input"] == int(in_len.sum())
        removed = np.where(has, rec["query_stop"] if front else in_len - rec["query_start"], 0)
        assert rep["basepair_counts"]["output"] == int((in_len - removed).sum())
        for a, ad in enumerate(rep["adapters_read1"]):
            sel = rec["adapter"] == a
            assert ad["name"] == names[a] and ad["total_matches"] == int(sel.sum())
            assert ad["on_reverse_complement"] == int((rec["is_rc"][sel] != 0).sum())
            end = ad["five_prime_end"] if front else ad["three_prime_end"]
            assert (ad["three_prime_end"] if front else ad["five_prime_end"]) is None
            assert end["sequence"] == seqs[a] and end["matches"] == int(sel.sum()) and end["error_rate"] == 0.1
            hist = {}
            for q, e in zip(removed[sel].tolist(), rec["errors"][sel].tolist()):
                hist.setdefault(q, {}).setdefault(e, 0)
                hist[q][e] += 1
            assert {t["len"]: {e: c for e, c in enumerate(t["counts"]) if c} for t in end["trimmed_lengths"]} == hist, names[a]
            for t in end["trimmed_lengths"]:
                assert abs(t["expect"] - rec.shape[0] * 0.25 ** min(t["len"], len(seqs[a]))) < 1e-6 * max(1.0, t["expect"])
            if adjacent is not None:
                assert end["adjacent_bases"] == adjacent[a], names[a]

    n5 = [n for n, _ in m13.sp5_forward()]
    s5 = [q for _, q in m13.sp5_forward()]
    check(rep, rec0, L, n5, s5, True)
    # round 2 on one of the round-1 bins: reads of that bin as round 1 left them (the oracle's view of them)
    a0 = int(np.bincount(rec0["adapter"][rec0["adapter"] >= 0]).argmax())
    sel = np.flatnonzero(rec0["adapter"] == a0)
    (tmp_path / "SP27").mkdir()
    r = subprocess.run([shim, "--action=trim", "-e", "0.1", "-j", "4", "--rc", "-a", "file:" + rev,
                        "-o", str(tmp_path / "SP27" / "{name}.fastq"), str(tmp_path / "SP5" / (n5[a0] + ".fastq")),
                        "--json=" + str(tmp_path / "r2.json")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rep2 = json.load(open(tmp_path / "r2.json"))
    in2 = (L - rec0["query_stop"])[sel]
    sub1 = rec1[sel]
    # the base before a 3' match, in the orientation that matched: rebuild the round-1 output of each read
    comp = {65: 84, 67: 71, 71: 67, 84: 65}
    adjacent = [{"A": 0, "C": 0, "G": 0, "T": 0, "": 0} for _ in range(12)]
    for i, m in zip(sel.tolist(), sub1):
        if m["adapter"] < 0:
            continue
        o = int(rs.offsets[i])
        s = rs.seq[o:o + int(L[i])]
        if rec0["is_rc"][i]:
            s = np.array([comp.get(int(c), int(c)) for c in s[::-1]], dtype=np.uint8)
        s = s[int(rec0["query_stop"][i]):]
        if m["is_rc"]:
            s = np.array([comp.get(int(c), int(c)) for c in s[::-1]], dtype=np.uint8)
        q = int(m["query_start"])
        base = chr(int(s[q - 1])) if q > 0 else ""
        adjacent[int(m["adapter"])][base if base in "ACGT" else ""] += 1
    n27 = [n for n, _ in m13.sp27_reverse_rc()]
    s27 = [q for _, q in m13.sp27_reverse_rc()]
    check(rep2, sub1, in2, n27, s27, False, adjacent)


def test_unsupported_exits_2(tmp_path):
    shim = os.path.join(H.PKG, "bin", "cutadapt")
    r = subprocess.run([shim, "-m", "20", "-g", "ACGT", "-o", str(tmp_path / "{name}.fq"), "in.fq"],
                       capture_output=True, text=True)
    assert r.returncode == 2 and "unsupported" in r.stderr


def test_raw_text_layout_equals_blob_layout(tmp_path):
    from orcdemux import engine as E
    from orcdemux import fastq as F
    rs = synth.generate(3000, 300, 900, seed=5)
    p = tmp_path / "x.fastq"
    p.write_bytes(rs.to_fastq_bytes())
    rd = F.FastqReader(str(p), max_reads=4096, max_bytes=1 << 23, keep=1, ahead=1)
    tb = next(iter(rd))
    assert tb.n_reads == rs.n_reads
    with E.Engine(E.m13_rounds(), max_reads=4096, max_bytes=1 << 23, n_slots=1) as eng:
        a = eng.run(tb)
        b = eng.run(rs)
    assert a.fastq.tobytes() == b.fastq.tobytes()
    assert np.array_equal(a.bin, b.bin)
    for x, y in zip(a.matches, b.matches):
        assert H.diff_matches(x, y)[1] == 0


def test_cli_anchored_and_no_indels(tmp_path):
    """config 4 call shape (-g ^file: --no-indels) and --no-indels on regular adapters via the shim."""
    import oracle
    shim = os.path.join(H.PKG, "bin", "cutadapt")
    fwd, rev, var = m13.write_tables(str(tmp_path / "adapters"))
    # anchored Hamming path
    rs = synth.generate(4000, 300, 600, seed=1004, anchored=True)
    infile = tmp_path / "a.fastq"
    infile.write_bytes(rs.to_fastq_bytes())
    out = tmp_path / "anch"
    out.mkdir()
    r = subprocess.run([shim, "--action=trim", "-e", "0.1", "--no-indels", "--rc", "-g", "^file:" + var,
                        "-o", str(out / "{name}.fastq"), str(infile), "--json=" + str(out / "r.json")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    names = [n for n, _ in m13.variable_all()]
    seqs = [q for _, q in m13.variable_all()]
    sets = [(oracle.AdapterSet(seqs, oracle.PREFIX, 0.1, 3, indels=False), 1)]
    rec0, _, oseq, oqual, olen = oracle.demux_batch(sets, rs.seq, rs.qual, rs.offsets, rs.lengths, n_threads=4)
    exp = {n: [] for n in names + ["unknown"]}
    for i in range(rs.n_reads):
        a = int(rec0["adapter"][i])
        nm = rs.read(i)[0] + (" rc" if rec0["is_rc"][i] else "")
        o, L = int(rs.offsets[i]), int(olen[i])
        exp[names[a] if a >= 0 else "unknown"].append(
            b"@" + nm.encode() + b"\n" + oseq[o:o + L].tobytes() + b"\n+\n" + oqual[o:o + L].tobytes() + b"\n")
    assert sorted(os.listdir(out)) == sorted([n + ".fastq" for n in exp] + ["r.json"])
    for n, recs in exp.items():
        assert (out / (n + ".fastq")).read_bytes() == b"".join(recs), n
    rep = json.load(open(out / "r.json"))
    assert rep["read_counts"]["input"] == rs.n_reads
    assert rep["read_counts"]["read1_with_adapter"] == int((rec0["adapter"] >= 0).sum())
    assert rep["read_counts"]["reverse_complemented"] == int((rec0["is_rc"] != 0).sum())
    for a, ad in enumerate(rep["adapters_read1"]):
        sel = rec0["adapter"] == a
        assert ad["name"] == names[a] and ad["total_matches"] == int(sel.sum())
        assert ad["on_reverse_complement"] == int((rec0["is_rc"][sel] != 0).sum())
        end = ad["five_prime_end"]
        assert end["type"] == "anchored_five_prime" and end["sequence"] == seqs[a] and ad["three_prime_end"] is None
        hist = {}
        for q, e in zip(rec0["query_stop"][sel].tolist(), rec0["errors"][sel].tolist()):
            hist.setdefault(q, {}).setdefault(e, 0)
            hist[q][e] += 1
        assert {t["len"]: {e: c for e, c in enumerate(t["counts"]) if c} for t in end["trimmed_lengths"]} == hist
    # anchored adapters with indels are refused, loudly
    r = subprocess.run([shim, "-g", "^file:" + var, "-o", str(out / "{name}.fq"), str(infile)],
                       capture_output=True, text=True)
    assert r.returncode == 2 and "unsupported" in r.stderr
    # --no-indels on the regular 5' adapters
    rs2 = synth.generate(3000, 300, 600, seed=8)
    in2 = tmp_path / "b.fastq.gz"
    with gzip.open(in2, "wb", compresslevel=1) as fh:
        fh.write(rs2.to_fastq_bytes())
    out2 = tmp_path / "noindel"
    out2.mkdir()
    r = subprocess.run([shim, "--action=trim", "-e", "0.1", "--no-indels", "--rc", "-g", "file:" + fwd,
                        "-o", str(out2 / "{name}.fastq.gz"), str(in2)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    n5 = [n for n, _ in m13.sp5_forward()]
    s5 = [q for _, q in m13.sp5_forward()]
    sets = [(oracle.AdapterSet(s5, oracle.FRONT, 0.1, 3, indels=False), 1)]
    rec0, _, oseq, oqual, olen = oracle.demux_batch(sets, rs2.seq, rs2.qual, rs2.offsets, rs2.lengths, n_threads=4)
    for a, nm in enumerate(n5):
        exp = b"".join(b"@" + (rs2.read(int(i))[0] + (" rc" if rec0["is_rc"][i] else "")).encode() + b"\n" +
                       oseq[int(rs2.offsets[i]):int(rs2.offsets[i]) + int(olen[i])].tobytes() + b"\n+\n" +
                       oqual[int(rs2.offsets[i]):int(rs2.offsets[i]) + int(olen[i])].tobytes() + b"\n"
                       for i in np.flatnonzero(rec0["adapter"] == a))
        assert _read_gz(out2 / (nm + ".fastq.gz")) == exp, nm


def test_two_round_on_empty_and_tiny_inputs(tmp_path):
    """The fused command (members coded on the GPU by default, and with --host-gzip) on an empty input and on a
    single read: all 96 files exist, every one is a valid .gz, the read lands where the oracle puts it."""
    fwd, rev, _ = m13.write_tables(str(tmp_path / "adapters"))
    one = synth.generate(40, 300, 900, seed=9)
    rec0, rec1, oseq, oqual, olen = H.run_oracle(H.m13_rounds(), one)
    keep = [i for i in range(one.n_reads) if rec0["adapter"][i] >= 0 and 0 <= rec1["adapter"][i] < 8][:1]
    assert keep
    i = keep[0]
    name, sq, ql = one.read(i)
    cases = {"empty": b"", "one": ("@%s\n%s\n+\n%s\n" % (name, sq, ql)).encode()}
    for tag, text in cases.items():
        for extra in ([], ["--host-gzip"]):
            (tmp_path / "pychopped").mkdir(exist_ok=True)
            infile = tmp_path / "pychopped" / ("pychopped_%s.fastq" % tag)
            infile.write_bytes(text)
            out = tmp_path / ("demuxed_%s_%d" % (tag, len(extra)))
            r = subprocess.run([sys.executable, "-m", "orcdemux.cli", "two-round", str(infile), "--sp5", fwd, "--sp27", rev,
                                "--outdir", str(out)] + extra, capture_output=True, text=True,
                               env=dict(os.environ, PYTHONPATH=H.PKG))
            assert r.returncode == 0, r.stderr
            files = sorted(glob.glob(str(out / "SP27" / "*.fastq.gz")))
            assert len(files) == 96
            total = b"".join(_read_gz(f) for f in files)
            if tag == "empty":
                assert total == b""
            else:
                o, L = int(one.offsets[i]), int(olen[i])
                nm = name + (" rc" if rec0["is_rc"][i] else "") + (" rc" if rec1["is_rc"][i] else "")
                assert total == b"@" + nm.encode() + b"\n" + oseq[o:o + L].tobytes() + b"\n+\n" + oqual[o:o + L].tobytes() + b"\n"
