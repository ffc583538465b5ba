"""GPU parity tests: the CUDA path, called through the C ABI (orcdemux.Engine is a thin
ctypes wrapper), against the CPU oracle on the same seeded inputs.  Bit-exact: adapter,
orientation, all six alignment fields, bin, trimmed bytes, order inside each bin."""
import os

import numpy as np
import pytest

import helpers as H
from orcdemux import engine as E
from orcdemux import synth

pytestmark = pytest.mark.gpu


def _engine(n_reads, n_bytes, **kw):
    return E.Engine(E.m13_rounds(), device=0, max_reads=max(n_reads, 1), max_bytes=max(n_bytes, 64),
                    n_slots=kw.pop("n_slots", 1), **kw)


def _expected_fastq(rs, rec0, rec1, oseq, oqual, olen, n_bins, bin_of):
    """Per-bin FASTQ bytes the reference pipeline would leave on disk (input order kept)."""
    bins = [[] for _ in range(n_bins)]
    names, noff = rs.names.tobytes(), rs.name_offsets.tolist()
    sq, ql = oseq.tobytes(), oqual.tobytes()
    a0 = rec0["adapter"].tolist()
    a1 = rec1["adapter"].tolist() if rec1 is not None else None
    rc0 = ((rec0["adapter"] >= 0) & (rec0["is_rc"] != 0)).tolist()
    rc1 = ((rec1["adapter"] >= 0) & (rec1["is_rc"] != 0)).tolist() if rec1 is not None else None
    offs, lens = rs.offsets.tolist(), olen.tolist()
    for r in range(rs.n_reads):
        name = names[noff[r]:noff[r + 1]]
        if rc0[r]:
            name += b" rc"
        if rc1 is not None and rc1[r]:
            name += b" rc"
        o, L = offs[r], lens[r]
        b = bin_of(a0[r], a1[r] if a1 is not None else -1)
        bins[b].append(b"@" + name + b"\n" + sq[o:o + L] + b"\n+\n" + ql[o:o + L] + b"\n")
    return [b"".join(x) for x in bins]


FULL = os.environ.get("ORC_TEST_FULL", "1") == "1"      # 0: the BASELINE-size tests run on 1/16 of their reads
NCPU = os.cpu_count() or 8


def _check(rs, eng=None):
    own = eng is None
    if own:
        eng = _engine(rs.n_reads, rs.seq.shape[0])
    try:
        res = eng.run(rs)
        rec0, rec1, oseq, oqual, olen = H.run_oracle(H.m13_rounds(), rs)
        idx, nbad = H.diff_matches(rec0, res.matches[0])
        assert nbad == 0, "round 1 differs at reads %s" % idx
        idx, nbad = H.diff_matches(rec1, res.matches[1])
        assert nbad == 0, "round 2 differs at reads %s" % idx
        assert np.array_equal(res.out_len, olen)
        exp_bin = np.array([eng.bin_id(int(a), int(b)) for a, b in zip(rec0["adapter"], rec1["adapter"])],
                           dtype=np.int32)
        assert np.array_equal(res.bin, exp_bin)
        exp = _expected_fastq(rs, rec0, rec1, oseq, oqual, olen, eng.n_bins, eng.bin_id)
        for b in range(eng.n_bins):
            assert res.bin_bytes(b) == exp[b], "bin %d bytes differ" % b
            assert int(res.bin_counts[b]) == int((exp_bin == b).sum())
        return res
    finally:
        if own:
            eng.close()


def test_synthetic_coi_20k():
    _check(synth.generate(20000, 300, 900, seed=1002))


def test_synthetic_rrna_2k():
    _check(synth.generate(2000, 1000, 3500, seed=1003))


def test_edge_cases():
    sp5 = [s for _, s in synth.m13.sp5_forward()]
    sp27 = [s for _, s in synth.m13.sp27_reverse_rc()]
    ins = "ACGTTGCA" * 40
    q = lambda s: "I" * len(s)
    recs = []
    def add(name, s):
        recs.append((name, s, q(s)))
    add("exact", sp5[6] + ins + sp27[2])
    add("exact_rc", synth.m13.revcomp(sp5[6] + ins + sp27[2]))
    add("empty", "")
    add("one", "A")
    add("only_adapter", sp5[0])
    add("adapter_twice", sp5[3] + ins + sp5[4] + ins + sp27[0])
    add("tie_cag", "CAG" + "T" * 300)
    add("ends_gtc", sp5[1] + "A" * 200 + "GTC")
    add("all_n", "N" * 500)
    add("lower", (sp5[9] + ins + sp27[7]).lower())
    add("trunc5", sp5[2][20:] + ins + sp27[11])
    add("trunc3", sp5[2] + ins + sp27[11][:30])
    add("short", sp5[5][:10])
    add("no_adapters", ins)
    add("iupac_read", sp5[8][:30] + "R" + sp5[8][31:] + ins + sp27[4])
    add("long", sp5[10] + ins * 60 + sp27[5])
    rs = synth.from_records(recs)
    res = _check(rs)
    assert res.matches[0]["adapter"][0] == 6 and res.matches[1]["adapter"][0] == 2


def test_empty_batch():
    rs = synth.from_records([])
    eng = _engine(16, 1024)
    try:
        res = eng.run(rs)
        assert res.n_reads == 0 and res.fastq.shape[0] == 0 and int(res.bin_counts.sum()) == 0
    finally:
        eng.close()


def test_slots_and_resident_relaunch():
    """Two batches in flight on two slots; re-launching a resident batch gives the same bytes."""
    a = synth.generate(3000, 300, 900, seed=7)
    b = synth.generate(2500, 300, 900, seed=8)
    eng = _engine(3000, max(a.seq.shape[0], b.seq.shape[0]), n_slots=2)
    try:
        eng.submit(0, a)
        eng.submit(1, b)
        ra, rb = eng.wait(0), eng.wait(1)
        eng.launch(0)
        eng.download(0)
        ra2 = eng.wait(0)
        assert ra.fastq.tobytes() == ra2.fastq.tobytes()
        assert np.array_equal(ra.bin, ra2.bin)
        t = eng.timings(0)
        assert t["kernel_launches"] >= 15 and t["total_ms"] > 0      # counted launch by launch in orc_launch
        tot = eng.counts()
        assert int(tot.sum()) == 2 * a.n_reads + b.n_reads
    finally:
        eng.close()
    _check(a)
    _check(b)


def test_qualities_read_in_place_from_pinned_host_memory():
    """orc_params.qual_zero_copy: the quality strings are not copied to the device, emit_kernel reads them in
    place from the caller's page-locked buffer.  Same bytes as the copying path and as the oracle -- for a
    whole pinned batch, for sub-batches that are views into one pinned buffer (arbitrary alignment of the
    quality pointer, the last one ending where the buffer's slack begins), with dropped bins, and for a
    pageable buffer (copied as before)."""
    rs = synth.generate(6000, 300, 900, seed=31)
    ref = _check(rs)                                            # copying path, checked against the oracle
    pinned = E.pin_readset(rs)
    drop = np.zeros(169, dtype=np.uint8)
    drop[0::13] = 1                                             # no SP5 match
    drop[:13] = 1                                               # no SP27 match
    for zc_rs in (pinned, rs):                                  # page-locked: in place; pageable: falls back to the copy
        eng = _engine(rs.n_reads, rs.seq.shape[0], qual_zero_copy=True)
        try:
            res = eng.run(zc_rs)
            assert res.fastq.tobytes() == ref.fastq.tobytes()
            assert np.array_equal(res.bin, ref.bin) and np.array_equal(res.out_len, ref.out_len)
        finally:
            eng.close()
    # sub-batches as views of the pinned blobs, two in flight, with dropped bins
    n_sub = 5
    per = (rs.n_reads + n_sub - 1) // n_sub
    engs = [_engine(per, rs.seq.shape[0], n_slots=2, drop_bins=drop, qual_zero_copy=z) for z in (False, True)]
    try:
        outs = [[], []]
        for i in range(n_sub):
            lo, hi = i * per, min(rs.n_reads, (i + 1) * per)
            b0 = int(pinned.offsets[lo])
            b1 = int(pinned.offsets[hi - 1]) + int(pinned.lengths[hi - 1])
            n0, n1 = int(pinned.name_offsets[lo]), int(pinned.name_offsets[hi])
            off = E.pinned_empty(hi - lo, np.uint64)
            off[...] = pinned.offsets[lo:hi] - np.uint64(b0)
            noff = E.pinned_empty(hi - lo + 1, np.uint64)
            noff[...] = pinned.name_offsets[lo:hi + 1] - np.uint64(n0)
            sub = synth.ReadSet(pinned.seq[b0:b1], pinned.qual[b0:b1], off, pinned.lengths[lo:hi], pinned.names[n0:n1], noff, {})
            for k, eng in enumerate(engs):
                eng.submit(i % 2, sub)
                r = eng.wait(i % 2)
                outs[k].append((r.fastq.tobytes(), r.bin.copy(), r.bin_offsets.copy()))
        for a, b in zip(*outs):
            assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        assert any((a[1] < 0).any() for a in outs[1]) and any(len(a[0]) for a in outs[1])
    finally:
        for eng in engs:
            eng.close()


def test_bins_as_gzip_members_made_on_the_device(tmp_path):
    """orc_params.emit_gzip: every bin of a batch comes back as one gzip member coded by the gz_* kernels
    (csrc/orc_gz.cuh).  zlib must inflate every member to exactly the bytes the plain path emits for that bin
    (which _check compares with the oracle), CRC-32 / ISIZE / the "OC" size field must hold, with dropped bins,
    empty bins, an empty batch, two slots in flight -- and the writer must put the members into .gz files as
    they are and into plain files inflated."""
    import gzip
    import zlib
    from test_gz import check_members
    from orcdemux import fastq as F
    rs = synth.generate(20000, 300, 900, seed=33)
    ref = _check(rs)
    drop = np.zeros(169, dtype=np.uint8)
    drop[0::13] = 1
    drop[:13] = 1
    plain = _engine(rs.n_reads, rs.seq.shape[0], drop_bins=drop)
    gz = _engine(rs.n_reads, rs.seq.shape[0], drop_bins=drop, emit_gzip=True, n_slots=2, want_matches=False)
    try:
        want = plain.run(rs)
        got = gz.run(rs)
        pieces = [want.fastq[int(want.bin_offsets[b]):int(want.bin_offsets[b + 1])].tobytes() for b in range(169)]
        assert sum(len(p) for p in pieces) > 0 and any(len(p) == 0 for p in pieces)
        check_members(got.fastq.tobytes(), got.bin_offsets, pieces)
        assert np.array_equal(got.bin, want.bin) and np.array_equal(got.bin_counts, want.bin_counts)
        t = gz.timings(0)
        assert t["gzip_bytes"] == got.fastq.shape[0] and t["gzip_ms"] > 0
        tl = gz.timeline(0)                 # upload reached, H2D, matching, emit, gzip, D2H: in this order
        assert all(b >= a for a, b in zip(tl, tl[1:])) and tl[4] - tl[3] > 0
        assert got.fastq.shape[0] < 0.62 * want.fastq.shape[0]
        # the writer: members as they are into .gz files, inflated into plain ones
        paths = [None if drop[b] else str(tmp_path / ("bin%03d.fastq%s" % (b, ".gz" if b % 2 else ""))) for b in range(169)]
        w = F.BinWriters(paths, 1, threads=4)
        t0 = w.write_batch(got, members=True)
        w.wait(t0)
        # a second batch (slot 1) appended to the same files
        sub = synth.generate(5000, 300, 900, seed=34)
        gz.submit(1, sub)
        got2 = gz.wait(1)
        want2 = plain.run(sub)
        pieces2 = [want2.fastq[int(want2.bin_offsets[b]):int(want2.bin_offsets[b + 1])].tobytes() for b in range(169)]
        check_members(got2.fastq.tobytes(), got2.bin_offsets, pieces2)
        w.wait(w.write_batch(got2, members=True))
        w.close()
        for b, path in enumerate(paths):
            if path is None:
                continue
            data = gzip.open(path, "rb").read() if path.endswith(".gz") else open(path, "rb").read()
            assert data == pieces[b] + pieces2[b], path
        assert w.bytes_written[14] == len(pieces[14]) + len(pieces2[14])
        # an empty batch: no members at all
        e = gz.run(synth.from_records([]))
        assert e.fastq.shape[0] == 0 and int(e.bin_offsets[-1]) == 0
    finally:
        plain.close()
        gz.close()
    assert ref.n_reads == rs.n_reads


def test_gzip_members_other_shapes():
    """emit_gzip beyond the COI two-round case: one round (13 bins, nothing dropped), kilobase reads, one read of
    120 kb beside short ones (a member far larger than a tile of chunks), reads with lower case / N / IUPAC bytes,
    and batches reusing one slot with different sizes (stale bytes behind a smaller batch must not leak)."""
    import zlib
    from test_gz import check_members
    import random
    rnd = random.Random(77)
    long_read = "".join(rnd.choice("ACGT") for _ in range(120000))
    odd = [("o%d" % i, "".join(rnd.choice("ACGTNacgtnRYKM") for _ in range(rnd.randint(1, 700))), None) for i in range(300)]
    odd = [(n, q, "".join(chr(33 + rnd.randint(0, 60)) for _ in q)) for n, q, _ in odd]
    sets = [synth.generate(3000, 1000, 3500, seed=1003),
            synth.from_records([("long", long_read, "I" * len(long_read))] + odd),
            synth.generate(700, 300, 900, seed=3)]
    for rounds in (E.m13_rounds(), E.m13_rounds()[:1]):
        cap_r = max(x.n_reads for x in sets)
        cap_b = max(int(x.seq.shape[0]) for x in sets) + 64
        kw = dict(device=0, max_reads=cap_r, max_bytes=cap_b, max_name_bytes=64 * cap_r, n_slots=1, want_matches=False)
        with E.Engine(rounds, **kw) as plain, E.Engine(rounds, emit_gzip=True, **kw) as gz:
            for rs in sets:                                         # one slot, batches of different sizes in turn
                want, got = plain.run(rs), gz.run(rs)
                pieces = [want.bin_bytes(b) for b in range(plain.n_bins)]
                check_members(got.fastq.tobytes(), got.bin_offsets, pieces)
                assert np.array_equal(got.bin_counts, want.bin_counts)
                assert sum(len(p) for p in pieces) == int(want.bin_offsets[-1])


def test_gzip_flat_code_fallback(monkeypatch):
    """What orc_wait() does when the members of a batch outgrow their arena (a sampled histogram far off the
    batch's bytes): the batch is coded again with 8- and 9-bit codes, which fit by construction.
    ORC_GZ_FORCE_FLAT takes that path for every batch."""
    from test_gz import check_members
    rs = synth.generate(5000, 300, 900, seed=35)
    with _engine(rs.n_reads, rs.seq.shape[0]) as plain, _engine(rs.n_reads, rs.seq.shape[0], emit_gzip=True) as gz:
        want = plain.run(rs)
        monkeypatch.setenv("ORC_GZ_FORCE_FLAT", "1")
        got = gz.run(rs)
        monkeypatch.delenv("ORC_GZ_FORCE_FLAT")
        pieces = [want.bin_bytes(b) for b in range(plain.n_bins)]
        check_members(got.fastq.tobytes(), got.bin_offsets, pieces)
        assert 1.0 < got.fastq.shape[0] / want.fastq.shape[0] < 1.15         # 8 bits per byte and the frames
        again = gz.run(rs)                                                    # and back to the batch's own code
        check_members(again.fastq.tobytes(), again.bin_offsets, pieces)
        assert again.fastq.shape[0] < 0.62 * want.fastq.shape[0]


def test_config2_full_size_every_read():
    """BASELINE configs[1] at its full size (1 Mi COI reads, seed 1002): the oracle on EVERY read -- all eight
    match fields of both rounds, trimmed length, bin, and the bytes of all 169 bins -- plus the
    size-independent invariants."""
    rs = synth.generate((1 << 20) if FULL else (1 << 16), 300, 900, seed=1002, workers=min(8, NCPU))
    eng = _engine(rs.n_reads, rs.seq.shape[0], want_matches=True)
    try:
        res = eng.run(rs)
    finally:
        eng.close()
    m0, m1 = res.matches
    n = rs.n_reads
    assert int(res.bin_counts.sum()) == n
    assert int(res.bin_offsets[-1]) == res.fastq.shape[0]
    # FASTQ well-formed: 4 lines per record, record count == n
    assert int((res.fastq == 10).sum()) == 4 * n
    # trimmed length arithmetic
    L = rs.lengths.astype(np.int64)
    l1 = np.where(m0["adapter"] >= 0, L - m0["query_stop"], L)
    l2 = np.where(m1["adapter"] >= 0, m1["query_start"], l1)
    assert np.array_equal(l2, res.out_len.astype(np.int64))
    # error-rate rule: errors <= floor(aligned adapter length * 0.1)
    for m in (m0, m1):
        has = m["adapter"] >= 0
        alen = (m["ref_stop"] - m["ref_start"])[has]
        assert np.all(m["errors"][has] <= alen // 10)
        assert np.all(alen >= 3)
    # unknown in round 1 never enters round 2
    assert np.all(m1["adapter"][m0["adapter"] < 0] == -1)
    # the oracle on every read
    rec0, rec1, oseq, oqual, olen = H.run_oracle(H.m13_rounds(), rs, n_threads=NCPU)
    idx, nbad = H.diff_matches(rec0, m0)
    assert nbad == 0, ("round 1", nbad, idx)
    idx, nbad = H.diff_matches(rec1, m1)
    assert nbad == 0, ("round 2", nbad, idx)
    assert np.array_equal(res.out_len, olen)
    exp_bin = ((rec0["adapter"] + 1) + 13 * (rec1["adapter"] + 1)).astype(np.int32)
    assert np.array_equal(res.bin, exp_bin)
    exp = _expected_fastq(rs, rec0, rec1, oseq, oqual, olen, 169, lambda a, b: (a + 1) + 13 * (b + 1))
    for b in range(169):
        assert res.bin_bytes(b) == exp[b], "bin %d bytes differ" % b
    # truth agreement is high (not exact: errors can push a read to unknown)
    t5 = rs.truth["sp5"]
    ok = (m0["adapter"] + 1 == t5) | (t5 == 0)
    assert ok.mean() > 0.95


def test_config3_rrna_256k_every_read():
    """BASELINE configs[2] (rRNA-cistron reads 1-3.5 kb, seed 1003) on 256 Ki reads, the oracle on every read
    (the full 10 M reads would keep the CPU oracle busy for about an hour; SURVEY 8d allows a subsample)."""
    rs = synth.generate((1 << 18) if FULL else (1 << 14), 1000, 3500, seed=1003, workers=min(8, NCPU))
    with _engine(rs.n_reads, rs.seq.shape[0], want_matches=True) as eng:
        res = eng.run(rs)
    rec0, rec1, oseq, oqual, olen = H.run_oracle(H.m13_rounds(), rs, n_threads=NCPU)
    idx, nbad = H.diff_matches(rec0, res.matches[0])
    assert nbad == 0, ("round 1", nbad, idx)
    idx, nbad = H.diff_matches(rec1, res.matches[1])
    assert nbad == 0, ("round 2", nbad, idx)
    assert np.array_equal(res.out_len, olen)
    exp = _expected_fastq(rs, rec0, rec1, oseq, oqual, olen, 169, lambda a, b: (a + 1) + 13 * (b + 1))
    for b in range(169):
        assert res.bin_bytes(b) == exp[b], "bin %d bytes differ" % b


def test_config4_both_arms_full_size():
    """BASELINE configs[3] on 1 Mi reads (seed 1004, bare index at offset 0), BOTH arms against the oracle on
    every read: `-g ^file:M13_variable_indices_all.fa --no-indels` (anchored Hamming path) and
    `-g file:M13_variable_indices_all.fa` (24 unanchored adapters with indels: the 64-bit scan, two lane banks)."""
    import oracle
    from orcdemux import m13
    from orcdemux.lib import ORC_FRONT, ORC_PREFIX
    var = m13.variable_all()
    names, seqs = [n for n, _ in var], [s for _, s in var]
    rs = synth.generate((1 << 20) if FULL else (1 << 16), 300, 900, seed=1004, anchored=True, workers=min(8, NCPU))
    for kind, okind, indels in ((ORC_PREFIX, oracle.PREFIX, False), (ORC_FRONT, oracle.FRONT, True)):
        rnd = E.Round(names, seqs, kind, 0.1, 3, indels, True)
        with E.Engine([rnd], max_reads=rs.n_reads, max_bytes=int(rs.seq.shape[0]), n_slots=1) as eng:
            res = eng.run(rs)
            assert eng.n_bins == 25
        sets = [(oracle.AdapterSet(seqs, okind, 0.1, 3, indels=indels), 1)]
        rec0, _, oseq, oqual, olen = oracle.demux_batch(sets, rs.seq, rs.qual, rs.offsets, rs.lengths, n_threads=NCPU)
        idx, nbad = H.diff_matches(rec0, res.matches[0])
        assert nbad == 0, (kind, nbad, idx)
        assert np.array_equal(res.out_len, olen)
        assert np.array_equal(res.bin, (rec0["adapter"] + 1).astype(np.int32))
        exp = _expected_fastq(rs, rec0, None, oseq, oqual, olen, 25, lambda a, b: a + 1)
        for b in range(25):
            assert res.bin_bytes(b) == exp[b], "arm %d bin %d bytes differ" % (kind, b)
        assert (rec0["adapter"] >= 0).mean() > 0.5 and (rec0["adapter"] >= 16).sum() > 1000


def test_pair_arena_overflow_is_rerun():
    """A batch with more candidate pairs than the slot's arenas hold is detected by its counters and run again
    with worst-case arenas (orc_api.cu grow_pair_arenas): same results as without the squeeze."""
    import random
    import test_hostsim as TH
    rnd = random.Random(31)
    f = [s for _, s in synth.m13.sp5_forward()]
    b = [s for _, s in synth.m13.sp27_reverse_rc()]
    rs = TH._adversarial_reads(rnd, f, b, 4000)
    ref = _check(rs)
    os.environ["ORC_PAIR_CAP"] = "64"
    try:
        eng = _engine(rs.n_reads, rs.seq.shape[0], n_slots=2)
    finally:
        os.environ.pop("ORC_PAIR_CAP", None)
    try:
        got = _check(rs, eng)           # first run overflows and is repeated inside orc_wait
        assert got.fastq.tobytes() == ref.fastq.tobytes()
        eng.submit(1, rs)               # the second slot grows on its own
        again = eng.wait(1)
        assert again.fastq.tobytes() == ref.fastq.tobytes()
        again = eng.run(rs)             # the grown arenas are kept
        assert again.fastq.tobytes() == ref.fastq.tobytes()
    finally:
        eng.close()


def test_anchored_hamming_path_config4():
    """BASELINE config 4: anchored -g ^file:M13_variable_indices_all.fa --no-indels (24 x 17-mers)."""
    import oracle
    from orcdemux import m13
    from orcdemux.lib import ORC_PREFIX
    var = m13.variable_all()
    rs = synth.generate(50000, 300, 900, seed=1004, anchored=True)
    rnd = E.Round([n for n, _ in var], [s for _, s in var], ORC_PREFIX, 0.1, 3, False, True)
    with E.Engine([rnd], max_reads=rs.n_reads, max_bytes=int(rs.seq.shape[0]), n_slots=1) as eng:
        res = eng.run(rs)
        t = eng.timings(0)
        assert eng.n_bins == 25
    sets = [(oracle.AdapterSet([s for _, s in var], oracle.PREFIX, 0.1, 3, indels=False), 1)]
    rec0, _, oseq, oqual, olen = oracle.demux_batch(sets, rs.seq, rs.qual, rs.offsets, rs.lengths, n_threads=8)
    assert H.diff_matches(rec0, res.matches[0])[1] == 0
    assert np.array_equal(res.out_len, olen)
    exp_bin = (rec0["adapter"] + 1).astype(np.int32)
    assert np.array_equal(res.bin, exp_bin)
    # bytes of one bin
    b = int(np.bincount(exp_bin).argmax())
    exp = b"".join(b"@" + rs.read(int(r))[0].encode() + (b" rc" if rec0["is_rc"][r] else b"") + b"\n" +
                   oseq[int(rs.offsets[r]):int(rs.offsets[r]) + int(olen[r])].tobytes() + b"\n+\n" +
                   oqual[int(rs.offsets[r]):int(rs.offsets[r]) + int(olen[r])].tobytes() + b"\n"
                   for r in np.flatnonzero(exp_bin == b))
    assert res.bin_bytes(b) == exp
    assert (rec0["adapter"] >= 0).mean() > 0.5 and t["total_ms"] > 0


def test_adversarial_and_random_adapter_sets():
    """The kernels (not only their host simulation) on adversarial reads and random adapter sets:
    shared prefixes of any length (trigger filter on or off), 1..32 adapters of 3..64 nt, error
    rates up to 0.4, both adapter types in either round order, with and without --rc."""
    import random
    import oracle
    import test_hostsim as TH
    from orcdemux.lib import ORC_BACK, ORC_FRONT
    rnd = random.Random(2024)
    for trial in range(10):
        nf = rnd.randint(1, 32)
        nb = rnd.randint(1, min(32, 512 // (nf + 1) - 1))        # at most 512 bins (MAX_BINS)
        mk = lambda: "".join(rnd.choice("ACGT") for _ in range(rnd.choice([3, 8, 17, 20, 33, 57, 64])))
        shared = mk()[:rnd.choice([4, 10, 17, 25, 32])]
        f = [((shared if rnd.random() < 0.8 else "") + mk())[:64] for _ in range(nf)]
        b = [((shared if rnd.random() < 0.8 else "") + mk())[:64] for _ in range(nb)]
        e = rnd.choice([0.0, 0.1, 0.1, 0.2, 0.3, 0.4])
        ov = rnd.choice([1, 3, 3, 5, 8])
        rc = rnd.choice([0, 1, 1])
        spec = [(f, oracle.FRONT, e, ov, rc), (b, oracle.BACK, e, ov, rc)]
        if trial % 3 == 1:
            spec = spec[::-1]
        if trial % 4 == 3:
            spec = spec[:1]
        rs = TH._adversarial_reads(rnd, f, b, 3000)
        rounds = [E.Round([str(i) for i in range(len(x[0]))], x[0], ORC_FRONT if x[1] == oracle.FRONT else ORC_BACK,
                          x[2], x[3], True, bool(x[4])) for x in spec]
        with E.Engine(rounds, max_reads=rs.n_reads, max_bytes=int(rs.seq.shape[0]) + 64, n_slots=1) as eng:
            res = eng.run(rs)
        rec = H.run_oracle(spec, rs)
        idx, nbad = H.diff_matches(rec[0], res.matches[0])
        assert nbad == 0, (trial, "round 1", idx[:3], rs.read(int(idx[0]))[1])
        if len(spec) > 1:
            idx, nbad = H.diff_matches(rec[1], res.matches[1])
            assert nbad == 0, (trial, "round 2", idx[:3])
        assert np.array_equal(res.out_len, rec[4])


def test_iupac_adapter_sets_gpu():
    """Adapters with IUPAC wildcards through the kernels (mask comparison, N taken out of the effective
    length, U read as T), FASTQ bytes included: reads with U that both rounds reverse-complement come
    out with T (dnaio's complement table), which the emit kernel has to reproduce."""
    import random
    import oracle
    import test_hostsim as TH
    from orcdemux.lib import ORC_BACK, ORC_FRONT
    rnd = random.Random(4040)
    done = 0
    for trial in range(8):
        f, b = TH._iupac_sets(rnd, n_in_front=(trial % 2 == 1))
        e = rnd.choice([0.0, 0.1, 0.1, 0.15, 0.2])
        ov = rnd.choice([1, 3, 3, 5])
        spec = [(f, oracle.FRONT, e, ov, 1), (b, oracle.BACK, e, ov, 1)]
        if trial % 4 == 3:
            spec = spec[::-1]
        rs = TH._adversarial_reads(rnd, TH._instances(rnd, f), TH._instances(rnd, b), 1500)
        recs = [rs.read(i) for i in range(rs.n_reads)]
        recs = [(nm, sq.replace("T", "U").replace("t", "u") if i % 5 == 0 else sq, q) for i, (nm, sq, q) in enumerate(recs)]
        for a in TH._instances(rnd, f, per=1):         # reads inside a 5' adapter (last-column test, own N count)
            for _ in range(6):
                cut = a[-rnd.randint(1, len(a)):]
                recs.append(("in%d" % len(recs), cut, "I" * len(cut)))
        rs = synth.from_records(recs)
        rounds = [E.Round([str(i) for i in range(len(x[0]))], x[0], ORC_FRONT if x[1] == oracle.FRONT else ORC_BACK,
                          x[2], x[3], True, bool(x[4])) for x in spec]
        with E.Engine(rounds, max_reads=rs.n_reads, max_bytes=int(rs.seq.shape[0]) + 64,
                      max_name_bytes=int(rs.names.shape[0]) + 64, n_slots=1, emit_fastq=True, want_matches=True) as eng:
            res = eng.run(rs)
            rec0, rec1, oseq, oqual, olen = H.run_oracle(spec, rs)
            idx, nbad = H.diff_matches(rec0, res.matches[0])
            assert nbad == 0, (trial, "round 1", idx[:3], rs.read(int(idx[0]))[1])
            idx, nbad = H.diff_matches(rec1, res.matches[1])
            assert nbad == 0, (trial, "round 2", idx[:3])
            assert np.array_equal(res.out_len, olen)
            exp = _expected_fastq(rs, rec0, rec1, oseq, oqual, olen, eng.n_bins, eng.bin_id)
            for bb in range(eng.n_bins):
                assert res.bin_bytes(bb) == exp[bb], "trial %d bin %d bytes differ" % (trial, bb)
        both = (rec0["is_rc"] == 1) & (rec1["is_rc"] == 1)
        done += int(both.sum())
    assert done > 0         # some reads were flipped twice


def test_plain_and_iupac_adapters_side_by_side_gpu():
    """Plain ACGT adapters beside IUPAC ones -- in one round, or one kind per round: all compared through the masks
    (cutadapt decides per adapter; the results differ only for a read with U).  Matches, lengths and bytes against
    the oracle; a batch with a U among its bases is refused by orc_wait (u_scan_kernel), and the slot stays usable."""
    import random
    import oracle
    import test_hostsim as TH
    from orcdemux.lib import ORC_BACK, ORC_FRONT
    rnd = random.Random(809)
    for trial in range(6):
        f, b = TH._iupac_sets(rnd, n_in_front=(trial % 2 == 1))
        pf = ["".join(rnd.choice("ACGT") for _ in range(rnd.choice([8, 17, 25, 40, 64]))) for _ in range(rnd.randint(1, 5))]
        pb = ["".join(rnd.choice("ACGT") for _ in range(rnd.choice([8, 17, 25, 40, 64]))) for _ in range(rnd.randint(1, 5))]
        if trial % 3 == 0:
            f, b = (f + pf)[:16], (pb + b)[:16]
        elif trial % 3 == 1:
            f = pf
        else:
            b = pb
        e = rnd.choice([0.0, 0.1, 0.1, 0.2])
        spec = [(f, oracle.FRONT, e, 3, 1), (b, oracle.BACK, e, 3, 1)]
        rs = TH._adversarial_reads(rnd, TH._instances(rnd, f), TH._instances(rnd, b), 1200)
        rounds = [E.Round([str(i) for i in range(len(x[0]))], x[0], ORC_FRONT if x[1] == oracle.FRONT else ORC_BACK,
                          x[2], x[3], True, True) for x in spec]
        with E.Engine(rounds, max_reads=rs.n_reads, max_bytes=int(rs.seq.shape[0]) + 64,
                      max_name_bytes=int(rs.names.shape[0]) + 64, n_slots=1, emit_fastq=True, want_matches=True) as eng:
            res = eng.run(rs)
            rec0, rec1, oseq, oqual, olen = H.run_oracle(spec, rs)
            assert H.diff_matches(rec0, res.matches[0])[1] == 0, (trial, "round 1")
            assert H.diff_matches(rec1, res.matches[1])[1] == 0, (trial, "round 2")
            assert np.array_equal(res.out_len, olen)
            exp = _expected_fastq(rs, rec0, rec1, oseq, oqual, olen, eng.n_bins, eng.bin_id)
            for bb in range(eng.n_bins):
                assert res.bin_bytes(bb) == exp[bb], "trial %d bin %d bytes differ" % (trial, bb)
            if trial == 0:
                with_u = synth.from_records([("u", "ACGUACGTTGCA" * 5, "I" * 60), ("v", "ACGTACGT", "I" * 8)])
                with pytest.raises(E.OrcError, match="U"):
                    eng.run(with_u)
                again = eng.run(rs)                     # the slot is usable afterwards
                assert np.array_equal(again.bin, res.bin)


def test_adapters_over_64_nt_gpu():
    """SURVEY 8f N4 "adapters > 64 nt": a round that holds one runs long_kernel (cutadapt's recurrence cell by
    cell) instead of the bit-parallel scan.  Matches, trimmed lengths and the FASTQ bytes of every bin against
    the oracle: both rounds long, a long round beside the M13 round, plain and IUPAC adapters, with and without
    indels, --rc on and off."""
    import random
    import oracle
    import test_hostsim as TH
    from orcdemux import m13
    from orcdemux.lib import ORC_BACK, ORC_FRONT
    rnd = random.Random(6402)
    hits = 0
    for trial in range(8):
        wild = trial % 3 == 2
        f, b = TH._long_sets(rnd, wild)
        e = rnd.choice([0.0, 0.1, 0.1, 0.2, 0.3])
        ov = rnd.choice([1, 3, 3, 10, 70])
        rc = rnd.choice([0, 1, 1])
        indels = trial % 4 != 3
        if trial % 5 == 1 and not wild:
            b = [s for _, s in m13.sp27_reverse_rc()]
        if trial % 5 == 4 and not wild:
            f = [s for _, s in m13.sp5_forward()]
        plain = lambda x: "".join(c if c in "ACGT" else rnd.choice("ACGT") for c in x)
        rs = TH._adversarial_reads(rnd, [plain(x) for x in f], [plain(x) for x in b], 1200)
        spec = [(f, oracle.FRONT, e, ov, rc), (b, oracle.BACK, e, ov, rc)]
        rounds = [E.Round([str(i) for i in range(len(x[0]))], x[0], ORC_FRONT if x[1] == oracle.FRONT else ORC_BACK,
                          x[2], x[3], indels, bool(x[4])) for x in spec]
        with E.Engine(rounds, max_reads=rs.n_reads, max_bytes=int(rs.seq.shape[0]) + 64,
                      max_name_bytes=int(rs.names.shape[0]) + 64, n_slots=1, emit_fastq=True, want_matches=True) as eng:
            res = eng.run(rs)
            rec0, rec1, oseq, oqual, olen = H.run_oracle(spec, rs, indels=indels)
            idx, nbad = H.diff_matches(rec0, res.matches[0])
            assert nbad == 0, (trial, "round 1", idx[:3])
            idx, nbad = H.diff_matches(rec1, res.matches[1])
            assert nbad == 0, (trial, "round 2", idx[:3])
            assert np.array_equal(res.out_len, olen)
            exp = _expected_fastq(rs, rec0, rec1, oseq, oqual, olen, eng.n_bins, eng.bin_id)
            for bb in range(eng.n_bins):
                assert res.bin_bytes(bb) == exp[bb], "trial %d bin %d bytes differ" % (trial, bb)
            t = eng.timings(0)
            assert t["cells"][0] > 0
        hits += int((rec0["adapter"] >= 0).sum()) + int((rec1["adapter"] >= 0).sum())
    assert hits > 5000
    with pytest.raises(E.OrcError, match="unsupported"):
        E.Engine([E.Round(["a"], ["A" * 257], ORC_BACK, 0.1, 3, True, True)], max_reads=16, max_bytes=1024)


def test_seeded_stage1_random_adapter_sets():
    """Adapter sets stage 1 can seed (long adapters, long shared prefix, low error rates; pieces per
    adapter minus errors 1 or 2) through the kernels, with and without the seed table."""
    import random
    import oracle
    import test_hostsim as TH
    from orcdemux.lib import ORC_BACK, ORC_FRONT
    rnd = random.Random(777)
    for trial in range(8):
        f, b, e, ov, rc = TH._seed_friendly_sets(rnd)
        spec = [(f, oracle.FRONT, e, ov, rc), (b, oracle.BACK, e, ov, rc)]
        if trial % 3 == 2:
            spec = spec[::-1]
        rs = TH._adversarial_reads(rnd, f, b, 3000)
        rounds = [E.Round([str(i) for i in range(len(x[0]))], x[0], ORC_FRONT if x[1] == oracle.FRONT else ORC_BACK,
                          x[2], x[3], True, bool(x[4])) for x in spec]
        rec = H.run_oracle(spec, rs)
        seen = {}
        for no_seed in ("0", "1"):
            os.environ["ORC_NO_SEED"] = no_seed
            try:
                with E.Engine(rounds, max_reads=rs.n_reads, max_bytes=int(rs.seq.shape[0]) + 64, n_slots=1) as eng:
                    res = eng.run(rs)
                    launches = eng.timings(0)["kernel_launches"]
            finally:
                os.environ.pop("ORC_NO_SEED", None)
            assert H.diff_matches(rec[0], res.matches[0])[1] == 0, (trial, no_seed, "round 1")
            assert H.diff_matches(rec[1], res.matches[1])[1] == 0, (trial, no_seed, "round 2")
            assert np.array_equal(res.out_len, rec[4])
            seen[no_seed] = launches
        assert 0 <= seen["0"] - seen["1"] <= 2          # one seed_kernel per round whose table could be built


def test_long_reads_and_capacity_errors():
    from orcdemux.engine import OrcError
    sp5 = [s for _, s in synth.m13.sp5_forward()]
    sp27 = [s for _, s in synth.m13.sp27_reverse_rc()]
    import random
    rnd = random.Random(1)
    body = "".join(rnd.choice("ACGT") for _ in range(150000))
    recs = [("long_fwd", sp5[3] + body + sp27[4], "I" * (59 + 150000 + 57)),
            ("long_rc", synth.m13.revcomp(sp5[3] + body + sp27[4]), "I" * (59 + 150000 + 57)),
            ("short", sp5[0] + "ACGT" * 30 + sp27[1], "I" * (59 + 120 + 57))]
    rs = synth.from_records(recs)
    res = _check(rs)
    assert list(res.matches[0]["adapter"]) == [3, 3, 0] and list(res.matches[1]["adapter"]) == [4, 4, 1]
    eng = _engine(2, 1024)
    try:
        with pytest.raises(OrcError, match="exceeds"):
            eng.run(rs)
    finally:
        eng.close()


def test_no_indels_unanchored_gpu():
    rs = synth.generate(20000, 300, 900, seed=21)
    rounds = E.m13_rounds()
    for r in rounds:
        r.indels = False
    with E.Engine(rounds, max_reads=rs.n_reads, max_bytes=int(rs.seq.shape[0]), n_slots=1) as eng:
        res = eng.run(rs)
        t = eng.timings(0)
    rec0, rec1, oseq, oqual, olen = H.run_oracle(H.m13_rounds(), rs, indels=False)
    assert H.diff_matches(rec0, res.matches[0])[1] == 0
    assert H.diff_matches(rec1, res.matches[1])[1] == 0
    assert np.array_equal(res.out_len, olen)
    assert t["n_tasks"] == [0, 0]


def test_device_generated_shards_config5():
    """BASELINE configs[4]'s input path: shards made on the GPU by orc_synth (seed = (1005 << 32) + shard) are
    deterministic, distinct, follow the read model, and demultiplex exactly like the oracle says for the
    very bytes orc_export hands back."""
    n = (1 << 17) if FULL else (1 << 14)
    eng = E.Engine(E.m13_rounds(), max_reads=n, max_bytes=n * 980, max_name_bytes=24 * n, n_slots=2,
                   emit_fastq=True, want_matches=True)
    try:
        eng.synth(0, (1005 << 32) + 3, n, 300, 900)
        eng.synth(1, (1005 << 32) + 4, n, 300, 900)
        a, b = eng.export(0), eng.export(1)
        assert a.n_reads == n and b.n_reads == n
        assert a.seq.shape[0] != b.seq.shape[0] or not np.array_equal(a.seq, b.seq)
        L = a.lengths.astype(np.int64)
        assert L.min() >= 150 and L.max() <= 1000 and 560 < L.mean() < 640
        assert np.array_equal(a.offsets[1:], np.cumsum(L[:-1]).astype(np.uint64)) and int(a.offsets[0]) == 0
        assert set(np.unique(a.seq).tolist()) <= set(b"ACGTN") and a.qual.min() >= 33 + 5 and a.qual.max() <= 33 + 40
        assert a.read(0)[0] == "r0 ch=0" and a.read(1)[0] == "r1" and a.read(7)[0] == "r7 ch=7"
        eng.launch(0)
        eng.download(0)
        res = eng.wait(0)
        rec0, rec1, oseq, oqual, olen = H.run_oracle(H.m13_rounds(), a, n_threads=NCPU)
        idx, nbad = H.diff_matches(rec0, res.matches[0])
        assert nbad == 0, ("round 1", nbad, idx)
        idx, nbad = H.diff_matches(rec1, res.matches[1])
        assert nbad == 0, ("round 2", nbad, idx)
        assert np.array_equal(res.out_len, olen)
        exp = _expected_fastq(a, rec0, rec1, oseq, oqual, olen, 169, lambda x, y: (x + 1) + 13 * (y + 1))
        for bb in range(169):
            assert res.bin_bytes(bb) == exp[bb], "bin %d bytes differ" % bb
        # the read model: most reads carry both adapters, about a tenth are reverse-complemented
        assert (rec0["adapter"] >= 0).mean() > 0.85 and 0.05 < rec0["is_rc"].mean() < 0.15
        valid = (rec0["adapter"] >= 0) & (rec1["adapter"] >= 0)
        assert valid.mean() > 0.75 and len(np.unique(rec0["adapter"][valid])) == 12 and len(np.unique(rec1["adapter"][valid])) == 12
        eng.synth(1, (1005 << 32) + 3, n, 300, 900)         # same seed: the same bytes
        c = eng.export(1)
        assert np.array_equal(a.seq, c.seq) and np.array_equal(a.qual, c.qual) and np.array_equal(a.names, c.names)
        with pytest.raises(E.OrcError, match="max_reads"):
            eng.synth(0, 1, n + 1, 300, 900)
    finally:
        eng.close()
