"""cutadapt-shaped report of the CLI: per-adapter end statistics derived from
the match records -- removed-length histograms by error count, matches on the reverse
complement, and for 3' adapters the base preceding the match (adapters.py
RemoveAfterMatch.adjacent_base: read[rstart - 1] in the orientation that matched)."""
import numpy as np

from orcdemux import cli, lib
from orcdemux import fastq as F


def _batch(records):
    text = np.frombuffer(b"".join(b"@%s\n%s\n+\n%s\n" % (n, s, b"I" * len(s)) for n, s in records), dtype=np.uint8)
    n, used, arr = F.index_text(text.copy(), len(text), len(records), True)
    return F.TextBatch(text, used, n, *arr)


def test_three_prime_statistics_and_text():
    names, seqs = ["a1", "a2"], ["ACGTACGTAC", "GGGGGCCCCC"]
    st = cli.EndStats(names, seqs, lib.ORC_BACK, 0.1, True, True)
    #            forward, base before = A   matched on the rc (AAAACGTACGTAC)   adapter at offset 0
    tb = _batch([(b"r0", b"AAAACGTACGTAC"), (b"r1", b"GTACGTACGTTTT"), (b"r2", b"ACGTACGTAC"),
                 (b"r3", b"ccnACGTACGTAC"), (b"r4", b"TTTTTTTT")])
    m = np.zeros(5, dtype=lib.MATCH_DTYPE)
    m["adapter"] = [0, 0, 0, 0, -1]
    m["is_rc"] = [0, 1, 0, 0, 0]
    m["query_start"] = [3, 3, 0, 3, 0]
    m["query_stop"] = [13, 13, 10, 13, 0]
    m["errors"] = [0, 0, 0, 1, 0]
    st.add(m, tb.lengths[:5], tb)
    end = st.as_json(5)[0]["three_prime_end"]
    assert end["adjacent_bases"] == {"A": 2, "C": 0, "G": 0, "T": 0, "": 2}       # 'n' and "no base" are none/other
    assert end["type"] == "regular_three_prime" and end["matches"] == 4 and end["error_lengths"] == [9, 10]
    assert end["trimmed_lengths"] == [{"len": 10, "expect": 5 * 0.25 ** 10, "counts": [3, 1]}]
    assert end["dominant_adjacent_base"] is None                                 # fewer than 20 matches
    js = st.as_json(5)
    assert js[0]["on_reverse_complement"] == 1 and js[0]["five_prime_end"] is None and js[1]["total_matches"] == 0
    txt = st.text(5, 3)
    assert "=== Adapter a1 ===" in txt and "Type: regular 3'; Length: 10; Trimmed: 4 times; Reverse-complemented: 1 times" in txt
    assert "1-9 bp: 0; 10 bp: 1" in txt and "  A: 50.0%" in txt and "  none/other: 50.0%" in txt
    assert "10\t4\t0.0\t1\t3 1" in txt
    assert txt.count("Overview of removed sequences") == 1                       # a2 never matched


def test_dominant_adjacent_base_and_five_prime():
    st = cli.EndStats(["x"], ["ACGTACGTAC"], lib.ORC_BACK, 0.1, True, False)
    tb = _batch([(b"r%d" % i, b"TTGACGTACGTAC") for i in range(25)])
    m = np.zeros(25, dtype=lib.MATCH_DTYPE)
    m["query_start"], m["query_stop"] = 3, 13
    st.add(m, tb.lengths[:25], tb)
    j = st.as_json(25)[0]
    assert j["three_prime_end"]["dominant_adjacent_base"] == "G" and j["on_reverse_complement"] is None
    assert "preceded by 'G' extremely often" in st.text(25, 3)
    f = cli.EndStats(["x"], ["ACGTACGTAC"], lib.ORC_FRONT, 0.1, True, True)
    m["query_start"], m["query_stop"] = 0, 7
    f.add(m, tb.lengths[:25], tb)
    e = f.as_json(25)[0]["five_prime_end"]
    assert e["adjacent_bases"] is None and e["trimmed_lengths"][0]["len"] == 7
    assert e["trimmed_lengths"][0]["expect"] == 25 * 0.25 ** 7
