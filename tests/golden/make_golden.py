"""Writes tests/golden/kat.json.

There is no importable reference (the arithmetic is cutadapt 4.9, un-vendored, SURVEY.md
8c), so these are NOT outputs of the reference: they are (a) the cutadapt user-guide examples
for regular 3'/5' adapters, (b) one known-answer test recalled from cutadapt's own
tests/test_align.py (indels are penalised: (0, 10, 0, 10, 8, 1)), (c) vectors derived by
hand from the published algorithm; the expected values were written by hand, the script only
lays them out.  tests/test_oracle.py checks all three oracle restatements against them."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "nanopore-barcoding-orc_b200"))
from orcdemux import m13  # noqa: E402

FRONT, BACK, PREFIX = 11, 14, 8
loc = []


def L(ref, query, rate, flags, mo, expect, note, indel_cost=1):
    loc.append(dict(ref=ref, query=query, rate=rate, flags=flags, min_overlap=mo, indel_cost=indel_cost,
                    expect=expect, note=note))


# (a) user guide, regular 3' adapter
L("ADAPTER", "MYSEQUENCEADAPTER", 0.1, BACK, 3, [0, 7, 10, 17, 7, 0], "guide: full 3' adapter")
L("ADAPTER", "MYSEQUENCEADAP", 0.1, BACK, 3, [0, 4, 10, 14, 4, 0], "guide: partial 3' adapter")
L("ADAPTER", "MYSEQUENCEADAPTERSOMETHINGELSE", 0.1, BACK, 3, [0, 7, 10, 17, 7, 0], "guide: adapter inside")
L("ADAPTER", "MYSEQUENCEAD", 0.1, BACK, 3, None, "guide: overlap 2 < -O 3")
L("ADAPTER", "MYSEQUENCE", 0.1, BACK, 3, None, "guide: no adapter")
# (a) user guide, regular 5' adapter
L("ADAPTER", "ADAPTERMYSEQUENCE", 0.1, FRONT, 3, [0, 7, 0, 7, 7, 0], "guide: full 5' adapter")
L("ADAPTER", "DAPTERMYSEQUENCE", 0.1, FRONT, 3, [1, 7, 0, 6, 6, 0], "guide: partial 5' adapter")
L("ADAPTER", "TERMYSEQUENCE", 0.1, FRONT, 3, [4, 7, 0, 3, 3, 0], "guide: partial 5' adapter")
L("ADAPTER", "SOMETHINGADAPTERMYSEQUENCE", 0.1, FRONT, 3, [0, 7, 9, 16, 7, 0], "guide: adapter inside")
L("ADAPTER", "ERMYSEQUENCE", 0.1, FRONT, 3, None, "overlap 2 < -O 3")
# (b) cutadapt tests/test_align.py TestAligner.test_indels_penalized
L("CCAGTCCTCT", "CCAGTCCTTTCCTGAGAGT", 0.3, PREFIX, 1, [0, 10, 0, 10, 8, 1], "upstream KAT: mismatch beats 3 deletions")
# (c) hand-derived
L("ACGT", "", 0.1, BACK, 3, None, "empty read")
L("ACGT", "ACG", 0.1, BACK, 3, [0, 3, 0, 3, 3, 0], "read is an adapter prefix")
L("ACGT", "CGT", 0.1, FRONT, 3, [1, 4, 0, 3, 3, 0], "read is an adapter suffix")
L("ACGTACGTAC", "TTTTACGTACGTACTTTT", 0.1, BACK, 3, [0, 10, 4, 14, 10, 0], "exact, early exit")
L("ACGTACGTAC", "TTTTACGTTCGTACTTTT", 0.1, BACK, 3, [0, 10, 4, 14, 8, 1], "1 substitution in 10: score 9-1")
L("ACGTACGTAC", "TTTTACGTTCGTACTTTT", 0.09, BACK, 3, None, "1 error needs rate*10 >= 1")
L("ACGTACGTAC", "TTTTACGTCGTACTTTT", 0.1, BACK, 3, [0, 10, 4, 13, 7, 1], "1 deletion in the read: 9 matches - 2")
L("ACGTACGTAC", "TTTTACGTAACGTACTTTT", 0.1, BACK, 3, [0, 10, 4, 15, 8, 1], "1 insertion in the read: 10 matches - 2")
L("AAAAAAAAAA", "CCCCAAAAAAAAAAAAAAACCCC", 0.1, BACK, 3, [0, 10, 4, 14, 10, 0], "leftmost occurrence in a homopolymer")
L("AAAAAAAAAA", "CCCCAAAAAAAAAAAAAAACCCC", 0.1, FRONT, 3, [0, 10, 4, 14, 10, 0], "leftmost occurrence, 5' adapter")
L("ACGTACGTAC", "GGGGGGGGACGTAC", 0.1, BACK, 3, [0, 6, 8, 14, 6, 0], "partial at the 3' end via the last column")
L("ACGTACGTAC", "GTACGGGGGGGG", 0.1, FRONT, 3, [6, 10, 0, 4, 4, 0], "partial at the 5' end via column 0")
L("TGTAAAACGACGGCCAG", "CAG" + "T" * 40, 0.1, FRONT, 3, [14, 17, 0, 3, 3, 0], "3-nt suffix overlap is a match at -O 3")
L("GTCATAGCTGTTTCCTG", "A" * 40 + "GTC", 0.1, BACK, 3, [0, 3, 40, 43, 3, 0], "3-nt prefix overlap at the read end")
L("ACGTACGTAC", "ACGTACGTAC", 0.0, FRONT, 3, [0, 10, 0, 10, 10, 0], "whole read is the adapter")
L("ACGTACGTAC", "ACGTACGTAC", 0.0, BACK, 3, [0, 10, 0, 10, 10, 0], "whole read is the adapter (3')")
L("ACGTACGTACGTACGTACGT", "TTACGTACGTACGTACGTACGTTT", 0.1, BACK, 3, [0, 20, 2, 22, 20, 0], "20-mer exact")
L("ACGTACGTACGTACGTACGT", "TTACGTACGAACGTACGAACGTTT", 0.1, BACK, 3, [0, 20, 2, 22, 16, 2], "2 substitutions in 20")
L("ACGTACGTACGTACGTACGT", "TTACGAACGAACGTACGAACGTTT", 0.1, BACK, 3, None, "3 substitutions in 20 exceed 0.1")
L("ACGT", "ACGT", 0.1, FRONT, 5, [0, 4, 0, 4, 4, 0], "min_overlap is clamped by the caller, 5 > m: still matches only if clamped",)
loc[-1]["min_overlap"] = 4
L("ACGTACGTAC", "TTTTACGTTCGTACTTTT", 0.1, BACK, 3, [0, 10, 4, 14, 8, 1], "no-indels keeps substitutions", 100000)
L("ACGTACGTAC", "TTTTACGTCGTACTTTT", 0.1, BACK, 3, [4, 10, 0, 0, 0, 0], "placeholder", 100000)
loc.pop()   # the no-indel deletion case is covered by the random cross-checks instead

sp5 = dict(m13.sp5_forward())
sp27 = dict(m13.sp27_reverse_rc())
ins = "ACGGTCTATCGGATTCAGCATCGATCGGATATTTCAGCGACTACGACTACGGGACTACTATCGAGGACTTTACGACGATCAGCGACTACTAGCATCATC" * 3
reads = []
full = sp5["SP5_007"] + ins + sp27["SP27_003"]
reads.append(dict(name="exact_pair", seq=full, sp5="SP5_007", sp27="SP27_003",
                  r1=[0, 0, 59, 0, 59, 59, 0], r2=[0, 0, 57, len(ins), len(ins) + 57, 57, 0], trimmed=ins))
reads.append(dict(name="exact_pair_revcomp", seq=m13.revcomp(full), sp5="SP5_007", sp27="SP27_003",
                  r1=[1, 0, 59, 0, 59, 59, 0], r2=[0, 0, 57, len(ins), len(ins) + 57, 57, 0], trimmed=ins))
reads.append(dict(name="no_adapters", seq=ins, sp5="unknown", sp27="unknown", trimmed=ins))
tail = "T" * 300
reads.append(dict(name="tie_to_first_cag", seq="CAG" + tail, sp5="SP5_001", sp27="unknown",
                  r1=[0, 56, 59, 0, 3, 3, 0], trimmed=tail))
reads.append(dict(name="sp5_only", seq=sp5["SP5_012"] + ins, sp5="SP5_012", sp27="unknown",
                  r1=[0, 0, 59, 0, 59, 59, 0], trimmed=ins))
body = "A" * 200
reads.append(dict(name="ends_gtc_ties_to_sp27_001", seq=sp5["SP5_002"] + body + "GTC", sp5="SP5_002", sp27="SP27_001",
                  r1=[0, 0, 59, 0, 59, 59, 0], r2=[0, 0, 3, 200, 203, 3, 0], trimmed=body))
a4 = sp5["SP5_004"]
flank_sub = list(a4)
for p in (2, 9, 16, 45, 52):      # 5 substitutions, all in the constant flanks
    flank_sub[p] = "A" if flank_sub[p] != "A" else "C"
# (d) adapters with IUPAC wildcards (wildcard_ref): recalled from cutadapt's tests/test_align.py
# test_n_wildcards_not_counted_aligner_back / _front -- the 14 N do not count, so the 20-mer allows
# int(0.1 * 6) = 0 errors -- plus hand-derived mask cases
wild = []


def LW(ref, query, rate, flags, mo, expect, note):
    wild.append(dict(ref=ref, query=query, rate=rate, flags=flags, min_overlap=mo, expect=expect, note=note))


NREF = "AGGNNNNNNNNNNNNNNTTC"
LW(NREF, "TTC", 0.1, BACK, 3, None, "upstream KAT (back): adapter end is not a 3' overlap")
LW(NREF, "AGG", 0.1, BACK, 3, [0, 3, 0, 3, 3, 0], "upstream KAT (back)")
LW(NREF, "AGGCCCCCCC", 0.1, BACK, 3, [0, 10, 0, 10, 10, 0], "upstream KAT (back): N match anything")
LW(NREF, "ATGCCCCCCC", 0.1, BACK, 3, None, "upstream KAT (back): 1 error in 3 effective characters")
LW(NREF, "AGGCCCCCCCCCCCCCCATC", 0.1, BACK, 3, None, "upstream KAT (back): 1 error in 6 effective characters")
LW(NREF, "CCC" + NREF.replace("N", "C") + "AAA", 0.1, BACK, 3, [0, 20, 3, 23, 20, 0], "upstream KAT (back): full")
LW(NREF, "TTC", 0.1, FRONT, 3, [17, 20, 0, 3, 3, 0], "upstream KAT (front)")
LW(NREF, "TGC", 0.1, FRONT, 3, None, "upstream KAT (front)")
LW(NREF, "CCCCCCCTTC", 0.1, FRONT, 3, [10, 20, 0, 10, 10, 0], "upstream KAT (front)")
LW(NREF, "CCCCCCCGTC", 0.1, FRONT, 3, None, "upstream KAT (front)")
LW(NREF, "CCC" + NREF.replace("N", "C") + "AAA", 0.1, FRONT, 3, [0, 20, 3, 23, 20, 0], "upstream KAT (front): full")
LW("ACRYACGT", "TTACGTACGTTT", 0.0, BACK, 3, [0, 8, 2, 10, 8, 0], "R = A|G, Y = C|T")
LW("ACRYACGT", "TTACCTACGTTT", 0.0, BACK, 3, None, "C is not in R")
LW("ACGTNCGT", "TTACGTNCGTTT", 0.0, BACK, 3, None, "a read N matches nothing, not even an adapter N")
LW("ACGUACGT", "TTACGTACGUTT", 0.0, BACK, 3, [0, 8, 2, 10, 8, 0], "U is T on both sides")
LW("ACGTXCGT", "TTACGTACGTTT", 0.2, BACK, 3, [0, 8, 2, 10, 6, 1], "X matches nothing: always one error")

reads.append(dict(name="five_flank_substitutions", seq="".join(flank_sub) + ins, sp5="SP5_004", sp27="unknown",
                  r1=[0, 0, 59, 0, 59, 49, 5], trimmed=ins))
reads.append(dict(name="truncated_5prime", seq=sp5["SP5_009"][20:] + ins + sp27["SP27_008"], sp5="SP5_009", sp27="SP27_008",
                  r1=[0, 20, 59, 0, 39, 39, 0], r2=[0, 0, 57, len(ins), len(ins) + 57, 57, 0], trimmed=ins))
reads.append(dict(name="truncated_3prime", seq=sp5["SP5_009"] + ins + sp27["SP27_008"][:40], sp5="SP5_009", sp27="SP27_008",
                  r1=[0, 0, 59, 0, 59, 59, 0], r2=[0, 0, 40, len(ins), len(ins) + 40, 40, 0], trimmed=ins))

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json")
with open(out, "w") as fh:
    json.dump(dict(source="hand-written expectations; see make_golden.py", locate=loc, locate_wildcard=wild, reads=reads), fh, indent=1)
print("wrote", out, len(loc), "locate vectors,", len(reads), "read vectors")
