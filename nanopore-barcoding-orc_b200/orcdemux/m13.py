"""The OrCA-seq M13 index tables used by the two-round demultiplex.

The reference keeps them as FASTA under adapters_primers/ (used at
/root/reference/scripts/02_cutadapt_loop.sh:43-44).  Every SP5 (5') adapter is
`SP5_LEFT + index + M13F`, every SP27 (3', as seen on the oriented read) adapter is
`M13R_RC + revcomp(index) + SP27_RIGHT`; only the 17-nt variable indices differ, so the
tables are stored here as flanks + indices and expanded on demand.
"""
from __future__ import annotations

import os

SP5_LEFT = "CATGTAATGCACGTACTTTCAGGGT"      # 25 nt constant 5' flank
M13F = "TGTAAAACGACGGCCAG"                   # 17 nt M13 forward
M13R_RC = "GTCATAGCTGTTTCCTG"                # 17 nt, "G" + revcomp of M13 reverse
SP27_RIGHT = "AGTCGTCGCAGCCTCACCTGATC"        # 23 nt constant 3' flank

# 17-nt variable indices, forward orientation, plate order 001..012
SP5_INDEX = """GAGCGTCTAATCGTAAT CTACCGTGGATATTCAA AATTCCACTTACAACGG AGTGTGCCGCCAACCAA
AGCCTCATTGGTTGTTC GATTCTACAAGTGGTGA ACAGGTTGCCGGAGTCT CAATCGTGACCATCCGG
AACAACAACAACAACCG GGTCAGGTAGTCCGTAT GCCTGTGCGGAGTAGAT CCAACGGACTACGAATT""".split()
SP27_INDEX = """CCTCCGTGCCTGGTTAA AACTTCAGGTCCACAGC ACGCGGTGGTGTAACGA AAGAATGGATAAGGAGG
ATAGGTCATTGCGCTTC CCGATCCTTCAGAGCCA CGCTGCTAGAATATGCC ATTGGACTGTTAGGAGG
CGGTACATCGCTCCTTA ATTGTAGCTTCTCCTTC CACCTAAGCGACACGTT GTTGTTCACGATACTAC""".split()

_COMP = str.maketrans("ACGT", "TGCA")


def revcomp(s: str) -> str:
    return s.translate(_COMP)[::-1]


def sp5_forward():
    """[(name, 59-nt sequence)] == M13_amplicon_indices_forward.fa (round 1, -g file:)."""
    return [("SP5_%03d" % (i + 1), SP5_LEFT + v + M13F) for i, v in enumerate(SP5_INDEX)]


def sp27_reverse_rc():
    """[(name, 57-nt sequence)] == M13_amplicon_indices_reverse_rc.fa (round 2, -a file:)."""
    return [("SP27_%03d" % (i + 1), M13R_RC + revcomp(v) + SP27_RIGHT) for i, v in enumerate(SP27_INDEX)]


def variable_all():
    """[(name, 17-nt index)] == M13_variable_indices_all.fa (BASELINE config 4)."""
    return [("SP5_%03d" % (i + 1), v) for i, v in enumerate(SP5_INDEX)] + \
           [("SP27_%03d" % (i + 1), v) for i, v in enumerate(SP27_INDEX)]


def write_fasta(path: str, records) -> str:
    """Write records as the reference lays them out (no trailing newline after the last)."""
    with open(path, "w") as fh:
        fh.write("\n".join(">%s\n%s" % (n, s) for n, s in records))
    return path


def write_tables(directory: str):
    """Materialise the three FASTA files; returns their paths (forward, reverse_rc, variable)."""
    os.makedirs(directory, exist_ok=True)
    return (write_fasta(os.path.join(directory, "M13_amplicon_indices_forward.fa"), sp5_forward()),
            write_fasta(os.path.join(directory, "M13_amplicon_indices_reverse_rc.fa"), sp27_reverse_rc()),
            write_fasta(os.path.join(directory, "M13_variable_indices_all.fa"), variable_all()))
