"""orcdemux -- host side of the B200-native two-round SP5xSP27 demultiplexer.

Drop-in for the two `cutadapt` call shapes of
/root/reference/scripts/02_cutadapt_loop.sh:64-72 and :94-102.  All matching, trimming
and binning runs in liborcdemux.so (hand-written sm_100a CUDA behind the C ABI of
include/orcdemux.h); this package only parses arguments and FASTQ, owns pinned buffers
and writes files.  There is no CPU fallback: importing `orcdemux.lib` without the built
library, or creating an engine without a CUDA device, raises.
"""
__version__ = "0.1.0"
