/*
 * orcdemux.h -- C ABI of liborcdemux.so, the B200-native drop-in for the adapter/index
 * matching, trimming and binning that the reference delegates to two `cutadapt`
 * invocations per dataset:
 *
 *   round 1  /root/reference/scripts/02_cutadapt_loop.sh:64-72
 *            cutadapt --action=trim -e 0.1 -j 24 --rc -g file:M13_amplicon_indices_forward.fa
 *                     -o SP5/{name}_<ds>.fastq.gz IN --json=...
 *   round 2  /root/reference/scripts/02_cutadapt_loop.sh:94-102   (once per SP5 bin)
 *            cutadapt --action=trim -e 0.1 -j 24 --rc -a file:M13_amplicon_indices_reverse_rc.fa
 *                     -o SP27/{name}_<SP5id>_<ds>.fastq.gz SP5/<SP5id>_<ds>.fastq.gz --json=...
 *
 * The reference has no in-process FFI for this path (the boundary is the cutadapt
 * command line); this header is what a host binding (ctypes/cgo/JNI) of a
 * cutadapt-compatible front end binds instead.  Plain pointers and sizes only.
 *
 * Threading: one ctx per GPU; a ctx is not thread-safe; different ctxs are independent.
 * Errors: functions return 0 on success and a negative ORC_E* code on failure;
 * orc_last_error() gives the text.  There is NO CPU fallback: without a usable CUDA
 * device orc_create() fails.
 */
#ifndef ORCDEMUX_H
#define ORCDEMUX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_ROUNDS 2
#define ORC_MAX_ADAPTERS 32      /* unanchored adapters per round (x2 orientations = 64 lanes) */
#define ORC_MAX_ANCHORED 64      /* anchored no-indel adapters per round */
#define ORC_MAX_ADAPTER_LEN 64   /* one 64-bit Myers word: the bit-parallel path */
#define ORC_MAX_LONG_ADAPTER_LEN 256 /* a 5' / 3' round that holds an adapter over 64 nt runs cutadapt's recurrence cell by
                                        cell on the GPU (long_kernel): exact, far slower, at most ORC_MAX_LONG_ADAPTERS */
#define ORC_MAX_LONG_ADAPTERS 16

typedef struct orc_ctx orc_ctx;   /* opaque; one per GPU */

/* adapter types == cutadapt's -g / -a / -g ^ / -a ...$ (adapters.py Front/Back/Prefix/SuffixAdapter) */
enum { ORC_FRONT = 0, ORC_BACK = 1, ORC_PREFIX = 2, ORC_SUFFIX = 3 };

enum { ORC_ACTION_TRIM = 0, ORC_ACTION_RETAIN = 1 };

enum {
    ORC_OK = 0,
    ORC_EINVAL = -1,       /* bad argument / unsupported option */
    ORC_ECUDA = -2,        /* CUDA runtime error (text in orc_last_error) */
    ORC_ECAPACITY = -3,    /* batch exceeds the capacities given to orc_create */
    ORC_ESTATE = -4        /* call sequence error (e.g. wait on an idle slot) */
};

/* One cutadapt invocation's matching options (replaces the argv at 02:64-72 / 02:94-102). */
typedef struct orc_round_params {
    int32_t n_adapters;             /* records of the `file:` FASTA, file order kept */
    int32_t type;                   /* ORC_FRONT (-g), ORC_BACK (-a); ORC_PREFIX (-g ^) / ORC_SUFFIX (-a ...$)
                                       only with indels == 0 (Hamming fast path, up to 64 adapters) */
    const char *const *names;       /* [n_adapters] header.split()[0]; may be NULL */
    const char *const *sequences;   /* [n_adapters] NUL-terminated, <= ORC_MAX_ADAPTER_LEN (ORC_MAX_LONG_ADAPTER_LEN, see there); ACGT, or IUPAC codes (cutadapt's
                                       adapter wildcards).  Plain and IUPAC adapters may stand side by side: all are then
                                       compared through the IUPAC masks, which equals cutadapt's per-adapter choice except
                                       for a read with U -- orc_wait() refuses a batch that holds one (ORC_EINVAL) */
    double max_error_rate;          /* -e  (values >= 1 are absolute error counts, as in cutadapt) */
    int32_t min_overlap;            /* -O  (cutadapt default 3) */
    int32_t indels;                 /* 1; 0 == --no-indels */
    int32_t revcomp;                /* --rc */
    int32_t action;                 /* ORC_ACTION_TRIM (--action=trim, 0) or ORC_ACTION_RETAIN (--action=retain: the
                                       read is cut at the far side of the match, the adapter itself stays) */
} orc_round_params;

typedef struct orc_params {
    int32_t device;                 /* CUDA device ordinal */
    int32_t n_rounds;               /* 1 = one cutadapt call; 2 = round 1 then round 2 on assigned reads */
    orc_round_params rounds[ORC_MAX_ROUNDS];
    uint32_t max_reads;             /* capacity of one batch */
    uint64_t max_bytes;             /* capacity of the seq (== qual) blob of one batch, bytes */
    uint64_t max_name_bytes;        /* capacity of the names blob of one batch */
    int32_t n_slots;                /* batches that may be in flight (>= 1) */
    int32_t emit_fastq;             /* build bin-major FASTQ text on the device */
    int32_t want_matches;           /* return the per-read match records */
    const uint8_t *drop_bins;       /* optional [n_bins]: 1 = do not emit that bin (02:110-118) */
    int32_t qual_zero_copy;         /* 1 = with emit_fastq and the separate-blob layout, the qualities are NOT copied to
                                       the device: emit_kernel reads the bytes it needs (the trimmed qualities of the reads
                                       whose bin is kept) straight from the caller's buffer over PCIe.  The buffer must be
                                       page-locked (orc_host_alloc, cudaHostAlloc/cudaHostRegister, torch pin_memory) and
                                       its allocation must extend 64 bytes beyond qual + n_bytes (16-byte loads over-read);
                                       a batch whose qual buffer is not page-locked is copied as before */
    int32_t emit_gzip;              /* 1 = with emit_fastq, every bin of a batch comes back as ONE gzip member (what
                                       02:64-72 / 02:94-102 leave on disk are .fastq.gz files): orc_result.fastq then holds
                                       the members back to back, bin_offsets their byte ranges (an empty bin has an empty
                                       range), fastq_bytes their total.  A member is a single dynamic-Huffman DEFLATE block
                                       of literals, coded on the device (csrc/orc_gz.cuh), with the "OC" size subfield
                                       orc_reader_* uses to inflate members in parallel; appending the members of successive
                                       batches to a bin's file gives a valid .fastq.gz.  Only the compressed bytes cross
                                       PCIe */
} orc_params;

/*
 * One batch of reads.  Bin ids: with one round bin = adapter + 1 (0 == "unknown");
 * with two rounds bin = (a1 + 1) + (n_adapters1 + 1) * (a2 + 1), a == -1 for unknown.
 * Buffers stay owned by the caller and must stay valid until orc_wait() returns; pinned
 * memory (orc_host_alloc or torch pin_memory) makes the copies asynchronous.
 *
 * Two layouts are accepted:
 *  - separate blobs: seq and qual of equal size, read r at offsets[r] in both;
 *  - raw FASTQ text: seq == qual == names == the text, offsets / qual_offsets / name_offsets /
 *    name_lengths point into it (orc_fastq_index() computes them); the text is uploaded once.
 */
typedef struct orc_batch {
    uint32_t n_reads;
    uint64_t n_bytes;               /* bytes of the seq blob (and of the qual blob) that are in use */
    const uint8_t *seq;             /* ASCII bases, read r at offsets[r] .. offsets[r]+lengths[r] */
    const uint8_t *qual;            /* ASCII qualities; may be the same pointer as seq */
    const uint64_t *offsets;        /* [n_reads] */
    const uint32_t *lengths;        /* [n_reads] */
    const uint64_t *qual_offsets;   /* [n_reads] start of the qualities inside qual; NULL = offsets */
    const uint8_t *names;           /* header lines without '@' and newline; NULL if !emit_fastq;
                                       may be the same pointer as seq */
    const uint64_t *name_offsets;   /* [n_reads + 1]; only [n_reads] are read when name_lengths is given */
    const uint32_t *name_lengths;   /* [n_reads] or NULL (name r ends where name r+1 starts) */
    uint64_t name_bytes;            /* bytes of the names blob; ignored when names == seq */
} orc_batch;

/* What Aligner.locate returns for the adapter that won (cutadapt _align.pyx), plus which. */
typedef struct orc_match {
    int32_t adapter;                /* index in file order, -1 = no match ("unknown") */
    int32_t is_rc;                  /* --rc chose the reverse complement (name gets " rc") */
    int32_t ref_start, ref_stop;    /* adapter interval */
    int32_t query_start, query_stop;/* read interval, in the chosen orientation */
    int32_t score, errors;
} orc_match;

/* Host-visible results of one batch; pointers are owned by the ctx and stay valid until
 * the next orc_submit()/orc_upload() on the same slot. */
typedef struct orc_result {
    uint32_t n_reads;
    int32_t n_bins;
    const orc_match *matches[ORC_MAX_ROUNDS]; /* [n_reads] per round; NULL unless want_matches */
    const int32_t *bin;             /* [n_reads] bin id, -1 = dropped */
    const uint32_t *out_len;        /* [n_reads] length of the trimmed read */
    const uint64_t *bin_counts;     /* [n_bins] reads of this batch per bin */
    const uint64_t *bin_offsets;    /* [n_bins + 1] byte ranges of the bins inside fastq */
    const uint8_t *fastq;           /* bin-major FASTQ text, input order kept inside a bin (orc_params.emit_gzip: that
                                       text as one gzip member per bin) */
    uint64_t fastq_bytes;
} orc_result;

/* kernels of one round, in launch order (orc_timings.kernel_ms) */
enum { ORC_K_SORT_READS = 0, ORC_K_SEED, ORC_K_TRIGGER, ORC_K_SORT_ITEMS, ORC_K_FILTER, ORC_K_SCAN,
       ORC_K_RESOLVE_BAND, ORC_K_RESOLVE_WIDE, ORC_K_SELECT, ORC_N_KERNELS };

/* Device time of the stages of the last orc_launch()/orc_submit() on a slot, from CUDA
 * events recorded on the ctx's own stream (milliseconds), and launch counts. */
typedef struct orc_timings {
    float pack_ms;                  /* ASCII -> 4-bit codes */
    float trigger_ms[ORC_MAX_ROUNDS]; /* scan stage 1: length sort + shared-prefix trigger scan (0 if off) */
    float scan_ms[ORC_MAX_ROUNDS];  /* scan stage 2: bit-parallel edit-distance scan of the windows */
    float resolve_ms[ORC_MAX_ROUNDS]; /* exact banded DP of the candidate pairs + selection */
    float bin_ms;                   /* per-bin counts, stable offsets */
    float emit_ms;                  /* trimmed FASTQ records into their bins */
    float total_ms;                 /* first kernel start .. last kernel end */
    float h2d_ms, d2h_ms;
    uint32_t kernel_launches;
    uint32_t n_tasks[ORC_MAX_ROUNDS]; /* candidate pairs that needed the resolver's exact walk */
    uint32_t n_candidates[ORC_MAX_ROUNDS]; /* pairs with any candidate cell (incl. those settled in the scan) */
    uint64_t cells[ORC_MAX_ROUNDS];   /* algorithmic DP cells: pairs * m * n   */
    uint64_t cells_executed[ORC_MAX_ROUNDS]; /* DP cells the two scan stages really updated */
    uint64_t pack_bytes, emit_bytes;  /* algorithmic bytes moved by pack / emit */
    /* per kernel of each round (ORC_K_*), device time between events recorded around every launch */
    float kernel_ms[ORC_MAX_ROUNDS][ORC_N_KERNELS];
    uint64_t window_columns[ORC_MAX_ROUNDS]; /* columns stage 2 looks at, summed over the (read, direction) items */
    uint64_t cells_2b[ORC_MAX_ROUNDS];       /* DP cells of stage 2b: m x window columns of the pairs that passed 2a */
    uint32_t n_pairs_2b[ORC_MAX_ROUNDS];     /* pairs that passed stage 2a */
    uint32_t n_tasks_wide[ORC_MAX_ROUNDS];   /* resolver tasks the band resolver could not take */
    /* when things happened on the device, milliseconds since orc_create(): H2D copies start, kernels start,
     * emit_kernel starts, emit_kernel ends, D2H copies end (0 where that part did not run) -- the timeline of a
     * pipelined orc_submit()/orc_wait() loop over several slots */
    float timeline_ms[5];
    float gzip_ms;                  /* emit_gzip: the bins' FASTQ text -> gzip members (0 if off) */
    uint64_t gzip_bytes;            /* emit_gzip: bytes of the members of this batch */
} orc_timings;

orc_ctx *orc_create(const orc_params *params, char *err, size_t err_len);
void orc_destroy(orc_ctx *ctx);
const char *orc_last_error(orc_ctx *ctx);
int orc_n_bins(orc_ctx *ctx);

/* whole step: H2D copies, kernels, D2H copies, all asynchronous on the slot's stream */
int orc_submit(orc_ctx *ctx, int slot, const orc_batch *batch);
int orc_wait(orc_ctx *ctx, int slot, orc_result *out);

/* the same step in three parts, for device-resident timing */
int orc_upload(orc_ctx *ctx, int slot, const orc_batch *batch);
int orc_launch(orc_ctx *ctx, int slot);
int orc_download(orc_ctx *ctx, int slot);
int orc_sync(orc_ctx *ctx, int slot);

int orc_get_timings(orc_ctx *ctx, int slot, orc_timings *out);
/* when the stages of the batch last waited for on a slot finished on the device, milliseconds since orc_create():
 * out6 = {stream reached the upload, H2D copies done, matching + binning kernels done, emit_kernel done, gzip stage
 * done, D2H copies done}.  Event queries only: usable inside a pipelined submit / wait loop, after orc_wait(). */
int orc_get_timeline(orc_ctx *ctx, int slot, float *out6);
/* a CUDA-event stopwatch on the slot's stream: start records an event, stop records a
 * second one, waits for it and returns the device time between them (milliseconds) */
int orc_timer_start(orc_ctx *ctx, int slot);
int orc_timer_stop(orc_ctx *ctx, int slot, float *ms);
/* the same over ALL slots: device time of everything launched on any slot between the two calls (every slot's
 * stream waits for the begin event, the end event waits for every stream) -- for throughput measurements with
 * several resident batches in flight, whose kernels overlap as they do in the submit/wait pipeline */
int orc_span_begin(orc_ctx *ctx);
int orc_span_end(orc_ctx *ctx, float *ms);
/* cumulative reads per bin over every batch waited on so far ([n_bins]) */
int orc_counts(orc_ctx *ctx, uint64_t *bins);

/*
 * Host-side FASTQ record indexer (the job of dnaio's C parser under cutadapt): finds the
 * complete 4-line records at the start of `text` and fills the per-read arrays in the raw-text
 * layout of orc_batch.  Stops after max_reads records or at the last complete record;
 * *consumed = bytes used.  Returns the number of records, or a negative ORC_E* code
 * (malformed record: err gets the reason).  A trailing '\r' of a line is not part of it.
 * `final` != 0 means no more text follows: a last record without a newline is accepted.
 */
int64_t orc_fastq_index(const uint8_t *text, uint64_t n_bytes, uint32_t max_reads, int final,
                        uint64_t *seq_offsets, uint32_t *lengths, uint64_t *qual_offsets,
                        uint64_t *name_offsets, uint32_t *name_lengths, uint64_t *consumed,
                        char *err, size_t err_len);

/*
 * Host-side FASTQ(.gz) streaming (csrc/orc_io.cpp): the reader -> workers -> ordered-writer
 * plumbing cutadapt's ParallelPipelineRunner, dnaio and xopen provide around the matching
 * (02_cutadapt_loop.sh:64-72 `-j 24`, input IN.fastq.gz, outputs {name}_<ds>.fastq.gz).
 *
 * Reader: a thread inflates `path` ("-" = stdin; gzip or plain) into a ring of n_buffers text
 * buffers (page-locked when pinned != 0 and a CUDA device is present) and indexes the records.
 * orc_reader_next() hands out the next batch in the raw-text layout of orc_batch (returns 1, or
 * 0 at the end of the input, or a negative ORC_E* code: text in orc_reader_error()); the batch
 * stays valid until its `buffer` is given back with orc_reader_release().
 */
typedef struct orc_reader orc_reader;
typedef struct orc_text_batch {
    uint8_t *text;                  /* the first n_bytes bytes are n_reads complete records */
    uint64_t n_bytes;
    uint32_t n_reads;
    int32_t buffer;                 /* ring slot, for orc_reader_release() */
    uint64_t *offsets;              /* [n_reads] start of the bases inside text */
    uint32_t *lengths;              /* [n_reads] */
    uint64_t *qual_offsets;         /* [n_reads] */
    uint64_t *name_offsets;         /* [n_reads] header without '@' */
    uint32_t *name_lengths;         /* [n_reads] */
    uint64_t total_bases;
} orc_text_batch;

orc_reader *orc_reader_open(const char *path, uint32_t max_reads, uint64_t max_bytes, int n_buffers,
                            int pinned, char *err, size_t err_len);
/* the same with the number of inflate threads given: a .gz written by orc_writer (its members carry a size
 * field) is inflated member by member on that many threads; any other .gz file of 4 MiB or more -- the
 * pychopped_<dataset>.fastq.gz that 02_cutadapt_loop.sh:64-72 reads -- is cut into chunks of its compressed
 * bytes that the pool inflates side by side (csrc/orc_pgz.h: block starts found by trial, text in front of a
 * chunk stood in for by markers, every chunk accepted only where it continues the one before it, CRC-32 and
 * ISIZE of every member checked); stdin, plain text, small files and inflate_threads < 2 go through one zlib
 * stream.  orc_reader_open() picks half the host's hardware threads, at most 8. */
orc_reader *orc_reader_open_threads(const char *path, uint32_t max_reads, uint64_t max_bytes, int n_buffers,
                                    int pinned, int inflate_threads, char *err, size_t err_len);
int orc_reader_next(orc_reader *r, orc_text_batch *out);
int orc_reader_release(orc_reader *r, int buffer);
/* 0: one zlib stream (or plain text), 1: member-parallel, 2: chunk-parallel; stats (may be NULL): text bytes
 * the pool inflated / the reader thread inflated itself so far (mode 2) */
int orc_reader_inflate_mode(orc_reader *r, uint64_t stats[2]);
const char *orc_reader_error(orc_reader *r);
/* A whole gzip file (at least 64 bytes) inflated by the chunk-parallel source alone (csrc/orc_pgz.h) into
 * out[0 .. cap): the number of bytes, ORC_ECAPACITY if they do not fit, ORC_EINVAL for a damaged stream (text in
 * err).  chunk_bytes 0 = 1 MiB.  For tools and tests; orc_reader does the same for its batches. */
int64_t orc_gunzip_file(const char *path, int threads, uint64_t chunk_bytes, uint8_t *out, uint64_t cap,
                        char *err, size_t err_len);
void orc_reader_close(orc_reader *r);

/*
 * Writer: one file per bin (paths[b] == NULL: bin not written), all created by orc_writer_open()
 * even if they stay empty (the reference's round-2 loop lists them, 02_cutadapt_loop.sh:75-85);
 * a path ending in ".gz" is written as gzip members of `level`, deflated by `threads` workers; every member
 * carries an FEXTRA subfield "OC" with its own size (BGZF's idea, 32-bit), which lets orc_reader inflate such
 * files on several threads and which every other gzip reader skips.
 * orc_writer_write() queues the bin-major FASTQ text of one batch (orc_result.fastq /
 * bin_offsets) and returns a ticket >= 0 at once; the text must stay untouched until
 * orc_writer_wait(ticket) returns.  Files receive their batches in submission order.
 * orc_writer_close() drains, closes the files and optionally reports the uncompressed bytes per bin.
 */
typedef struct orc_writer orc_writer;
orc_writer *orc_writer_open(const char *const *paths, int n_bins, int level, int threads, char *err,
                            size_t err_len);
/* orc_writer_set_index(w, 1) before the first write: orc_writer_close() also leaves PATH.idx beside every bin
 * file, (ticket, bytes) as two little-endian uint64 per chunk in file order -- what merging the part files
 * of several GPUs' ranks into one file in batch order needs (orcdemux/cli.py).  orc_empty_gzip_member():
 * the empty member a bin without reads consists of (returns its length, 0 on failure; cap >= 40). */
int orc_writer_set_index(orc_writer *w, int on);
size_t orc_empty_gzip_member(uint8_t *out, size_t cap, int level);
int64_t orc_writer_write(orc_writer *w, const uint8_t *fastq, const uint64_t *bin_offsets);
/* the same for a batch that comes back as gzip members (orc_params.emit_gzip): members = orc_result.fastq,
 * member_offsets = orc_result.bin_offsets.  A .gz bin file receives its member as it is (no host deflate), a
 * plain file the inflated text. */
int64_t orc_writer_write_members(orc_writer *w, const uint8_t *members, const uint64_t *member_offsets);
int orc_writer_wait(orc_writer *w, int64_t ticket);
const char *orc_writer_error(orc_writer *w);
int orc_writer_close(orc_writer *w, uint64_t *bytes_per_bin);

/*
 * Pairwise unit-cost edit distance of whole sequences (csrc/orc_edit.cuh), the drop-in for what
 * amplicon_sorter asks of edlib for every pair it compares
 * (/root/reference/scripts/auxiliary_code/amplicon_sorter.py:225-235 distance(), mode NW;
 *  :838-849 distance_finetune(), mode HW):
 *      edlib.align(A1, A2, task='distance', mode=mode)['editDistance'], A1 the shorter sequence.
 * seqs/offsets/lengths: n_seqs sequences in one blob (any bytes, compared for equality; at most 8
 * distinct bytes per call); pair k compares sequences pair_a[k] and pair_b[k] (the shorter one is
 * the query; on equal lengths pair_a[k] is) and dist[k] receives the distance.  Host buffers in
 * and out; *kernel_ms (optional) = device time of the kernels.  Returns 0 or a negative ORC_E*
 * code with the reason in err.  Queries longer than 8192 are refused; no CPU fallback.
 */
enum { ORC_EDIT_NW = 0, ORC_EDIT_HW = 1 };
int orc_edit_distances(int device, const uint8_t *seqs, const uint64_t *offsets, const uint32_t *lengths,
                       uint32_t n_seqs, const uint32_t *pair_a, const uint32_t *pair_b, uint64_t n_pairs,
                       int mode, uint32_t *dist, float *kernel_ms, char *err, size_t err_len);

/*
 * Synthetic input (benchmarks and tests; the reference ships no reads).  orc_synth() makes n_reads ONT-like
 * reads of the SURVEY 8d read model -- 5' adapter of round 1 + random insert + 3' adapter of round 2, iid
 * sequencing errors, truncations, missing adapters, reverse-complemented reads, total length uniform in
 * [len_min, len_max] -- on the device, straight into the slot, which is then in the state orc_upload() leaves
 * it in (orc_launch / orc_download / orc_wait follow).  Deterministic in (seed, n_reads, len_min, len_max).
 * Needs two rounds of plain unanchored adapters and emit_fastq.  orc_resident() gives the sizes of the batch
 * resident in a slot, orc_export() copies it to host buffers in the separate-blob layout of orc_batch
 * (name_offsets: n_reads + 1 entries), e.g. to hand the same bytes to a CPU checker.
 */
int orc_synth(orc_ctx *ctx, int slot, uint64_t seed, uint32_t n_reads, uint32_t len_min, uint32_t len_max);
int orc_resident(orc_ctx *ctx, int slot, uint32_t *n_reads, uint64_t *n_bytes, uint64_t *name_bytes);
int orc_export(orc_ctx *ctx, int slot, uint8_t *seq, uint8_t *qual, uint64_t *offsets, uint32_t *lengths,
               uint8_t *names, uint64_t *name_offsets);

/* pinned host memory for callers that do not bring their own */
void *orc_host_alloc(size_t bytes);
void orc_host_free(void *p);

/* micro-benchmark: issue rate of dependent-free 32-bit integer instructions, in lane-ops per
 * second (the DP kernel's roofline denominator).  mode 0: LOP3 only (the ALU
 * pipe the Myers recurrence lives on); mode 1: LOP3 + IMAD mix (ALU and FMA pipes together) */
double orc_measure_int32_peak(int device, int mode, double *sm_clock_mhz);
/* GB/s of kernel loads from PINNED host memory: every warp reads `chunk` of every `stride` bytes (multiples of
 * 16) of host[0 .. bytes) with 16-byte loads per lane, like emit_kernel's record copies (the ceiling of
 * orc_params.qual_zero_copy).  -2 if the buffer is not pinned, -1 on other failures. */
double orc_probe_hostread(int device, const void *host, uint64_t bytes, uint32_t chunk, uint64_t stride);

const char *orc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ORCDEMUX_H */
