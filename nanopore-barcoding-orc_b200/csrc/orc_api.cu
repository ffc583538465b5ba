// orc_api.cu -- C ABI of liborcdemux.so (include/orcdemux.h): context, device arena,
// streams, and the launch sequence of one batch.  Replaces the inside of the two cutadapt
// invocations of /root/reference/scripts/02_cutadapt_loop.sh:64-72 and :94-102.
//
// One ctx per GPU; n_slots batches can be in flight, each on its own stream with its own
// device arena, so that the H2D copy of batch i+1 and the D2H copy of batch i-1 overlap the
// kernels of batch i.  No CPU fallback anywhere: every entry point needs the device.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/orcdemux.h"

#include "orc_kernels.cuh"
#include "orc_gz.cuh"
#include "orc_synth.cuh"
#include "orc_table.h"

using namespace orc;

static_assert(sizeof(orc_match) == sizeof(Match), "orc_match layout");
static_assert(sizeof(View) == 16, "View layout");
static_assert(sizeof(Task) == 32, "Task layout");
static_assert(sizeof(PairResult) == 32, "PairResult layout");
static_assert(sizeof(WinList) == 32, "WinList layout");
static_assert(ORC_MAX_ADAPTERS == MAX_AD, "adapter limit");
static_assert(ORC_MAX_LONG_ADAPTER_LEN == MAX_M_LONG && ORC_MAX_LONG_ADAPTERS == MAX_AD_LONG, "long-adapter limits");
static_assert(MAX_BINS_GZ == MAX_BINS, "bins of the gzip encoder");

namespace {

enum { EV_START = 0, EV_H2D, EV_PACK, EV_TRIG0, EV_SCAN0, EV_RES0, EV_TRIG1, EV_SCAN1, EV_RES1, EV_BIN, EV_EMIT, EV_HDR, EV_END, EV_T0, EV_T1, EV_GZ, EV_COUNT };
// d_counters: 16 per-round counters, then for round r and ordering o (0: reads by length, 1: items by
// window columns) a histogram and a cursor array of SORT_BUCKETS words each
constexpr size_t N_COUNTERS = 16 + 2 * 2 * 2 * (size_t)SORT_BUCKETS;
static inline uint32_t *sort_hist(uint32_t *counters, int round, int ordering, int cursor)
{
    return counters + 16 + (size_t)(((round * 2 + ordering) * 2 + cursor)) * SORT_BUCKETS;
}

enum { SLOT_IDLE = 0, SLOT_UPLOADED, SLOT_LAUNCHED, SLOT_DOWNLOADING };

struct Slot {
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;             // the wide resolver runs here, beside the band resolver
    cudaEvent_t ev_fork[ORC_MAX_ROUNDS] = {}, ev_join[ORC_MAX_ROUNDS] = {}, ev_w0[ORC_MAX_ROUNDS] = {};
    int state = SLOT_IDLE;
    uint32_t n_reads = 0;
    uint64_t n_bytes = 0, name_bytes = 0, in_bases = 0;
    bool has_names = false;
    bool did_h2d = false, did_kernels = false, did_d2h = false, fresh_upload = false;
    // device
    uint8_t *d_seq = nullptr, *d_qual = nullptr, *d_names = nullptr, *d_fastq = nullptr;
    uint8_t *u_qual = nullptr, *u_names = nullptr;      // what the kernels use (may alias d_seq, or the caller's pinned buffer)
    bool qual_in_place = false;                         // u_qual is the caller's page-locked buffer (orc_params.qual_zero_copy)
    uint64_t *d_qual_offsets = nullptr, *u_qual_offsets = nullptr;
    uint32_t *d_name_lengths = nullptr, *u_name_lengths = nullptr;
    uint32_t *d_codes_alloc = nullptr;
    uint64_t *d_offsets = nullptr, *d_name_offsets = nullptr, *d_dest = nullptr;
    uint32_t *d_lengths = nullptr;
    View *d_views[3] = {nullptr, nullptr, nullptr};
    Match *d_match[2] = {nullptr, nullptr};
    uint32_t *d_wcols = nullptr, *d_wcols_sorted = nullptr, *d_item_order = nullptr;
    unsigned long long *d_best_key = nullptr;
    uint32_t *d_order = nullptr;             // reads in order of decreasing view length (stage 1)
    WinList *d_wins = nullptr;
    SeedWins *d_seedwins = nullptr;  // stage 1s: seed windows per (segment, read, direction)
    uint32_t max_len = 0;            // longest read of the resident batch (how many segments seed_kernel needs)
    Task *d_tasks = nullptr;
    uint32_t *d_jobs = nullptr;              // stage 2a survivors: job numbers (item * n_adapters + adapter)
    PairResult *d_results = nullptr;
    uint32_t *d_counters = nullptr;          // [0..15] per-round counters, then the size-class histograms and cursors
                                             // of the two orderings of each round (N_COUNTERS words, zeroed per launch)
    unsigned long long *d_cells = nullptr;   // [0..1] sum of view lengths entering each round, [2..3] window columns,
                                             // [4..5] cells stage 2b updated (rows x columns of the pairs that passed 2a)
    int32_t *d_bin = nullptr;
    uint32_t *d_out_len = nullptr, *d_rec_bytes = nullptr, *d_hist_cnt = nullptr;
    uint64_t *d_hist_bytes = nullptr, *d_bin_counts = nullptr, *d_bin_offsets = nullptr, *d_bin_bytes = nullptr;
    // pinned host
    Match *h_match[2] = {nullptr, nullptr};
    int32_t *h_bin = nullptr;
    uint32_t *h_out_len = nullptr, *h_counters = nullptr;
    uint64_t *h_bin_counts = nullptr, *h_bin_offsets = nullptr;
    unsigned long long *h_cells = nullptr;
    uint8_t *h_fastq = nullptr;
    // orc_params.emit_gzip (orc_gz.cuh): the bins of the batch as gzip members
    uint8_t *d_gz = nullptr;
    GzTable *d_gz_table = nullptr;
    unsigned long long *d_gz_hist = nullptr;
    uint32_t *d_gz_chunk_base = nullptr, *d_gz_chunk_local = nullptr, *d_gz_tile_bits = nullptr, *d_gz_member_crc = nullptr;
    uint64_t *d_gz_tile_off = nullptr, *d_gz_member_pos = nullptr, *d_gz_member_bits = nullptr, *d_gz_member_bytes = nullptr, *d_gz_offsets = nullptr;
    uint64_t *h_gz_offsets = nullptr;
    uint32_t n_launches = 0;                 // own kernels of the last orc_launch()
    size_t cap_pairs = 0;                    // entries of d_tasks / d_results (see alloc_slot)
    cudaEvent_t ev[EV_COUNT] = {};
    cudaEvent_t evk[ORC_MAX_ROUNDS][ORC_N_KERNELS + 1] = {};   // [r][0]: before the round's first kernel, [r][1 + K]: after kernel K
};

}  // namespace

struct orc_ctx {
    int device = 0;
    int n_rounds = 1, n_slots = 1, n_bins = 1, sm_count = 148;
    int emit_fastq = 1, want_matches = 1, qual_zero_copy = 0, emit_gzip = 0;
    uint64_t gz_cap = 0, gz_max_chunks = 0;  // bytes of a slot's gzip arena, chunks of its text at most
    int emit_zc_blocks = 8;                  // emit_kernel blocks per SM when it reads the qualities from host memory (measured: 8 > 4 > 2 > 1)
    uint32_t max_reads = 0;
    uint64_t max_bytes = 0, max_name_bytes = 0, fastq_cap = 0;
    RoundTable h_tab[2];
    SeedTable h_seed[2];             // stage 1s (on == 0: the round keeps the flank scan)
    SeedTable *d_seed[2] = {nullptr, nullptr};
    RoundTable *d_tab[2] = {nullptr, nullptr};
    AnchoredTable h_anch[2];
    AnchoredTable *d_anch[2] = {nullptr, nullptr};
    bool anchored[2] = {false, false};
    LongTable h_long[2];
    LongTable *d_long[2] = {nullptr, nullptr};
    bool need_u_check = false;              // plain and IUPAC adapters side by side: a batch whose reads hold U is refused
    bool wild_codes = false;                // the reads are packed with U = T (some round compares through the IUPAC masks)
    bool longr[2] = {false, false};         // a round with an adapter over 64 nt (long_kernel)
    bool special[2] = {false, false};       // anchored or long: the round does not take the bit-parallel pipeline
    uint8_t *d_pack_lut = nullptr, *d_comp_lut = nullptr, *d_drop = nullptr;
    cudaEvent_t ev_span[2] = {nullptr, nullptr};   // orc_span_begin / orc_span_end
    cudaEvent_t ev_ref = nullptr;                  // recorded in orc_create: origin of orc_timings.timeline_ms
    std::vector<cudaEvent_t> ev_span_slot;
    SynthTable *d_synth = nullptr;   // orc_synth(): built on first use
    uint64_t *d_synth_totals = nullptr;
    std::vector<Slot> slots;
    std::vector<uint64_t> total_counts;
    std::string err;
    int scan_blocks = 0, resolve_blocks = 0, filter_blocks = 0, band_blocks[2] = {0, 0};
    int na_max = 1;                  // most adapters of an unanchored round (sizes the pair arenas)
};

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                \
            return ORC_ECUDA;                                                             \
        }                                                                                 \
    } while (0)

template <typename T>
static cudaError_t dalloc(T **p, size_t n)
{
    return cudaMalloc(reinterpret_cast<void **>(p), n * sizeof(T) + 64);
}
template <typename T>
static cudaError_t halloc(T **p, size_t n)
{
    return cudaHostAlloc(reinterpret_cast<void **>(p), n * sizeof(T) + 64, cudaHostAllocDefault);
}

static int alloc_slot(orc_ctx *ctx, Slot &s)
{
    const size_t R = ctx->max_reads, B = (size_t)ctx->max_bytes;
    const size_t n_chunks = (R + BIN_CHUNK - 1) / BIN_CHUNK;
    // Pair arenas (Task / PairResult, 32 bytes each): the worst case is every (read, direction, adapter) pair
    // holding a candidate, 2 * na per read; real data has about 1.4 per read and round.  Slots start with
    // room for 4 per read (and never less than 65536, so small batches of adversarial reads fit outright);
    // a batch that needs more is caught by its counters in orc_wait(), which grows the arenas to the worst
    // case and runs the batch again (grow_pair_arenas).
    const size_t worst = std::max<size_t>(R * 2 * (size_t)ctx->na_max, 2 * R);
    size_t n_tasks = std::min(worst, std::max<size_t>(4 * R, 65536));
    if (const char *e = getenv("ORC_PAIR_CAP")) {       // tests: force the overflow path
        const size_t floor_ = (ctx->special[0] || ctx->special[1]) ? 2 * R : 1;
        n_tasks = std::min(worst, std::max<size_t>((size_t)atoll(e), floor_));
    }
    s.cap_pairs = n_tasks;
    CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s.side, cudaStreamNonBlocking));
    for (int r = 0; r < ORC_MAX_ROUNDS; r++) {
        CK(cudaEventCreateWithFlags(&s.ev_fork[r], cudaEventDisableTiming));
        CK(cudaEventCreate(&s.ev_join[r]));
        CK(cudaEventCreate(&s.ev_w0[r]));
    }
    for (int i = 0; i < EV_COUNT; i++) CK(cudaEventCreate(&s.ev[i]));
    for (int r = 0; r < ORC_MAX_ROUNDS; r++)
        for (int i = 0; i <= ORC_N_KERNELS; i++) CK(cudaEventCreate(&s.evk[r][i]));
    CK(dalloc(&s.d_seq, B + 64));
    CK(dalloc(&s.d_qual, B + 64));
    CK(dalloc(&s.d_codes_alloc, B / 8 + 16 + 2 * GUARD_WORDS));
    CK(cudaMemset(s.d_codes_alloc, 0, (B / 8 + 16 + 2 * GUARD_WORDS) * sizeof(uint32_t)));
    CK(dalloc(&s.d_offsets, R));
    CK(dalloc(&s.d_qual_offsets, R));
    CK(dalloc(&s.d_lengths, R));
    for (int i = 0; i < 3; i++) CK(dalloc(&s.d_views[i], R));
    for (int i = 0; i < 2; i++) CK(dalloc(&s.d_match[i], R));
    CK(dalloc(&s.d_wcols, 2 * R)); CK(dalloc(&s.d_wcols_sorted, 2 * R));
    CK(dalloc(&s.d_item_order, 2 * R));
    CK(dalloc(&s.d_best_key, 2 * R));
    CK(dalloc(&s.d_order, R));
    CK(dalloc(&s.d_wins, 2 * R));
    CK(dalloc(&s.d_seedwins, (size_t)SEED_SEGS_MAX * 2 * R));    // one plane of 2 * R per segment (seed_kernel)
    CK(dalloc(&s.d_tasks, 2 * n_tasks));           // front half: band resolver, back half: wide resolver
    CK(dalloc(&s.d_jobs, 2 * R * (size_t)ctx->na_max));
    CK(dalloc(&s.d_results, n_tasks));
    CK(dalloc(&s.d_counters, N_COUNTERS));
    CK(dalloc(&s.d_cells, 6));
    CK(dalloc(&s.d_bin, R));
    CK(dalloc(&s.d_out_len, R));
    CK(dalloc(&s.d_rec_bytes, R));
    CK(dalloc(&s.d_dest, R));
    CK(dalloc(&s.d_hist_cnt, n_chunks * ctx->n_bins));
    CK(dalloc(&s.d_hist_bytes, n_chunks * ctx->n_bins));
    CK(dalloc(&s.d_bin_counts, (size_t)ctx->n_bins));
    CK(dalloc(&s.d_bin_bytes, (size_t)ctx->n_bins));
    CK(dalloc(&s.d_bin_offsets, (size_t)ctx->n_bins + 1));
    CK(halloc(&s.h_bin, R));
    CK(halloc(&s.h_out_len, R));
    CK(halloc(&s.h_counters, 16));
    CK(halloc(&s.h_cells, 6));
    CK(halloc(&s.h_bin_counts, (size_t)ctx->n_bins));
    CK(halloc(&s.h_bin_offsets, (size_t)ctx->n_bins + 1));
    if (ctx->want_matches)
        for (int i = 0; i < 2; i++) CK(halloc(&s.h_match[i], R));
    if (ctx->emit_fastq) {
        CK(dalloc(&s.d_names, (size_t)ctx->max_name_bytes + 64));
        CK(dalloc(&s.d_name_offsets, R + 1));
        CK(dalloc(&s.d_name_lengths, R));
        CK(dalloc(&s.d_fastq, (size_t)ctx->fastq_cap));
        if (ctx->emit_gzip) {
            CK(dalloc(&s.d_gz, (size_t)ctx->gz_cap));
            CK(dalloc(&s.d_gz_table, 1));
            CK(dalloc(&s.d_gz_hist, 256));
            CK(dalloc(&s.d_gz_chunk_base, (size_t)ctx->n_bins + 1));
            CK(dalloc(&s.d_gz_chunk_local, (size_t)ctx->gz_max_chunks));
            CK(dalloc(&s.d_gz_tile_bits, (size_t)ctx->gz_max_chunks / GZ_TILE + 2));
            CK(dalloc(&s.d_gz_tile_off, (size_t)ctx->gz_max_chunks / GZ_TILE + 2));
            CK(dalloc(&s.d_gz_member_pos, (size_t)ctx->n_bins));
            CK(dalloc(&s.d_gz_member_bits, (size_t)ctx->n_bins));
            CK(dalloc(&s.d_gz_member_bytes, (size_t)ctx->n_bins));
            CK(dalloc(&s.d_gz_member_crc, (size_t)ctx->n_bins));
            CK(dalloc(&s.d_gz_offsets, (size_t)ctx->n_bins + 1));
            CK(halloc(&s.h_gz_offsets, (size_t)ctx->n_bins + 1));
            GzTable *T = new GzTable();
            memset(T, 0, sizeof(GzTable));
            gz_fill_crc_tables(*T);
            const cudaError_t e = cudaMemcpy(s.d_gz_table, T, sizeof(GzTable), cudaMemcpyHostToDevice);
            delete T;
            CK(e);
        }
        // h_fastq (page-locked, as large as the device arena) is allocated by the first orc_wait() that has
        // text to fetch: slots that are only ever launched (device-resident shards) never pay for it
    }
    return ORC_OK;
}

static void free_slot(Slot &s)
{
    cudaFree(s.d_seq); cudaFree(s.d_qual); cudaFree(s.d_names); cudaFree(s.d_fastq);
    cudaFree(s.d_gz); cudaFree(s.d_gz_table); cudaFree(s.d_gz_hist); cudaFree(s.d_gz_chunk_base); cudaFree(s.d_gz_chunk_local);
    cudaFree(s.d_gz_tile_bits); cudaFree(s.d_gz_tile_off); cudaFree(s.d_gz_member_pos); cudaFree(s.d_gz_member_bits); cudaFree(s.d_gz_member_bytes);
    cudaFree(s.d_gz_member_crc); cudaFree(s.d_gz_offsets); cudaFreeHost(s.h_gz_offsets);
    cudaFree(s.d_codes_alloc); cudaFree(s.d_offsets); cudaFree(s.d_name_offsets); cudaFree(s.d_dest);
    cudaFree(s.d_lengths); cudaFree(s.d_qual_offsets); cudaFree(s.d_name_lengths);
    for (int i = 0; i < 3; i++) cudaFree(s.d_views[i]);
    for (int i = 0; i < 2; i++) { cudaFree(s.d_match[i]); cudaFreeHost(s.h_match[i]); }
    cudaFree(s.d_wcols); cudaFree(s.d_wcols_sorted); cudaFree(s.d_item_order);
    cudaFree(s.d_best_key); cudaFree(s.d_tasks); cudaFree(s.d_results); cudaFree(s.d_jobs);
    cudaFree(s.d_order);
    cudaFree(s.d_wins); cudaFree(s.d_seedwins);
    cudaFree(s.d_counters); cudaFree(s.d_cells); cudaFree(s.d_bin); cudaFree(s.d_out_len);
    cudaFree(s.d_rec_bytes); cudaFree(s.d_hist_cnt); cudaFree(s.d_hist_bytes);
    cudaFree(s.d_bin_counts); cudaFree(s.d_bin_offsets); cudaFree(s.d_bin_bytes);
    cudaFreeHost(s.h_bin); cudaFreeHost(s.h_out_len); cudaFreeHost(s.h_counters); cudaFreeHost(s.h_cells);
    cudaFreeHost(s.h_bin_counts); cudaFreeHost(s.h_bin_offsets); cudaFreeHost(s.h_fastq);
    for (int i = 0; i < EV_COUNT; i++) if (s.ev[i]) cudaEventDestroy(s.ev[i]);
    for (int r = 0; r < ORC_MAX_ROUNDS; r++)
        for (int i = 0; i <= ORC_N_KERNELS; i++) if (s.evk[r][i]) cudaEventDestroy(s.evk[r][i]);
    for (int r = 0; r < ORC_MAX_ROUNDS; r++) {
        if (s.ev_fork[r]) cudaEventDestroy(s.ev_fork[r]);
        if (s.ev_join[r]) cudaEventDestroy(s.ev_join[r]);
        if (s.ev_w0[r]) cudaEventDestroy(s.ev_w0[r]);
    }
    if (s.side) cudaStreamDestroy(s.side);
    if (s.stream) cudaStreamDestroy(s.stream);
}

static int ctx_init(orc_ctx *ctx, const orc_params *p)
{
    if (p->n_rounds < 1 || p->n_rounds > ORC_MAX_ROUNDS) { ctx->err = "n_rounds must be 1 or 2"; return ORC_EINVAL; }
    if (p->max_reads == 0 || p->max_bytes == 0) { ctx->err = "max_reads and max_bytes must be > 0"; return ORC_EINVAL; }
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        ctx->err = std::string("no CUDA device (liborcdemux has no CPU fallback): ") + cudaGetErrorString(e);
        return ORC_ECUDA;
    }
    ctx->device = p->device;
    CK(cudaSetDevice(p->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, p->device));
    if (prop.major < 10) {
        ctx->err = "liborcdemux is built for sm_100a (B200) only; device is sm_" + std::to_string(prop.major) +
                   std::to_string(prop.minor);
        return ORC_ECUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->n_rounds = p->n_rounds;
    ctx->n_slots = p->n_slots < 1 ? 1 : p->n_slots;
    ctx->emit_fastq = p->emit_fastq ? 1 : 0;
    ctx->want_matches = p->want_matches ? 1 : 0;
    ctx->qual_zero_copy = p->qual_zero_copy ? 1 : 0;
    ctx->emit_gzip = (p->emit_gzip && p->emit_fastq) ? 1 : 0;
    if (const char *e = getenv("ORC_EMIT_ZC_BLOCKS")) ctx->emit_zc_blocks = std::min(8, std::max(1, atoi(e)));
    ctx->max_reads = p->max_reads;
    ctx->max_bytes = (p->max_bytes + 63) & ~63ull;
    ctx->max_name_bytes = p->max_name_bytes ? p->max_name_bytes : 64;
    for (int r = 0; r < p->n_rounds; r++) {
        const orc_round_params &rp = p->rounds[r];
        if (rp.sequences == nullptr) { ctx->err = "round without adapter sequences"; return ORC_EINVAL; }
        std::string why;
        ctx->anchored[r] = (rp.type == ORC_PREFIX || rp.type == ORC_SUFFIX);
        ctx->longr[r] = !ctx->anchored[r] && rp.n_adapters >= 1 && round_is_long(rp.n_adapters, rp.sequences);
        ctx->special[r] = ctx->anchored[r] || ctx->longr[r];
        if (ctx->anchored[r])
            why = build_anchored_table(ctx->h_anch[r], ctx->h_tab[r], rp.n_adapters, rp.type == ORC_SUFFIX,
                                       rp.sequences, rp.max_error_rate, rp.indels, rp.revcomp);
        else if (ctx->longr[r])
            why = build_long_table(ctx->h_long[r], ctx->h_tab[r], rp.n_adapters, rp.type, rp.sequences,
                                   rp.max_error_rate, rp.min_overlap, rp.indels, rp.revcomp);
        else
            why = build_round_table(ctx->h_tab[r], rp.n_adapters, rp.type, rp.sequences,
                                    rp.max_error_rate, rp.min_overlap, rp.indels, rp.revcomp);
        if (!why.empty()) { ctx->err = why; return ORC_EINVAL; }
        if (rp.action != ORC_ACTION_TRIM && rp.action != ORC_ACTION_RETAIN) { ctx->err = "unsupported: action must be trim or retain"; return ORC_EINVAL; }
        ctx->h_tab[r].action = rp.action;
    }
    {
        // one code array serves every round: with IUPAC adapters anywhere, all rounds compare through the masks; the
        // plain adapters among them differ from cutadapt's ASCII comparison only on a read with U (u_scan_kernel)
        bool uses_codes[ORC_MAX_ROUNDS] = {};
        for (int r = 0; r < p->n_rounds; r++) uses_codes[r] = !ctx->anchored[r];
        ctx->need_u_check = unify_wildcards(ctx->h_tab, uses_codes, p->n_rounds);
        ctx->wild_codes = false;
        for (int r = 0; r < p->n_rounds; r++) ctx->wild_codes = ctx->wild_codes || (uses_codes[r] && ctx->h_tab[r].wild != 0);
    }
    for (int r = 0; r < p->n_rounds; r++) {
        ctx->h_seed[r].on = 0;
        if (!ctx->special[r]) {
            const char *off = getenv("ORC_NO_SEED");        // A/B measurements: keep the flank scan
            build_seed_table(ctx->h_tab[r], ctx->h_seed[r], !(off && off[0] == '1'));
        }
    }
    for (int r = 0; r < p->n_rounds; r++)
        if (!ctx->special[r] && ctx->h_tab[r].n_adapters > ctx->na_max) ctx->na_max = ctx->h_tab[r].n_adapters;
    ctx->n_bins = ctx->h_tab[0].n_adapters + 1;
    if (p->n_rounds == 2) ctx->n_bins *= ctx->h_tab[1].n_adapters + 1;
    if (ctx->n_bins > MAX_BINS) { ctx->err = "unsupported: more than 512 bins"; return ORC_EINVAL; }
    ctx->total_counts.assign((size_t)ctx->n_bins, 0);
    ctx->fastq_cap = ctx->max_name_bytes + 2 * ctx->max_bytes + 16ull * ctx->max_reads + 64;
    // the flat code orc_wait() falls back to takes 9 bits per byte at most; frame and block header per member
    ctx->gz_cap = (ctx->fastq_cap + ctx->fastq_cap / 8 + 256ull * (uint64_t)ctx->n_bins + 64 + 15) & ~15ull;
    ctx->gz_max_chunks = ctx->fastq_cap / GZ_CHUNK + (uint64_t)ctx->n_bins + 2;
    if (ctx->emit_gzip && ctx->gz_cap >= (1ull << 32)) {
        // a member's "OC" size field and its ISIZE are 32 bits wide
        ctx->err = "emit_gzip needs batches under 4 GiB of FASTQ text (smaller max_bytes)";
        return ORC_EINVAL;
    }
    for (int r = 0; r < p->n_rounds; r++) {
        CK(dalloc(&ctx->d_tab[r], 1));
        CK(cudaMemcpy(ctx->d_tab[r], &ctx->h_tab[r], sizeof(RoundTable), cudaMemcpyHostToDevice));
        if (!ctx->special[r] && ctx->h_seed[r].on) {
            CK(dalloc(&ctx->d_seed[r], 1));
            CK(cudaMemcpy(ctx->d_seed[r], &ctx->h_seed[r], sizeof(SeedTable), cudaMemcpyHostToDevice));
        }
        if (ctx->anchored[r]) {
            CK(dalloc(&ctx->d_anch[r], 1));
            CK(cudaMemcpy(ctx->d_anch[r], &ctx->h_anch[r], sizeof(AnchoredTable), cudaMemcpyHostToDevice));
        }
        if (ctx->longr[r]) {
            CK(dalloc(&ctx->d_long[r], 1));
            CK(cudaMemcpy(ctx->d_long[r], &ctx->h_long[r], sizeof(LongTable), cudaMemcpyHostToDevice));
        }
    }
    uint8_t lut[256];
    build_pack_lut(lut, ctx->wild_codes);
    CK(dalloc(&ctx->d_pack_lut, 256));
    CK(cudaMemcpy(ctx->d_pack_lut, lut, 256, cudaMemcpyHostToDevice));
    build_complement_lut(lut);
    CK(dalloc(&ctx->d_comp_lut, 256));
    CK(cudaMemcpy(ctx->d_comp_lut, lut, 256, cudaMemcpyHostToDevice));
    std::vector<uint8_t> drop((size_t)ctx->n_bins, 0);
    if (p->drop_bins) memcpy(drop.data(), p->drop_bins, (size_t)ctx->n_bins);
    CK(dalloc(&ctx->d_drop, (size_t)ctx->n_bins));
    CK(cudaMemcpy(ctx->d_drop, drop.data(), (size_t)ctx->n_bins, cudaMemcpyHostToDevice));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, scan_kernel, SCAN_THREADS, 0));
    ctx->scan_blocks = ctx->sm_count * (occ > 0 ? occ : 1);
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, filter_kernel, SCAN_THREADS, 0));
    ctx->filter_blocks = ctx->sm_count * (occ > 0 ? occ : 1);
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, resolve_kernel, 128, 0));
    {   // more resident threads can thrash the per-thread ring in L1/L2 (ORC_RESOLVE_OCC: tuning knob)
        const char *e = getenv("ORC_RESOLVE_OCC");
        const int cap = e ? atoi(e) : 6;
        if (cap > 0 && occ > cap) occ = cap;
    }
    ctx->resolve_blocks = ctx->sm_count * (occ > 0 ? occ : 1);
    CK(cudaFuncSetAttribute(resolve_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)band_smem_bytes(MAX_LANES, MAX_AD)));
    for (int r = 0; r < p->n_rounds; r++) {
        if (ctx->special[r]) continue;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, resolve_band_kernel, BAND_THREADS,
                                                         band_smem_bytes(ctx->h_tab[r].n_lanes, ctx->h_tab[r].n_adapters)));
        ctx->band_blocks[r] = ctx->sm_count * (occ > 0 ? occ : 1);
    }
    ctx->slots.resize((size_t)ctx->n_slots);
    for (auto &s : ctx->slots) {
        int rc = alloc_slot(ctx, s);
        if (rc != ORC_OK) return rc;
    }
    CK(cudaEventCreate(&ctx->ev_ref));
    CK(cudaEventRecord(ctx->ev_ref, ctx->slots[0].stream));
    CK(cudaEventSynchronize(ctx->ev_ref));
    return ORC_OK;
}

extern "C" orc_ctx *orc_create(const orc_params *params, char *err, size_t err_len)
{
    orc_ctx *ctx = new orc_ctx();
    int rc = params ? ctx_init(ctx, params) : ORC_EINVAL;
    if (rc != ORC_OK) {
        if (err && err_len) {
            snprintf(err, err_len, "%s", params ? ctx->err.c_str() : "params is NULL");
        }
        orc_destroy(ctx);
        return nullptr;
    }
    if (err && err_len) err[0] = 0;
    return ctx;
}

extern "C" void orc_destroy(orc_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (auto &s : ctx->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        free_slot(s);
    }
    for (int r = 0; r < 2; r++) { cudaFree(ctx->d_tab[r]); cudaFree(ctx->d_anch[r]); cudaFree(ctx->d_long[r]); cudaFree(ctx->d_seed[r]); }
    cudaFree(ctx->d_pack_lut); cudaFree(ctx->d_comp_lut); cudaFree(ctx->d_drop);
    cudaFree(ctx->d_synth); cudaFree(ctx->d_synth_totals);
    for (int i = 0; i < 2; i++) if (ctx->ev_span[i]) cudaEventDestroy(ctx->ev_span[i]);
    if (ctx->ev_ref) cudaEventDestroy(ctx->ev_ref);
    for (cudaEvent_t e : ctx->ev_span_slot) if (e) cudaEventDestroy(e);
    delete ctx;
}

extern "C" const char *orc_last_error(orc_ctx *ctx) { return ctx ? ctx->err.c_str() : "ctx is NULL"; }
extern "C" int orc_n_bins(orc_ctx *ctx) { return ctx ? ctx->n_bins : ORC_EINVAL; }
extern "C" const char *orc_version(void) { return "orcdemux 0.1.0 (sm_100a)"; }

static Slot *get_slot(orc_ctx *ctx, int slot)
{
    if (!ctx) return nullptr;
    if (slot < 0 || slot >= ctx->n_slots) { ctx->err = "slot out of range"; return nullptr; }
    return &ctx->slots[(size_t)slot];
}

extern "C" int orc_upload(orc_ctx *ctx, int slot, const orc_batch *b)
{
    Slot *sp = get_slot(ctx, slot);
    if (!sp) return ORC_EINVAL;
    Slot &s = *sp;
    if (!b) { ctx->err = "batch is NULL"; return ORC_EINVAL; }
    if (b->n_reads > ctx->max_reads || b->n_bytes > ctx->max_bytes) {
        ctx->err = "batch exceeds max_reads/max_bytes given to orc_create";
        return ORC_ECAPACITY;
    }
    if (b->n_reads && (!b->seq || !b->qual || !b->offsets || !b->lengths)) {
        ctx->err = "batch buffers missing"; return ORC_EINVAL;
    }
    s.has_names = ctx->emit_fastq != 0;
    const bool names_alias = s.has_names && b->names == b->seq;
    uint64_t name_bytes = 0;
    if (s.has_names) {
        if (b->n_reads && (!b->names || !b->name_offsets)) {
            ctx->err = "emit_fastq needs names and name_offsets"; return ORC_EINVAL;
        }
        if (!names_alias) {
            name_bytes = b->name_lengths ? b->name_bytes : (b->n_reads ? b->name_offsets[b->n_reads] : 0);
            if (name_bytes > ctx->max_name_bytes) { ctx->err = "names exceed max_name_bytes"; return ORC_ECAPACITY; }
        }
    }
    CK(cudaSetDevice(ctx->device));
    s.n_reads = b->n_reads;
    s.n_bytes = b->n_bytes;
    s.name_bytes = name_bytes;
    uint64_t bases = 0;
    uint32_t max_len = 0;
    for (uint32_t r = 0; r < b->n_reads; r++) {
        if (b->offsets[r] + b->lengths[r] > b->n_bytes ||
            (b->qual_offsets && b->qual_offsets[r] + b->lengths[r] > b->n_bytes)) {
            ctx->err = "read extends past n_bytes"; return ORC_EINVAL;
        }
        bases += b->lengths[r];
        if (b->lengths[r] > max_len) max_len = b->lengths[r];
    }
    if (s.has_names && b->n_reads) {
        // the header lines: inside the text (raw FASTQ layout) or inside the names blob, never past it
        const uint64_t bound = names_alias ? b->n_bytes : name_bytes;
        for (uint32_t r = 0; r < b->n_reads; r++) {
            uint64_t len;
            if (b->name_lengths) len = b->name_lengths[r];
            else {
                if (b->name_offsets[r + 1] < b->name_offsets[r]) { ctx->err = "name_offsets must not decrease"; return ORC_EINVAL; }
                len = b->name_offsets[r + 1] - b->name_offsets[r];
            }
            if (b->name_offsets[r] > bound || len > bound - b->name_offsets[r]) {
                ctx->err = "read name extends past the names blob"; return ORC_EINVAL;
            }
        }
    }
    s.in_bases = bases;
    s.max_len = max_len;
    CK(cudaEventRecord(s.ev[EV_START], s.stream));
    s.u_qual = s.d_qual; s.u_names = s.d_names; s.u_qual_offsets = nullptr; s.u_name_lengths = nullptr;
    if (b->n_reads) {
        CK(cudaMemcpyAsync(s.d_seq, b->seq, b->n_bytes, cudaMemcpyHostToDevice, s.stream));
        s.qual_in_place = false;
        if (b->qual == b->seq) s.u_qual = s.d_seq;       // raw FASTQ text: one blob, uploaded once
        else {
            if (ctx->qual_zero_copy && s.has_names) {
                // page-locked memory is mapped into the device's address space (UVA): emit_kernel reads it in place
                cudaPointerAttributes at;
                if (cudaPointerGetAttributes(&at, b->qual) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
                    s.u_qual = static_cast<uint8_t *>(at.devicePointer);
                    s.qual_in_place = true;
                } else cudaGetLastError();
            }
            if (!s.qual_in_place)
                CK(cudaMemcpyAsync(s.d_qual, b->qual, b->n_bytes, cudaMemcpyHostToDevice, s.stream));
        }
        CK(cudaMemcpyAsync(s.d_offsets, b->offsets, sizeof(uint64_t) * b->n_reads, cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemcpyAsync(s.d_lengths, b->lengths, sizeof(uint32_t) * b->n_reads, cudaMemcpyHostToDevice, s.stream));
        if (b->qual_offsets) {
            CK(cudaMemcpyAsync(s.d_qual_offsets, b->qual_offsets, sizeof(uint64_t) * b->n_reads,
                               cudaMemcpyHostToDevice, s.stream));
            s.u_qual_offsets = s.d_qual_offsets;
        }
        if (s.has_names) {
            if (names_alias) s.u_names = s.d_seq;
            else if (name_bytes)
                CK(cudaMemcpyAsync(s.d_names, b->names, name_bytes, cudaMemcpyHostToDevice, s.stream));
            CK(cudaMemcpyAsync(s.d_name_offsets, b->name_offsets,
                               sizeof(uint64_t) * (b->n_reads + (b->name_lengths ? 0 : 1)),
                               cudaMemcpyHostToDevice, s.stream));
            if (b->name_lengths) {
                CK(cudaMemcpyAsync(s.d_name_lengths, b->name_lengths, sizeof(uint32_t) * b->n_reads,
                                   cudaMemcpyHostToDevice, s.stream));
                s.u_name_lengths = s.d_name_lengths;
            }
        }
    }
    CK(cudaEventRecord(s.ev[EV_H2D], s.stream));
    s.state = SLOT_UPLOADED;
    s.did_h2d = true;
    s.fresh_upload = true;
    s.did_kernels = false;
    s.did_d2h = false;
    return ORC_OK;
}

// Synthetic reads made on the device (orc_synth.cuh), left in the slot exactly as orc_upload() would
// leave an uploaded batch.
extern "C" int orc_synth(orc_ctx *ctx, int slot, uint64_t seed, uint32_t n_reads, uint32_t len_min, uint32_t len_max)
{
    Slot *sp = get_slot(ctx, slot);
    if (!sp) return ORC_EINVAL;
    Slot &s = *sp;
    if (ctx->n_rounds != 2 || ctx->special[0] || ctx->special[1] || ctx->h_tab[0].wild || ctx->h_tab[1].wild) {
        ctx->err = "orc_synth needs two rounds of plain (ACGT, unanchored, <= 64 nt) adapters: the read model is 5' adapter + insert + 3' adapter";
        return ORC_EINVAL;
    }
    if (!ctx->emit_fastq) { ctx->err = "orc_synth needs emit_fastq (it also makes the read names)"; return ORC_EINVAL; }
    if (n_reads == 0 || n_reads > ctx->max_reads) { ctx->err = "orc_synth: n_reads exceeds max_reads"; return ORC_ECAPACITY; }
    if (len_min < 1 || len_max < len_min) { ctx->err = "orc_synth: bad length range"; return ORC_EINVAL; }
    // a first plausibility check on the mean read; the exact total is checked before anything is written
    if ((uint64_t)n_reads * (((uint64_t)len_min + len_max) / 2) > ctx->max_bytes) {
        ctx->err = "orc_synth: max_bytes too small for n_reads reads of (len_min + len_max) / 2 bases on average";
        return ORC_ECAPACITY;
    }
    if ((uint64_t)n_reads * 24 > ctx->max_name_bytes) { ctx->err = "orc_synth: max_name_bytes too small (24 per read)"; return ORC_ECAPACITY; }
    CK(cudaSetDevice(ctx->device));
    if (!ctx->d_synth) {
        SynthTable T;
        memset(&T, 0, sizeof(T));
        auto letter = [](uint8_t mask) -> uint8_t { return mask == 1 ? 0 : mask == 2 ? 1 : mask == 4 ? 2 : 3; };
        T.n5 = ctx->h_tab[0].n_adapters; T.n27 = ctx->h_tab[1].n_adapters;
        for (int a = 0; a < T.n5; a++) {
            T.m5[a] = ctx->h_tab[0].m[a];
            for (int i = 0; i < T.m5[a]; i++) T.a5[a][i] = letter(ctx->h_tab[0].code[a][i]);
        }
        for (int a = 0; a < T.n27; a++) {
            T.m27[a] = ctx->h_tab[1].m[a];
            for (int i = 0; i < T.m27[a]; i++) T.a27[a][i] = letter(ctx->h_tab[1].code[a][i]);
        }
        CK(dalloc(&ctx->d_synth, 1));
        CK(cudaMemcpy(ctx->d_synth, &T, sizeof(T), cudaMemcpyHostToDevice));
        CK(dalloc(&ctx->d_synth_totals, 2));
    }
    cudaStream_t st = s.stream;
    SynthArgs A;
    A.seed = seed; A.n_reads = n_reads; A.len_min = len_min; A.len_max = len_max;
    const uint32_t warp_blocks = (n_reads + 7) / 8;         // 8 warps per block, one warp per read
    CK(cudaEventRecord(s.ev[EV_START], st));
    synth_lengths_kernel<<<warp_blocks, 256, 0, st>>>(ctx->d_synth, A, s.d_lengths, s.d_out_len, s.d_name_lengths);
    synth_offsets_kernel<<<1, 1024, 0, st>>>(s.d_lengths, s.d_name_lengths, n_reads, s.d_offsets, s.d_name_offsets,
                                             ctx->d_synth_totals);
    CK(cudaGetLastError());
    uint64_t totals[2] = {0, 0};
    CK(cudaMemcpyAsync(totals, ctx->d_synth_totals, sizeof(totals), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (totals[0] > ctx->max_bytes || totals[1] > ctx->max_name_bytes) {
        ctx->err = "orc_synth: the generated reads exceed max_bytes / max_name_bytes";
        s.state = SLOT_IDLE;
        return ORC_ECAPACITY;
    }
    synth_write_kernel<<<warp_blocks, 256, 0, st>>>(ctx->d_synth, A, s.d_lengths, s.d_out_len, s.d_offsets,
                                                    s.d_name_offsets, s.d_seq, s.d_qual, s.d_names);
    CK(cudaGetLastError());
    s.n_reads = n_reads;
    s.n_bytes = totals[0];
    s.name_bytes = totals[1];
    s.in_bases = totals[0];
    s.max_len = len_max;
    s.has_names = true;
    s.u_qual = s.d_qual; s.u_names = s.d_names; s.u_qual_offsets = nullptr; s.u_name_lengths = nullptr;
    CK(cudaEventRecord(s.ev[EV_H2D], st));
    s.state = SLOT_UPLOADED;
    s.did_h2d = false; s.fresh_upload = false; s.did_kernels = false; s.did_d2h = false;
    return ORC_OK;
}

// Size of the batch resident in a slot (separate-blob layout), for orc_export().
extern "C" int orc_resident(orc_ctx *ctx, int slot, uint32_t *n_reads, uint64_t *n_bytes, uint64_t *name_bytes)
{
    Slot *sp = get_slot(ctx, slot);
    if (!sp) return ORC_EINVAL;
    if (sp->state == SLOT_IDLE) { ctx->err = "no batch resident in this slot"; return ORC_ESTATE; }
    if (n_reads) *n_reads = sp->n_reads;
    if (n_bytes) *n_bytes = sp->n_bytes;
    if (name_bytes) *name_bytes = sp->name_bytes;
    return ORC_OK;
}

// Copies the resident batch of a slot back to host buffers in the separate-blob layout of orc_batch
// (sizes from orc_resident(); name_offsets has n_reads + 1 entries).  For batches made by orc_synth().
extern "C" int orc_export(orc_ctx *ctx, int slot, uint8_t *seq, uint8_t *qual, uint64_t *offsets, uint32_t *lengths,
                          uint8_t *names, uint64_t *name_offsets)
{
    Slot *sp = get_slot(ctx, slot);
    if (!sp) return ORC_EINVAL;
    Slot &s = *sp;
    if (s.state == SLOT_IDLE) { ctx->err = "no batch resident in this slot"; return ORC_ESTATE; }
    if (s.u_qual != s.d_qual || s.u_name_lengths != nullptr || !s.has_names) {
        ctx->err = "orc_export handles the separate-blob layout with names only"; return ORC_EINVAL;
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = s.stream;
    if (seq) CK(cudaMemcpyAsync(seq, s.d_seq, s.n_bytes, cudaMemcpyDeviceToHost, st));
    if (qual) CK(cudaMemcpyAsync(qual, s.d_qual, s.n_bytes, cudaMemcpyDeviceToHost, st));
    if (offsets) CK(cudaMemcpyAsync(offsets, s.d_offsets, sizeof(uint64_t) * s.n_reads, cudaMemcpyDeviceToHost, st));
    if (lengths) CK(cudaMemcpyAsync(lengths, s.d_lengths, sizeof(uint32_t) * s.n_reads, cudaMemcpyDeviceToHost, st));
    if (names) CK(cudaMemcpyAsync(names, s.d_names, s.name_bytes, cudaMemcpyDeviceToHost, st));
    if (name_offsets) CK(cudaMemcpyAsync(name_offsets, s.d_name_offsets, sizeof(uint64_t) * (s.n_reads + 1), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return ORC_OK;
}

// The bins of the slot's FASTQ text as gzip members (orc_gz.cuh): sampled histogram -> the batch's Huffman code ->
// bits per chunk (prefix sums per tile) and the members' CRCs -> where tiles and members start -> the codes.
// flat: a code of 8- and 9-bit lengths instead of the histogram's (see gz_table_kernel).
static int launch_gz(orc_ctx *ctx, Slot &s, int flat, uint32_t &nl)
{
    cudaStream_t st = s.stream;
    const int nb = ctx->n_bins;
    const uint64_t *total = s.d_bin_offsets + nb;
    CK(cudaMemsetAsync(s.d_gz_hist, 0, 256 * sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(s.d_gz_member_crc, 0, (size_t)nb * sizeof(uint32_t), st));
    if (!flat) { gz_hist_kernel<<<ctx->sm_count * 2, 256, 0, st>>>(s.d_fastq, total, s.d_gz_hist); nl++; }
    gz_table_kernel<<<1, 512, 0, st>>>(s.d_gz_hist, s.d_gz_table, nb, s.d_bin_offsets, s.d_gz_chunk_base, flat); nl++;
    gz_measure_kernel<<<ctx->sm_count * 8, GZ_TILE, 0, st>>>(s.d_fastq, s.d_bin_offsets, nb, s.d_gz_chunk_base, s.d_gz_table,
                                                            s.d_gz_chunk_local, s.d_gz_tile_bits, s.d_gz_member_crc); nl++;
    gz_layout_kernel<<<1, 1024, 0, st>>>(s.d_bin_offsets, nb, s.d_gz_chunk_base, s.d_gz_table, s.d_gz_chunk_local,
                                         s.d_gz_tile_bits, s.d_gz_tile_off, s.d_gz_member_pos, s.d_gz_member_bits,
                                         s.d_gz_member_bytes, s.d_gz_offsets); nl++;
    gz_zero_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(s.d_gz_offsets, nb, ctx->gz_cap, reinterpret_cast<uint4 *>(s.d_gz)); nl++;
    gz_encode_kernel<<<ctx->sm_count * 16, 128, 0, st>>>(s.d_fastq, s.d_bin_offsets, nb, s.d_gz_chunk_base, s.d_gz_table,
                                                        s.d_gz_chunk_local, s.d_gz_tile_off, s.d_gz_member_pos,
                                                        s.d_gz_member_bits, s.d_gz_member_crc, s.d_gz_member_bytes,
                                                        s.d_gz_offsets, ctx->gz_cap, s.d_gz); nl++;
    CK(cudaGetLastError());
    return ORC_OK;
}

extern "C" int orc_launch(orc_ctx *ctx, int slot)
{
    Slot *sp = get_slot(ctx, slot);
    if (!sp) return ORC_EINVAL;
    Slot &s = *sp;
    if (s.state == SLOT_IDLE) { ctx->err = "orc_launch on a slot without an uploaded batch"; return ORC_ESTATE; }
    CK(cudaSetDevice(ctx->device));
    const uint32_t n = s.n_reads;
    uint32_t nl = 0;                             // own kernels launched (CUB's are not counted)
    uint32_t *W = s.d_codes_alloc + GUARD_WORDS;
    cudaStream_t st = s.stream;
    CK(cudaMemsetAsync(s.d_counters, 0, N_COUNTERS * sizeof(uint32_t), st));
    CK(cudaMemsetAsync(s.d_cells, 0, 6 * sizeof(unsigned long long), st));
    CK(cudaEventRecord(s.ev[EV_H2D], st));       // kernels start here (re-recorded when launched alone)
    s.did_h2d = s.fresh_upload;                  // h2d_ms is only meaningful right after an upload
    s.fresh_upload = false;
    s.did_d2h = false;
    if (n) {
        init_views_kernel<<<std::min<uint32_t>((n + 255) / 256, (uint32_t)ctx->sm_count * 8), 256, 0, st>>>(s.d_offsets, s.d_lengths, n, s.d_views[0],
                                                           sort_hist(s.d_counters, 0, 0, 0)); nl++;
        bool need_codes = false;        // anchored rounds read the ASCII bases directly
        for (int r = 0; r < ctx->n_rounds; r++) need_codes = need_codes || !ctx->anchored[r];
        if (need_codes) {
            const uint64_t n16 = (s.n_bytes + 15) / 16;
            const int pack_blocks = (int)std::min<uint64_t>((n16 + 255) / 256, (uint64_t)ctx->sm_count * 16);
            pack_kernel<<<pack_blocks, 256, 0, st>>>(s.d_seq, W, n16, ctx->d_pack_lut); nl++;
        }
        if (ctx->need_u_check) {
            // counters[7]: some read holds a U (see unify_wildcards); orc_wait() refuses the batch
            u_scan_kernel<<<std::min<uint32_t>((n + 7) / 8, (uint32_t)ctx->sm_count * 8), 256, 0, st>>>(
                s.d_seq, s.d_offsets, s.d_lengths, n, s.d_counters + 7); nl++;
        }
    }
    CK(cudaEventRecord(s.ev[EV_PACK], st));
    for (int r = 0; r < ctx->n_rounds; r++) {
        const Match *prev = r == 0 ? nullptr : s.d_match[r - 1];
        const bool filter = ctx->h_tab[r].use_filter != 0;
        const bool anch = ctx->special[r];      // none of the scan stages below
        // per round: [0] stage-2b job counter, [1] result slots, [2] band-resolver tasks, [3] wide-resolver tasks,
        // [4] stage-2a job counter, [5] pairs that passed stage 2a
        uint32_t *cnt = s.d_counters + 8 * r;
        // an event after every kernel of the round, launched or not (orc_timings.kernel_ms)
        auto mark = [&](int k) -> cudaError_t { return cudaEventRecord(s.evk[r][k + 1], st); };
        if (n) CK(cudaMemsetAsync(s.d_best_key, 0, sizeof(unsigned long long) * 2 * n, st));
        CK(cudaEventRecord(s.evk[r][0], st));
        // stage 1: (with a usable shared flank) visit the reads in order of decreasing length; exact 8-mer
        // seeds and the read-end tests mark the column windows stage 2 must look at
        if (n && !anch && filter) {
            bucket_scatter_kernel<true><<<(n + SORT_BUCKETS - 1) / SORT_BUCKETS, 256, 0, st>>>(
                s.d_views[r], prev, nullptr, n, sort_hist(s.d_counters, r, 0, 0), sort_hist(s.d_counters, r, 0, 1),
                s.d_order, nullptr); nl++;
        }
        CK(mark(ORC_K_SORT_READS));
        const bool seeded = !anch && filter && ctx->h_seed[r].on != 0;
        if (n && seeded) {
            // long reads are probed in segments, one thread each (grid.y; orc_core.cuh seed_segments)
            const dim3 seed_grid((n + 255) / 256, (unsigned)seed_segments(s.max_len));
            seed_kernel<<<seed_grid, 256, 0, st>>>(ctx->d_seed[r], W, s.d_views[r], prev, s.d_order, n,
                                                  s.d_seedwins, (size_t)2 * ctx->max_reads); nl++;
        }
        CK(mark(ORC_K_SEED));
        if (n && !anch) {
            trigger_kernel<<<(2 * n + 127) / 128, 128, 0, st>>>(ctx->d_tab[r], W, s.d_views[r], prev,
                                                               filter ? s.d_order : nullptr, n, s.d_wins,
                                                               s.d_wcols, s.d_cells + 2 + r,
                                                               seeded ? s.d_seedwins : nullptr,
                                                               sort_hist(s.d_counters, r, 1, 0), (size_t)2 * ctx->max_reads); nl++;
        }
        CK(mark(ORC_K_TRIGGER));
        if (n && !anch) {
            // order the (read, direction) items by the columns they have to scan
            bucket_scatter_kernel<false><<<(2 * n + SORT_BUCKETS - 1) / SORT_BUCKETS, 256, 0, st>>>(
                nullptr, nullptr, s.d_wcols, 2 * n, sort_hist(s.d_counters, r, 1, 0), sort_hist(s.d_counters, r, 1, 1),
                s.d_item_order, s.d_wcols_sorted); nl++;
        }
        CK(mark(ORC_K_SORT_ITEMS));
        CK(cudaEventRecord(s.ev[r == 0 ? EV_TRIG0 : EV_TRIG1], st));
        // stage 2a drops the pairs that cannot hold a candidate, stage 2b scans the rest
        const bool prefilter = !anch && ctx->h_tab[r].indels != 0;
        if (n && prefilter) {
            filter_kernel<<<ctx->filter_blocks, SCAN_THREADS, 0, st>>>(
                ctx->d_tab[r], W, s.d_views[r], s.d_wins, s.d_wcols_sorted, s.d_item_order, 2 * n, s.d_jobs, cnt,
                s.d_cells + 4 + r, getenv("ORC_NO_TRIM") ? 0 : 1); nl++;
        }
        CK(mark(ORC_K_FILTER));
        if (n && ctx->longr[r]) {
            // adapters over 64 nt: cutadapt's recurrence as it is, one thread per (read, orientation)
            long_kernel<<<std::min<uint32_t>((2 * n + 127) / 128, (uint32_t)ctx->sm_count * 8), 128, 0, st>>>(
                ctx->d_long[r], W, s.d_views[r], prev, n, s.d_results, s.d_best_key); nl++;
        } else if (n && anch) {
            // anchored adapters without indels: Hamming compare of the anchored end, no alignment
            anchored_kernel<<<std::min<uint32_t>((2 * n + 127) / 128, (uint32_t)ctx->sm_count * 16), 128, 0, st>>>(ctx->d_anch[r], s.d_seq, ctx->d_comp_lut, s.d_views[r],
                                                                prev, n, s.d_results, s.d_best_key); nl++;
        } else if (n) {
            scan_kernel<<<ctx->scan_blocks, SCAN_THREADS, 0, st>>>(
                ctx->d_tab[r], W, s.d_views[r], s.d_wins, s.d_wcols_sorted, s.d_item_order, 2 * n, s.d_results,
                s.d_tasks, s.d_best_key, cnt, prefilter ? s.d_jobs : nullptr, (uint32_t)s.cap_pairs); nl++;
        }
        CK(mark(ORC_K_SCAN));
        CK(cudaEventRecord(s.ev[r == 0 ? EV_SCAN0 : EV_SCAN1], st));
        if (n && prefilter) {           // --no-indels settles every candidate in the scan: no tasks
            // the few tasks the band resolver cannot take run beside it on the slot's side stream: their
            // kernel is all latency (one trip per warp), which the band resolver's work hides
            CK(cudaEventRecord(s.ev_fork[r], st));
            CK(cudaStreamWaitEvent(s.side, s.ev_fork[r], 0));
            CK(cudaEventRecord(s.ev_w0[r], s.side));
            resolve_kernel<<<ctx->resolve_blocks, 128, 0, s.side>>>(ctx->d_tab[r], W, s.d_views[r],
                                                                   s.d_tasks + s.cap_pairs, cnt + 3,
                                                                   s.d_results, s.d_best_key, (uint32_t)s.cap_pairs); nl++;
            CK(cudaEventRecord(s.ev_join[r], s.side));
            resolve_band_kernel<<<ctx->band_blocks[r], BAND_THREADS,
                                  band_smem_bytes(ctx->h_tab[r].n_lanes, ctx->h_tab[r].n_adapters), st>>>(
                ctx->d_tab[r], W, s.d_views[r], s.d_tasks, cnt + 2, s.d_results, s.d_best_key, (uint32_t)s.cap_pairs); nl++;
        }
        CK(mark(ORC_K_RESOLVE_BAND));
        if (n && prefilter) CK(cudaStreamWaitEvent(st, s.ev_join[r], 0));
        CK(mark(ORC_K_RESOLVE_WIDE));
        if (n) {
            SelectArgs A;
            A.type = ctx->h_tab[r].type;
            A.revcomp = ctx->h_tab[r].revcomp;
            A.action = ctx->h_tab[r].action;
            A.views_in = s.d_views[r];
            A.views_out = s.d_views[r + 1];
            A.prev = prev;
            A.out = s.d_match[r];
            A.best_key = s.d_best_key;
            A.results = s.d_results;
            A.n_reads = n;
            A.last_round = (r == ctx->n_rounds - 1);
            A.round_index = r;
            A.n_ad0 = ctx->h_tab[0].n_adapters;
            A.match0 = s.d_match[0];
            A.drop_bins = ctx->d_drop;
            A.name_offsets = s.has_names ? s.d_name_offsets : nullptr;
            A.name_lengths = s.has_names ? s.u_name_lengths : nullptr;
            A.bin = s.d_bin;
            A.out_len = s.d_out_len;
            A.rec_bytes = s.d_rec_bytes;
            A.next_bases = (r + 1 < ctx->n_rounds) ? s.d_cells + r + 1 : nullptr;
            A.next_len_hist = (r + 1 < ctx->n_rounds && !ctx->special[r + 1]) ? sort_hist(s.d_counters, r + 1, 0, 0) : nullptr;
            select_kernel<<<std::min<uint32_t>((n + 127) / 128, (uint32_t)ctx->sm_count * 16), 128, 0, st>>>(A); nl++;
        }
        CK(mark(ORC_K_SELECT));
        CK(cudaEventRecord(s.ev[r == 0 ? EV_RES0 : EV_RES1], st));
    }
    if (ctx->n_rounds == 1) {
        CK(cudaEventRecord(s.ev[EV_TRIG1], st));
        CK(cudaEventRecord(s.ev[EV_SCAN1], st));
        CK(cudaEventRecord(s.ev[EV_RES1], st));
    }
    const uint32_t n_chunks = (n + BIN_CHUNK - 1) / BIN_CHUNK;
    if (n) {
        bin_count_kernel<<<(n_chunks + 3) / 4, 128, 0, st>>>(s.d_bin, s.d_rec_bytes, n, ctx->n_bins, n_chunks,
                                                            s.d_hist_cnt, s.d_hist_bytes); nl++;
    }
    bin_scan_kernel<<<ctx->n_bins, 256, 0, st>>>(n_chunks, s.d_hist_cnt, s.d_hist_bytes, s.d_bin_counts, s.d_bin_bytes); nl++;
    bin_offsets_kernel<<<1, 32, 0, st>>>(ctx->n_bins, s.d_bin_bytes, s.d_bin_offsets); nl++;
    if (n) {
        bin_place_kernel<<<(n_chunks + 3) / 4, 128, 0, st>>>(s.d_bin, s.d_rec_bytes, n, ctx->n_bins, n_chunks,
                                                            s.d_hist_bytes, s.d_bin_offsets, s.d_dest); nl++;
    }
    CK(cudaEventRecord(s.ev[EV_BIN], st));
    if (n && s.has_names) {
        // (qualities read in place from host memory: the kernel then runs at PCIe speed; smaller grids that
        // would leave thread slots to the other slots' kernels measured slower, ORC_EMIT_ZC_BLOCKS)
        const int emit_blocks = ctx->sm_count * (s.qual_in_place ? ctx->emit_zc_blocks : 8);
        emit_kernel<<<emit_blocks, 256, 0, st>>>(s.d_seq, s.u_qual, s.u_names, s.d_name_offsets,
                                                       s.u_name_lengths, s.d_offsets, s.u_qual_offsets,
                                                       s.d_views[ctx->n_rounds], s.d_dest, n, ctx->d_comp_lut,
                                                       s.d_fastq); nl++;
    }
    CK(cudaEventRecord(s.ev[EV_EMIT], st));
    if (n && s.has_names && ctx->emit_gzip) {
        const int rc = launch_gz(ctx, s, 0, nl);
        if (rc != ORC_OK) return rc;
    }
    CK(cudaEventRecord(s.ev[EV_GZ], st));
    CK(cudaGetLastError());
    s.n_launches = nl;
    s.state = SLOT_LAUNCHED;
    s.did_kernels = true;
    return ORC_OK;
}

extern "C" int orc_download(orc_ctx *ctx, int slot)
{
    Slot *sp = get_slot(ctx, slot);
    if (!sp) return ORC_EINVAL;
    Slot &s = *sp;
    if (s.state != SLOT_LAUNCHED) { ctx->err = "orc_download before orc_launch"; return ORC_ESTATE; }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = s.stream;
    const uint32_t n = s.n_reads;
    CK(cudaMemcpyAsync(s.h_bin_counts, s.d_bin_counts, sizeof(uint64_t) * ctx->n_bins, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(s.h_bin_offsets, s.d_bin_offsets, sizeof(uint64_t) * (ctx->n_bins + 1), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(s.h_counters, s.d_counters, sizeof(uint32_t) * 16, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(s.h_cells, s.d_cells, sizeof(unsigned long long) * 6, cudaMemcpyDeviceToHost, st));
    if (ctx->emit_gzip && n && s.has_names)
        CK(cudaMemcpyAsync(s.h_gz_offsets, s.d_gz_offsets, sizeof(uint64_t) * (ctx->n_bins + 1), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(s.ev[EV_HDR], st));
    if (n) {
        CK(cudaMemcpyAsync(s.h_bin, s.d_bin, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(s.h_out_len, s.d_out_len, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
        if (ctx->want_matches)
            for (int r = 0; r < ctx->n_rounds; r++)
                CK(cudaMemcpyAsync(s.h_match[r], s.d_match[r], sizeof(Match) * n, cudaMemcpyDeviceToHost, st));
    }
    s.state = SLOT_DOWNLOADING;
    return ORC_OK;
}

// More candidate pairs than the slot's arenas hold (counters [8r + 1] count every pair, stored or not):
// the results of such a launch are incomplete.  Grow the arenas to the worst case; the caller runs the
// batch again (it is still resident).  Returns 1 if it grew them, 0 if nothing overflowed.
static int grow_pair_arenas(orc_ctx *ctx, Slot &s, const uint32_t *counters)
{
    bool over = false;
    for (int r = 0; r < ctx->n_rounds; r++) over = over || counters[8 * r + 1] > s.cap_pairs;
    if (!over) return 0;
    const size_t R = ctx->max_reads;
    const size_t worst = std::max<size_t>(R * 2 * (size_t)ctx->na_max, 2 * R);
    CK(cudaStreamSynchronize(s.stream));
    cudaFree(s.d_tasks); cudaFree(s.d_results);
    s.d_tasks = nullptr; s.d_results = nullptr;
    CK(dalloc(&s.d_tasks, 2 * worst));
    CK(dalloc(&s.d_results, worst));
    s.cap_pairs = worst;
    return 1;
}

extern "C" int orc_submit(orc_ctx *ctx, int slot, const orc_batch *batch)
{
    int rc = orc_upload(ctx, slot, batch);
    if (rc != ORC_OK) return rc;
    rc = orc_launch(ctx, slot);
    if (rc != ORC_OK) return rc;
    return orc_download(ctx, slot);
}

extern "C" int orc_sync(orc_ctx *ctx, int slot)
{
    Slot *sp = get_slot(ctx, slot);
    if (!sp) return ORC_EINVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(sp->stream));
    return ORC_OK;
}

extern "C" int orc_wait(orc_ctx *ctx, int slot, orc_result *out)
{
    Slot *sp = get_slot(ctx, slot);
    if (!sp) return ORC_EINVAL;
    Slot &s = *sp;
    if (s.state != SLOT_DOWNLOADING) { ctx->err = "orc_wait on a slot with nothing submitted"; return ORC_ESTATE; }
    CK(cudaSetDevice(ctx->device));
    // the FASTQ size is only known once the header has landed
    CK(cudaEventSynchronize(s.ev[EV_HDR]));
    {
        const int grew = grow_pair_arenas(ctx, s, s.h_counters);
        if (grew < 0) return grew;
        if (grew) {                 // rare: run the resident batch again with worst-case arenas
            s.state = SLOT_UPLOADED;
            int rc = orc_launch(ctx, slot);
            if (rc == ORC_OK) rc = orc_download(ctx, slot);
            if (rc != ORC_OK) return rc;
            CK(cudaEventSynchronize(s.ev[EV_HDR]));
        }
    }
    if (ctx->need_u_check && s.h_counters[7]) {
        ctx->err = "unsupported: a read holds U, and the adapters mix plain ACGT sequences (cutadapt compares them as "
                   "ASCII: U is not T) with IUPAC ones (compared through masks: U is T)";
        s.state = SLOT_UPLOADED;
        return ORC_EINVAL;
    }
    const bool gz = ctx->emit_gzip && s.has_names && s.n_reads;     // the bins come back as gzip members
    if (s.has_names && s.h_bin_offsets[ctx->n_bins] > ctx->fastq_cap) { ctx->err = "internal: FASTQ output exceeds its arena"; return ORC_ECAPACITY; }
    uint64_t fq = !s.has_names ? 0 : gz ? s.h_gz_offsets[ctx->n_bins] : s.h_bin_offsets[ctx->n_bins];
    if (gz && (fq > ctx->gz_cap || getenv("ORC_GZ_FORCE_FLAT"))) {
        // the sampled histogram was so far off the batch's bytes that the members outgrew the arena (the encoder wrote
        // nothing): code the batch again with 8- and 9-bit codes, which fit by construction
        uint32_t nl = 0;
        const int rc = launch_gz(ctx, s, 1, nl);
        if (rc != ORC_OK) return rc;
        CK(cudaMemcpyAsync(s.h_gz_offsets, s.d_gz_offsets, sizeof(uint64_t) * (ctx->n_bins + 1), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        fq = s.h_gz_offsets[ctx->n_bins];
        if (fq > ctx->gz_cap) { ctx->err = "internal: gzip output exceeds its arena"; return ORC_ECAPACITY; }
    }
    if (fq && !s.h_fastq) CK(halloc(&s.h_fastq, (size_t)std::max(ctx->fastq_cap, ctx->emit_gzip ? ctx->gz_cap : 0)));
    if (fq) CK(cudaMemcpyAsync(s.h_fastq, gz ? s.d_gz : s.d_fastq, fq, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaEventRecord(s.ev[EV_END], s.stream));
    CK(cudaStreamSynchronize(s.stream));
    s.did_d2h = true;
    for (int b = 0; b < ctx->n_bins; b++) ctx->total_counts[(size_t)b] += s.h_bin_counts[b];
    if (out) {
        memset(out, 0, sizeof(*out));
        out->n_reads = s.n_reads;
        out->n_bins = ctx->n_bins;
        if (ctx->want_matches)
            for (int r = 0; r < ctx->n_rounds; r++) out->matches[r] = reinterpret_cast<const orc_match *>(s.h_match[r]);
        out->bin = s.h_bin;
        out->out_len = s.h_out_len;
        out->bin_counts = s.h_bin_counts;
        out->bin_offsets = gz ? s.h_gz_offsets : s.h_bin_offsets;
        out->fastq = s.h_fastq;
        out->fastq_bytes = fq;
    }
    s.state = SLOT_UPLOADED;   // the batch is still resident: orc_launch may run it again
    return ORC_OK;
}

extern "C" int orc_get_timings(orc_ctx *ctx, int slot, orc_timings *t)
{
    Slot *sp = get_slot(ctx, slot);
    if (!sp || !t) return ORC_EINVAL;
    Slot &s = *sp;
    memset(t, 0, sizeof(*t));
    if (!s.did_kernels) { ctx->err = "no launch recorded on this slot"; return ORC_ESTATE; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(s.stream));
    auto el = [&](int a, int b, float *dst) -> cudaError_t { return cudaEventElapsedTime(dst, s.ev[a], s.ev[b]); };
    CK(el(EV_H2D, EV_PACK, &t->pack_ms));
    CK(el(EV_PACK, EV_TRIG0, &t->trigger_ms[0]));
    CK(el(EV_TRIG0, EV_SCAN0, &t->scan_ms[0]));
    CK(el(EV_SCAN0, EV_RES0, &t->resolve_ms[0]));
    CK(el(EV_RES0, EV_TRIG1, &t->trigger_ms[1]));
    CK(el(EV_TRIG1, EV_SCAN1, &t->scan_ms[1]));
    CK(el(EV_SCAN1, EV_RES1, &t->resolve_ms[1]));
    CK(el(EV_RES1, EV_BIN, &t->bin_ms));
    CK(el(EV_BIN, EV_EMIT, &t->emit_ms));
    CK(el(EV_EMIT, EV_GZ, &t->gzip_ms));
    CK(el(EV_H2D, EV_GZ, &t->total_ms));
    if (s.did_h2d) CK(el(EV_START, EV_H2D, &t->h2d_ms));
    if (s.did_d2h) CK(el(EV_GZ, EV_END, &t->d2h_ms));
    {
        auto at = [&](int e, float *dst) -> cudaError_t { return cudaEventElapsedTime(dst, ctx->ev_ref, s.ev[e]); };
        if (s.did_h2d) CK(at(EV_START, &t->timeline_ms[0]));
        CK(at(EV_H2D, &t->timeline_ms[1]));
        CK(at(EV_BIN, &t->timeline_ms[2]));
        CK(at(EV_EMIT, &t->timeline_ms[3]));
        if (s.did_d2h) CK(at(EV_END, &t->timeline_ms[4]));
    }
    // counters need a device read when the caller never downloaded
    uint32_t counters[16];
    unsigned long long cells[6];
    CK(cudaMemcpy(counters, s.d_counters, sizeof(counters), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cells, s.d_cells, sizeof(cells), cudaMemcpyDeviceToHost));
    for (int r = 0; r < ctx->n_rounds; r++)
        if (counters[8 * r + 1] > s.cap_pairs) {
            // a launch-only caller (device-resident timing) never passes orc_wait(), which would have grown
            // the arenas and run the batch again
            ctx->err = "candidate-pair arena overflow in the last launch: call orc_download()/orc_wait() once";
            return ORC_ECAPACITY;
        }
    uint64_t emit_bytes = 0;
    CK(cudaMemcpy(&emit_bytes, s.d_bin_offsets + ctx->n_bins, sizeof(uint64_t), cudaMemcpyDeviceToHost));
    t->kernel_launches = s.n_launches;
    for (int r = 0; r < ctx->n_rounds; r++) {
        t->n_tasks[r] = counters[8 * r + 2] + counters[8 * r + 3];
        t->n_candidates[r] = counters[8 * r + 1];
        const RoundTable &T = ctx->h_tab[r];
        // algorithmic cells: pairs * m * n summed over the reads entering the round
        uint64_t msum = 0;
        if (ctx->longr[r]) {
            // every cell of every pair (Ukkonen's cut-off aside)
            uint64_t ml = 0;
            for (int a = 0; a < ctx->h_long[r].n_adapters; a++) ml += (uint64_t)ctx->h_long[r].m[a];
            const uint64_t bases = (r == 0) ? s.in_bases : (uint64_t)cells[1];
            t->cells[r] = t->cells_executed[r] = (T.revcomp ? 2ull : 1ull) * ml * bases;
            continue;
        }
        if (ctx->anchored[r]) {
            // Hamming path: m characters per adapter and orientation, independent of the read length
            for (int a = 0; a < ctx->h_anch[r].n_adapters; a++) msum += (uint64_t)ctx->h_anch[r].m[a];
            t->cells[r] = t->cells_executed[r] = (T.revcomp ? 2ull : 1ull) * msum * (uint64_t)s.n_reads;
            continue;
        }
        for (int a = 0; a < T.n_adapters; a++) msum += (uint64_t)T.m[a];
        const uint64_t bases = (r == 0) ? s.in_bases : (uint64_t)cells[1];
        t->cells[r] = (T.revcomp ? 2ull : 1ull) * msum * bases;
        // cells actually updated: stage 1 (the shared flank's rows, every column, both directions) +
        // stage 2a (every adapter's block rows over the window columns) + stage 2b (the m rows of
        // the pairs that passed); without stage 2a every adapter's m rows over the window columns
        uint64_t bsum = 0;
        for (int a = 0; a < T.n_adapters; a++) bsum += (uint64_t)(T.m[a] < 32 ? T.m[a] : 32);
        // (a seeded stage 1 updates no DP cells in its main pass; the short flank scans at the read
        // ends are not counted)
        const uint64_t rows1 = ctx->h_seed[r].on ? 0ull : (uint64_t)(T.sfx_primary ? T.lcs : T.lcp);
        t->cells_executed[r] = (T.use_filter ? rows1 * (T.revcomp ? 2ull : 1ull) * bases : 0ull) +
                               (T.indels ? bsum * (uint64_t)cells[2 + r] + (uint64_t)cells[4 + r]
                                         : msum * (uint64_t)cells[2 + r]);
    }
    for (int r = 0; r < ctx->n_rounds; r++) {
        for (int k = 0; k < ORC_N_KERNELS; k++) CK(cudaEventElapsedTime(&t->kernel_ms[r][k], s.evk[r][k], s.evk[r][k + 1]));
        // (RESOLVE_WIDE: the wide resolver runs on the side stream beside the band resolver; this interval is what
        // is left of it after the band resolver has finished)
        t->window_columns[r] = (uint64_t)cells[2 + r];
        t->cells_2b[r] = (uint64_t)cells[4 + r];
        t->n_pairs_2b[r] = counters[8 * r + 5];
        t->n_tasks_wide[r] = counters[8 * r + 3];
    }
    t->pack_bytes = s.n_bytes + s.n_bytes / 2;
    t->emit_bytes = 2 * emit_bytes;    // every FASTQ byte is read once and written once
    if (ctx->emit_gzip && s.n_reads && s.has_names)
        CK(cudaMemcpy(&t->gzip_bytes, s.d_gz_offsets + ctx->n_bins, sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return ORC_OK;
}

// When the stages of the batch last waited for on a slot finished on the device, milliseconds since orc_create():
// [0] the slot's stream reached the upload, [1] H2D copies done, [2] matching and binning kernels done,
// [3] emit_kernel done, [4] gzip stage done, [5] D2H copies done.  Event queries only (no device read, unlike
// orc_get_timings): safe to call inside a pipelined submit / wait loop after orc_wait() of that slot.
extern "C" int orc_get_timeline(orc_ctx *ctx, int slot, float *out6)
{
    Slot *sp = get_slot(ctx, slot);
    if (!sp || !out6) return ORC_EINVAL;
    Slot &s = *sp;
    for (int i = 0; i < 6; i++) out6[i] = 0.0f;
    if (!s.did_kernels || !s.did_d2h) { ctx->err = "orc_get_timeline: no finished batch on this slot"; return ORC_ESTATE; }
    const int evs[6] = {EV_START, EV_H2D, EV_BIN, EV_EMIT, EV_GZ, EV_END};
    for (int i = 0; i < 6; i++) {
        if (i == 0 && !s.did_h2d) continue;
        CK(cudaEventElapsedTime(&out6[i], ctx->ev_ref, s.ev[evs[i]]));
    }
    return ORC_OK;
}

extern "C" int orc_counts(orc_ctx *ctx, uint64_t *bins)
{
    if (!ctx || !bins) return ORC_EINVAL;
    memcpy(bins, ctx->total_counts.data(), sizeof(uint64_t) * (size_t)ctx->n_bins);
    return ORC_OK;
}

// FASTQ record indexer: 4 lines per record, '@' header, '+' separator, |seq| == |qual|.
extern "C" int64_t orc_fastq_index(const uint8_t *text, uint64_t n_bytes, uint32_t max_reads, int final,
                                   uint64_t *seq_offsets, uint32_t *lengths, uint64_t *qual_offsets,
                                   uint64_t *name_offsets, uint32_t *name_lengths, uint64_t *consumed,
                                   char *err, size_t err_len)
{
    auto fail = [&](const char *what, uint64_t rec) -> int64_t {
        if (err && err_len) snprintf(err, err_len, "FASTQ record %llu: %s", (unsigned long long)rec, what);
        return ORC_EINVAL;
    };
    uint64_t pos = 0;
    int64_t n = 0;
    while (n < (int64_t)max_reads && pos < n_bytes) {
        uint64_t ls[4], le[4];      // line starts / ends (exclusive, newline and '\r' stripped)
        uint64_t p = pos;
        int got = 0;
        for (int l = 0; l < 4; l++) {
            if (p > n_bytes) break;
            const uint8_t *nl = p < n_bytes ? (const uint8_t *)memchr(text + p, '\n', n_bytes - p) : nullptr;
            uint64_t end;
            if (nl) end = (uint64_t)(nl - text);
            else if (final && l == 3 && p <= n_bytes) end = n_bytes;     // last line without newline
            else break;
            ls[l] = p;
            le[l] = (end > p && text[end - 1] == '\r') ? end - 1 : end;
            p = end + 1;
            got++;
        }
        if (got < 4) {
            if (final) {
                // no more text follows: tolerate trailing blank lines, reject anything else -- a record cut off
                // anywhere, also in the middle of its header line (got == 0), is a truncated input, as it is
                // for dnaio ("Premature end of file")
                bool blank = true;
                for (uint64_t i = pos; i < n_bytes; i++) if (text[i] != '\n' && text[i] != '\r') blank = false;
                if (!blank) return fail("truncated record at end of input", (uint64_t)n);
                pos = n_bytes;
            }
            break;
        }
        if (le[0] == ls[0] || text[ls[0]] != '@') return fail("header line does not start with '@'", (uint64_t)n);
        if (le[2] == ls[2] || text[ls[2]] != '+') return fail("third line does not start with '+'", (uint64_t)n);
        if (le[1] - ls[1] != le[3] - ls[3]) return fail("sequence and qualities differ in length", (uint64_t)n);
        if (le[1] - ls[1] > 0xffffffffull) return fail("read longer than 4 Gi bases", (uint64_t)n);
        name_offsets[n] = ls[0] + 1;
        name_lengths[n] = (uint32_t)(le[0] - ls[0] - 1);
        seq_offsets[n] = ls[1];
        lengths[n] = (uint32_t)(le[1] - ls[1]);
        qual_offsets[n] = ls[3];
        pos = p > n_bytes ? n_bytes : p;
        n++;
    }
    if (consumed) *consumed = pos;
    return n;
}

extern "C" void *orc_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

extern "C" void orc_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

extern "C" int orc_timer_start(orc_ctx *ctx, int slot)
{
    Slot *sp = get_slot(ctx, slot);
    if (!sp) return ORC_EINVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(sp->ev[EV_T0], sp->stream));
    return ORC_OK;
}

extern "C" int orc_timer_stop(orc_ctx *ctx, int slot, float *ms)
{
    Slot *sp = get_slot(ctx, slot);
    if (!sp || !ms) return ORC_EINVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(sp->ev[EV_T1], sp->stream));
    CK(cudaEventSynchronize(sp->ev[EV_T1]));
    CK(cudaEventElapsedTime(ms, sp->ev[EV_T0], sp->ev[EV_T1]));
    return ORC_OK;
}

// Device time of everything launched on ANY slot between the two calls: the streams of all slots wait for the
// begin event, and the end event waits for all of them.  For throughput measurements with several resident
// batches in flight (their kernels overlap, as they do in the submit/wait pipeline).
extern "C" int orc_span_begin(orc_ctx *ctx)
{
    if (!ctx) return ORC_EINVAL;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->ev_span[0]) {
        CK(cudaEventCreate(&ctx->ev_span[0]));
        CK(cudaEventCreate(&ctx->ev_span[1]));
        ctx->ev_span_slot.assign((size_t)ctx->n_slots, nullptr);
        for (auto &e : ctx->ev_span_slot) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    for (auto &s : ctx->slots) CK(cudaStreamSynchronize(s.stream));
    CK(cudaEventRecord(ctx->ev_span[0], ctx->slots[0].stream));
    for (size_t i = 1; i < ctx->slots.size(); i++) CK(cudaStreamWaitEvent(ctx->slots[i].stream, ctx->ev_span[0], 0));
    return ORC_OK;
}

extern "C" int orc_span_end(orc_ctx *ctx, float *ms)
{
    if (!ctx || !ms) return ORC_EINVAL;
    if (!ctx->ev_span[0]) { ctx->err = "orc_span_end without orc_span_begin"; return ORC_ESTATE; }
    CK(cudaSetDevice(ctx->device));
    for (size_t i = 1; i < ctx->slots.size(); i++) {
        CK(cudaEventRecord(ctx->ev_span_slot[i], ctx->slots[i].stream));
        CK(cudaStreamWaitEvent(ctx->slots[0].stream, ctx->ev_span_slot[i], 0));
    }
    CK(cudaEventRecord(ctx->ev_span[1], ctx->slots[0].stream));
    CK(cudaEventSynchronize(ctx->ev_span[1]));
    CK(cudaEventElapsedTime(ms, ctx->ev_span[0], ctx->ev_span[1]));
    return ORC_OK;
}

// Bandwidth of kernel loads from pinned HOST memory (what emit_kernel does with the qualities when
// orc_params.qual_zero_copy is set): every warp reads `chunk` bytes out of every `stride` bytes with 16-byte
// loads per lane, like the record copies of emit_kernel.  Returns GB/s of the bytes asked for, < 0 on failure.
__global__ void __launch_bounds__(256) hostread_probe_kernel(const uint8_t *__restrict__ src, uint64_t n_chunks,
                                                             uint32_t chunk, uint64_t stride, uint32_t *out, int dup)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t acc = 0;
    for (uint64_t c = warp; c < n_chunks; c += n_warps) {
        const uint4 *p = reinterpret_cast<const uint4 *>(src + c * stride);
        for (uint32_t q = lane; q < chunk / 16u; q += 32u) {
            const uint4 v = p[q];
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
            if (dup) {                          // every unit also by the neighbouring lane, as a realigning copy does
                const uint4 w = p[q + 1];
                acc ^= w.x + w.y + w.z + w.w;
            }
        }
    }
    if (acc == 0x12345678u) out[0] = acc;       // keeps the loads alive
}

// The same reads as bulk asynchronous copies (cp.async.bulk, the TMA's 1-D mode) into shared memory and on into
// device memory: one lane per warp issues them, two buffers per warp, an mbarrier per buffer.  ORC_PROBE_BULK=1.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256) hostread_bulk_kernel(const uint8_t *__restrict__ src, uint64_t n_chunks, uint32_t chunk,
                                                            uint64_t stride, uint8_t *__restrict__ dst)
{
    extern __shared__ __align__(128) uint8_t bulk_smem[];       // [warp][2][chunk]
    __shared__ __align__(8) uint64_t bars[8][2];
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    if (lane != 0) return;
    for (int b = 0; b < 2; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[w][b])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    uint8_t *buf[2] = {bulk_smem + (size_t)w * 2 * chunk, bulk_smem + (size_t)w * 2 * chunk + chunk};
    uint32_t phase[2] = {0, 0};
    auto issue = [&](uint64_t c, int b) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[w][b])), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(buf[b])), "l"(src + c * stride), "r"(chunk), "r"(smem_u32(&bars[w][b])) : "memory");
    };
    uint64_t c = warp;
    int b = 0;
    if (c < n_chunks) issue(c, 0);
    for (; c < n_chunks; c += n_warps, b ^= 1) {
        const uint64_t nxt = c + n_warps;
        if (nxt < n_chunks) {
            // the other buffer's store to global must have read it before the next load lands there
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            issue(nxt, b ^ 1);
        }
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(smem_u32(&bars[w][b])), "r"(phase[b]) : "memory");
        phase[b] ^= 1u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     ::"l"(dst + c * (uint64_t)chunk), "r"(smem_u32(buf[b])), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

extern "C" double orc_probe_hostread(int device, const void *host, uint64_t bytes, uint32_t chunk, uint64_t stride)
{
    const int dup = getenv("ORC_PROBE_DUP") ? 1 : 0;
    if (!host || chunk < 16 || (chunk & 15u) || stride < chunk || (stride & 15u) || bytes < stride) return -1.0;
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) {
        cudaGetLastError();
        return -2.0;                            // not pinned / not mapped
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
    uint32_t *d = nullptr;
    if (cudaMalloc(&d, 64) != cudaSuccess) return -1.0;
    const uint64_t n_chunks = bytes / stride;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    const int bulk = getenv("ORC_PROBE_BULK") ? atoi(getenv("ORC_PROBE_BULK")) : 0;
    uint8_t *d_bulk = nullptr;
    if (bulk) {
        if (cudaMalloc(&d_bulk, (size_t)n_chunks * chunk) != cudaSuccess) return -1.0;
        cudaFuncSetAttribute(hostread_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * (int)chunk);
    }
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a);
        if (bulk)       // `bulk` blocks per SM, 8 warps each, two buffers of `chunk` bytes per warp
            hostread_bulk_kernel<<<prop.multiProcessorCount * bulk, 256, 16 * (size_t)chunk>>>(
                static_cast<const uint8_t *>(at.devicePointer), n_chunks, chunk, stride, d_bulk);
        else
        hostread_probe_kernel<<<prop.multiProcessorCount * 8, 256>>>(static_cast<const uint8_t *>(at.devicePointer), n_chunks,
                                                                    chunk, stride, d, dup);
        cudaEventRecord(b);
        if (cudaEventSynchronize(b) != cudaSuccess) { best = -1.f; break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    cudaFree(d_bulk);
    if (cudaGetLastError() != cudaSuccess) return -3.0;
    if (best <= 0) return -1.0;
    return (double)n_chunks * chunk / (best * 1e-3) / 1e9;
}

extern "C" double orc_measure_int32_peak(int device, int mode, double *sm_clock_mhz)
{
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    uint32_t *d = nullptr;
    if (cudaMalloc(&d, (size_t)blocks * threads * sizeof(uint32_t)) != cudaSuccess) return -1.0;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(a);
        if (mode == 0) int32_peak_kernel<0><<<blocks, threads>>>(d, iters, 12345u + rep);
        else int32_peak_kernel<1><<<blocks, threads>>>(d, iters, 12345u + rep);
        cudaEventRecord(b);
        if (cudaEventSynchronize(b) != cudaSuccess) { best = -1.f; break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    if (best <= 0) return -1.0;
    if (sm_clock_mhz) *sm_clock_mhz = prop.clockRate / 1000.0;
    const double ops = (double)blocks * threads * (double)iters * 64.0;   // 8 unrolled x 8 ops
    return ops / (best * 1e-3);
}
