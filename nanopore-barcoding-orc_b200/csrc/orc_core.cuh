// orc_core.cuh -- per-lane arithmetic of the demultiplexer, shared by every kernel.
//
// Everything here is a pure function of its arguments (no warp intrinsics, no global
// state) and is marked __host__ __device__ so that tests/hostsim.cpp can run the very same
// code on the CPU and compare it with the oracle; the product only ever calls it from
// kernels (orc_kernels.cu).
//
// What is computed (rules R1..R10; upstream cutadapt 4.9 _align.pyx
// Aligner.locate as invoked by /root/reference/scripts/02_cutadapt_loop.sh:64-72, :94-102):
//
//  scan_lane()     Myers/Hyyro bit-parallel column scan of ONE (read view, adapter,
//                  orientation) pair.  The adapter (m <= 64 rows) is the bit-vector, left
//                  aligned so that row m is bit 63; the read is streamed 8 columns per
//                  32-bit word of 4-bit codes.  It yields the exact unit-cost cost D[m][j]
//                  of every column and D[i][n] of the last column, and from them the hull
//                  of the cells that can possibly pass cutadapt's acceptance test
//                  (R5: last row, R6: last column).  Most pairs have no such cell.
//  resolve_pair()  For a pair that has candidate cells: re-scan the few columns around them
//                  keeping the vertical deltas, walk cutadapt's own predecessor rule (R3,
//                  same tie-breaks) back from each candidate that can still win to get its
//                  (score, origin), then R5/R6/R7 verbatim.  DESIGN.md section 4 argues why
//                  this reproduces the forward recurrence exactly.
//  select_read()   R8 (best of the adapters) and R9 (--rc: strictly higher score wins),
//                  R10 (trim -> the next view of the read).
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ORC_HD __host__ __device__ __forceinline__
#else
#define ORC_HD inline
#endif
// "does any lane of the group still want to go on?" -- a warp vote on the device, which also
// brings the lanes of `mask` back together at this point; the single-lane answer on the host.
#if defined(__CUDA_ARCH__)
#define ORC_ANY(mask, p) (__any_sync((mask), (p)) != 0)
#else
#define ORC_ANY(mask, p) (p)
#endif

namespace orc {

constexpr int MAX_AD = 32;          // adapters per round
constexpr int MAX_M = 64;           // adapter length (one 64-bit word)
constexpr int MAX_LANES = 64;       // 2 orientations x MAX_AD
constexpr int INF_COST = 1 << 20;
constexpr int TYPE_FRONT = 0, TYPE_BACK = 1;

// Per-round constant tables, built on the host (orc_table.h), staged in shared memory.
// Lane l < n_adapters searches adapter l in storage direction 0 (left to right over the
// packed codes); lane n_adapters + a searches adapter a in direction 1 (right to left,
// complemented -- the complement is folded into peq).
struct RoundTable {
    int32_t n_adapters;
    int32_t type;                   // TYPE_FRONT / TYPE_BACK
    int32_t revcomp;                // --rc
    int32_t min_overlap;            // -O, not yet clamped per adapter
    int32_t n_lanes;                // 2 * n_adapters
    int32_t action;                 // 0: trim, 1: retain (select_read)
    int32_t mixed;                  // 1: the round holds adapters with AND without IUPAC wildcards (all compared through the
                                    // masks, which differs from cutadapt's ASCII comparison of the plain ones only for a read
                                    // with U: orc_api.cu refuses such a batch)
    int32_t pad_;
    int32_t m[MAX_AD];              // adapter length
    int32_t k[MAX_AD];              // int(max_error_rate * m)
    int32_t min_ov[MAX_AD];         // min(min_overlap, m)
    // kmax[a][0][L] = max cost with cost <= L * rate (fp64, R5/R6); with N wildcards in a 5' adapter the larger of
    // the two exact tables behind it -- what every pruning test reads.  kmax[a][1] / kmax[a][2]: the exact limits
    // of the last-row test (R5) and the last-column test (R6): _align.pyx takes different N counts off the overlap
    // length in the two; they differ from kmax[a][0] only for 5' adapters with N wildcards.  r5_update / r6_update
    // find them KMAX_R5 / KMAX_R6 bytes behind the row they are handed (also in compact copies of an adapter's rows).
    uint8_t kmax[MAX_AD][3][MAX_M + 8];
    // the same, 8 codes per word, for the resolver's 16-cells-at-a-time diagonal walk:
    // code4: nibble 16+q = adapter[q] (16 zero nibbles in front); rcode4: nibble q =
    // comp(adapter[m-1-q]) (zero nibbles behind) -- what a direction-1 lane compares raw codes with
    uint32_t code4[MAX_AD][12];
    uint32_t rcode4[MAX_AD][12];
    uint64_t peq[MAX_LANES / 32][16][32];   // [lane / 32][read code][lane % 32]: match bits, row i at bit 64-m+i-1,
                                    // the 64-m low padding bits always 1.  Banks of 32 lanes: one PRMT builds
                                    // the byte offset code*256 + (lane%32)*8 inside a bank
    uint64_t pv0[MAX_LANES];        // vertical deltas of column 0 (R2)
    int32_t d0[MAX_LANES];          // D[m][0]
    // shared-prefix trigger filter (stage 1 of the scan)
    int32_t use_filter;             // 1 if the adapters share a prefix long enough to filter on
    int32_t lcp;                    // its length Lp (<= 32)
    int32_t k_max, m_max;           // largest k and m of the round
    int32_t m_min;                  // shortest adapter
    int32_t indels;                 // 0: --no-indels (only diagonal moves; Hamming distance along diagonals)
    int32_t sfx_primary;            // 3' round: stage 1 scans the shared suffix, not the shared prefix
    uint32_t peq32[16][64];         // [read code][lane]: match bits of the prefix, row Lp at bit 31;
                                    // even lanes: direction 0, odd lanes: direction 1 (complemented)
    // what decides whether the mandatory first (5') / last (3') window is needed at all
    int32_t lcs;                    // length Ls of the suffix all adapters share (<= 32, 0: none)
    int32_t min_ov_min;             // smallest min_ov of the round
    uint8_t kmax_any[MAX_M + 8];    // max over the adapters of kmax[a][L]
    uint32_t peq32s[16][64];        // like peq32, for the shared suffix (5' rounds only)
    // 5' rounds: first_lim[j] = the largest cost c <= k_max with which an alignment that starts in column 0 and
    // ends in column j could be acceptable for some adapter (it aligns at most min(m_max, j + c) adapter
    // characters), -1 if there is none.  A scan whose cost in column j is D can only be looking at such an
    // alignment if D <= first_lim[j] (the alignment's cost is >= D).
    int8_t first_lim[MAX_M + 32];
    // chunk_lut[P | M << 4], P / M = the Ph / Mh top bits of four consecutive columns (oldest in
    // bit 3): low nibble = 4 + the lowest prefix sum of the four deltas, high nibble = 4 + their sum
    uint8_t chunk_lut[256];
    // Stage 2a: per lane, the match bits of a block of Lb = min(32, m) adapter rows -- the LAST
    // rows of a 5' adapter, the FIRST rows of a 3' adapter -- row Lb at bit 31, like peq32.
    uint32_t peq32b[16][64];
    int32_t block_len[MAX_AD];    // Lb, or 0: no block test for this adapter (k >= Lb)
    int32_t wild;                 // 1: the adapters hold IUPAC wildcards (reads compared through the ACGT masks, U = T)
    // the adapters as 4-bit IUPAC masks (A1 C2 G4 T8), one per byte: read by the table builders only, kept out
    // of the head that resolve_kernel stages in shared memory
    uint8_t code[MAX_AD][MAX_M];
};

// A read (or what a previous round left of it) as a window of the packed code array:
// element p of the view is code[lo+p], or, if rc is set, comp(code[lo+len-1-p]).
// the 64-bit match table of a lane's bank (scan_window / resolve_columns add code*256 + (lane%32)*8)
ORC_HD const char *peq_bank(const RoundTable &T, int lane)
{
    return reinterpret_cast<const char *>(&T.peq[lane >> 5][0][0]);
}

struct View {
    uint64_t lo;        // absolute index into the flat code / seq / qual arrays
    uint32_t len;
    uint32_t rc;        // bit 0: reversed+complemented; bits 8..: number of " rc" suffixes so far
};

// Candidate hull of one pair (output of the scan, input of the resolver).
struct Task {
    uint32_t read;
    uint32_t lane;      // adapter + n_adapters * storage direction
    int32_t jf, jl;     // columns (1-based) of the last row that may be acceptable; jf > jl: none
    int32_t i1, i2;     // rows of the last column that may be acceptable (BACK); i1 > i2: none
    uint32_t slot;      // where the pair's PairResult goes
    int32_t pad_;
};

// Aligner.locate's return value (R7) for one pair.
struct PairResult {
    int32_t has;
    int32_t ref_start, ref_stop, query_start, query_stop, score, errors;
    int32_t pad_;
};

struct Match {          // == orc_match (include/orcdemux.h)
    int32_t adapter, is_rc, ref_start, ref_stop, query_start, query_stop, score, errors;
};

ORC_HD int imin(int a, int b) { return a < b ? a : b; }
ORC_HD int imax(int a, int b) { return a > b ? a : b; }

ORC_HD uint32_t comp4(uint32_t c)   // complement of a 4-bit IUPAC mask: A<->T, C<->G
{
    return ((c & 1u) << 3) | ((c & 2u) << 1) | ((c & 4u) >> 1) | ((c & 8u) >> 3);
}

ORC_HD uint32_t nib(const uint32_t *W, int64_t idx)
{
    return (W[idx >> 3] >> ((uint32_t)(idx & 7) * 4u)) & 15u;
}

// code of element p of the sequence that a lane with storage direction `dir` sees
ORC_HD uint32_t lane_code(const uint32_t *W, uint64_t lo, uint32_t len, int dir, int p)
{
    if (!dir) return nib(W, (int64_t)lo + p);
    return comp4(nib(W, (int64_t)lo + (int64_t)len - 1 - p));
}

#if defined(__CUDA_ARCH__)
ORC_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
ORC_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
#else
ORC_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
    sh &= 31u;
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
}
ORC_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s)
{
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        uint32_t sel = (s >> (4 * i)) & 7u;
        r |= (uint32_t)((v >> (8 * sel)) & 0xffu) << (8 * i);
    }
    return r;
}
#endif

ORC_HD int popc64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

ORC_HD int clz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}

ORC_HD int ctz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return x ? __ffsll((long long)x) - 1 : 64;
#else
    return x ? __builtin_ctzll(x) : 64;
#endif
}

ORC_HD uint64_t brev64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    return __builtin_bswap64(x);
#endif
}

// 16 consecutive 4-bit codes starting at code index idx of a word array (idx may be unaligned)
ORC_HD uint64_t nib16(const uint32_t *A, int64_t idx)
{
    const int64_t w = idx >> 3;
    const uint32_t sh = (uint32_t)(idx & 7) * 4u;
    const uint32_t a = A[w], b = A[w + 1], c = A[w + 2];
    return (uint64_t)funnel_r(a, b, sh) | ((uint64_t)funnel_r(b, c, sh) << 32);
}


// distance from kmax[a] to kmax_r5[a] / kmax_r6[a]
constexpr int KMAX_R5 = MAX_M + 8, KMAX_R6 = 2 * KMAX_R5;

// cutadapt's running best match of one Aligner.locate call (R5-R7) and one DP cell.
struct Best { int32_t score, cost, origin, ref_stop, query_stop; };
struct Cell { int32_t cost, score, origin; };

// R5 update (last row).  Returns true when cutadapt stops scanning (exact full match).
ORC_HD bool r5_update(Best &b, int m, int n, const Cell &c, int j, int min_ov, const uint8_t *kmax)
{
    const int length = m + imin(c.origin, 0);
    if (!(length >= min_ov && c.cost <= (int)kmax[KMAX_R5 + length])) return false;
    const int best_length = m + imin(b.origin, 0);
    if (b.cost == m + n + 1 ||
        (c.origin <= b.origin + m / 2 && c.score > b.score) ||
        (length > best_length && c.score > b.score)) {
        b.score = c.score; b.cost = c.cost; b.origin = c.origin; b.ref_stop = m; b.query_stop = j;
        return c.cost == 0 && c.origin >= 0;
    }
    return false;
}

// R6 update (last column, row i)
ORC_HD void r6_update(Best &b, int n, const Cell &c, int i, int min_ov, const uint8_t *kmax)
{
    const int length = i + imin(c.origin, 0);
    if (!(length >= min_ov && c.cost <= (int)kmax[KMAX_R6 + length])) return;
    if (c.score > b.score || (c.score == b.score && c.cost < b.cost)) {
        b.score = c.score; b.cost = c.cost; b.origin = c.origin; b.ref_stop = i; b.query_stop = n;
    }
}

// ------------------------------------------------------------------------------------
// The scan, in two stages.
//
// Stage 1 (trigger_lane): when all adapters of a round share a prefix P of Lp <= 32 characters
// (the M13 tables do: 25 nt for SP5, 17 nt for SP27rc), rows 1..Lp of the DP matrix are the
// same for every adapter.  One 32-bit Myers scan of P per (read, direction) gives D[Lp][j] for
// every column.  An acceptable alignment of ANY adapter that starts in row 0 (or above row Lp
// in column 0) passes through row Lp at some column j' with D[Lp][j'] <= k ("trigger") and ends
// within (m - Lp) + k columns of it, so the columns that the per-adapter scan has to look at
// are windows around the triggers, plus the first columns of a 5' adapter (alignments that
// start below row Lp in column 0) and the last columns of a 3' adapter (rows <= Lp of the last
// column).  The windows are a superset of what is needed; a missed window would lose matches,
// a superfluous one only costs time.
//
// Stage 2 (scan_window): the 64-bit Myers scan of one (read, adapter, direction) pair over one
// window [s, e] of columns.  A window with s > 0 starts from column costs D[i][s] = i; that
// over-estimates, but is exact for every cell whose optimal path starts at a column >= s, and
// every candidate inside the window is of that kind because windows reach Lp + 2k + 1 columns
// back from their trigger.
//
// W: packed codes (8 per word) with guard words on both sides.  The match-bit tables live in
// shared memory on the device; the entry of (code c, lane l) is at c*256 + l*8 (64-bit table)
// or c*256 + l*4 (32-bit table), an address that one PRMT builds from a code byte and l*8.
// ------------------------------------------------------------------------------------
struct ScanHull { int32_t jf, jl, i1, i2; };

constexpr int MAX_WIN = 3;
struct alignas(16) WinList {        // windows of one (read, direction), increasing, disjoint; moved as two 16-byte words
    uint32_t n;
    uint32_t s[MAX_WIN], e[MAX_WIN];   // columns s+1 .. e are scanned; s == 0 is the true column 0
    uint32_t flags;                    // bit 0: a last-column cell (i, n) with i <= 32 may be a candidate (3' rounds)
};

ORC_HD uint32_t funnel_l1(uint32_t acc, uint32_t top)   // (acc << 1) | (top >> 31)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(top, acc, 1);
#else
    return (acc << 1) | (top >> 31);
#endif
}

ORC_HD int popc32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

// Lowest D[row][j] over the (up to eight) columns of a chunk, exactly, and the total change:
// accP / accM hold one Ph / Mh top bit per column, newest in bit 0.  Columns a short chunk does
// not have are zero deltas in the older half; they only add the chunk's starting cost to the
// minimum, which can trigger a replay that finds nothing but never hides a column.
ORC_HD int chunk_min(const uint8_t *lut, uint32_t accP, uint32_t accM, int D, int &sum)
{
    const uint32_t e1 = lut[(accP >> 4) | (accM & 0xF0u)];            // the older four columns
    const uint32_t e2 = lut[(accP & 15u) | ((accM << 4) & 0xF0u)];    // the newer four
    const int s1 = (int)(e1 >> 4) - 4;
    sum = s1 + (int)(e2 >> 4) - 4;
    return D + imin((int)(e1 & 15u), s1 + (int)(e2 & 15u)) - 4;
}

// Reads 8 codes of a lane's sequence, view positions p .. p+7, as two words of byte-wide
// codes: A holds positions p, p+2, p+4, p+6 and B holds p+1, p+3, p+5, p+7, in the byte order
// that the lane's PRMT selectors expect (reversed for direction 1).
struct ChunkReader {
    // Eight codes per step, one packed word apart: the bit offset inside the word never changes
    // and consecutive steps share a word, so a step costs one load and one funnel shift.
    const uint32_t *p;  // the word after (direction 0) / before (direction 1) the ones held
    uint32_t lo_, hi_;  // the two words the current chunk straddles
    uint32_t sh, shA, shB;
    int32_t dir_;
    ORC_HD void init(const uint32_t *W_, uint64_t lo, uint32_t len, int dir, uint32_t p0)
    {
        // storage index of the lowest-addressed code of the first chunk
        const int64_t s = dir ? (int64_t)lo + (int64_t)len - 8 - (int64_t)p0 : (int64_t)lo + (int64_t)p0;
        const uint32_t *w = W_ + (s >> 3);
        sh = (uint32_t)(s & 7) * 4u;
        shA = dir ? 4u : 0u;
        shB = dir ? 0u : 4u;
        dir_ = dir;
        if (!dir) { hi_ = w[0]; p = w + 1; }        // next(): lo_ <- hi_, hi_ <- *p++
        else      { lo_ = w[1]; p = w; }            // next(): hi_ <- lo_, lo_ <- *p--
    }
    ORC_HD void next(uint32_t &A, uint32_t &B)
    {
        const uint32_t nw = *p;
        if (!dir_) { lo_ = hi_; hi_ = nw; p += 1; }
        else       { hi_ = lo_; lo_ = nw; p -= 1; }
        const uint32_t x = funnel_r(lo_, hi_, sh);
        A = (x >> shA) & 0x0F0F0F0Fu;
        B = (x >> shB) & 0x0F0F0F0Fu;
    }
};

// The same with the next word already on its way while the current chunk is worked on (one word past the
// range is read: guard words).
struct ChunkReaderAhead {
    ChunkReader r;
    uint32_t pre;
    ORC_HD void init(const uint32_t *W_, uint64_t lo, uint32_t len, int dir, uint32_t p0)
    {
        r.init(W_, lo, len, dir, p0);
        pre = *r.p;
    }
    ORC_HD void next(uint32_t &A, uint32_t &B)
    {
        const uint32_t nw = pre;
        if (!r.dir_) { r.lo_ = r.hi_; r.hi_ = nw; r.p += 1; }
        else         { r.hi_ = r.lo_; r.lo_ = nw; r.p -= 1; }
        pre = *r.p;
        const uint32_t x = funnel_r(r.lo_, r.hi_, r.sh);
        A = (x >> r.shA) & 0x0F0F0F0Fu;
        B = (x >> r.shB) & 0x0F0F0F0Fu;
    }
};

ORC_HD void win_add(WinList &L, bool &open, uint32_t &cs, uint32_t &ce, uint32_t s, uint32_t e)
{
    if (!open) { cs = s; ce = e; open = true; return; }
    if (s <= ce + 1u || L.n >= (uint32_t)(MAX_WIN - 1)) {   // overlapping, or out of slots: merge
        if (e > ce) ce = e;
        if (s < cs) cs = s;
        return;
    }
    L.s[L.n] = cs; L.e[L.n] = ce; L.n++;
    cs = s; ce = e;
}

// ------------------------------------------------------------------------------------
// Stage 1s: exact seeds instead of the column-by-column flank scan of the main loop below.
//
// Cut every adapter into its floor(m / 8) leading pieces of 8 rows.  An alignment of the whole
// adapter (rows 1..m) with at most k errors damages at most k pieces (a substitution or deletion
// the piece of its row, an insertion the piece it falls inside), so at least np - k pieces appear
// in the read exactly and contiguously, and the diagonals of two such pieces differ by at most k
// (the indels between them).  A piece of 8 codes is one 32-bit window of the packed code array,
// so "does a piece start here" is one probe of a perfect hash table per column -- for both
// directions at once, because a piece found in the reverse complement of the read is the
// reverse-complemented piece found in the read.  Where np - k >= 2 a hit only counts together
// with a second one (another piece of the same adapter, same direction, diagonal within k):
// random 8-mers practically never come in such pairs.
//
// Every confirmed hit with diagonal c0 (row i of the adapter at column c0 + i) opens the window
// [c0 - 2k - 1, c0 + m_max + k]: the candidates end within k of c0 + m, and any optimal path to
// them starts within m + k columns before its end, i.e. at or after c0 - 2k; such a path is
// itself an alignment with <= k errors, so it has seeds of its own whose window overlaps this
// one and is merged with it -- inside the merged windows every candidate's cost is exact
// (DESIGN.md section 4, item 5b).  Alignments that do NOT span the whole adapter are the ones
// that run off an end of the read; trigger_lane() decides those windows exactly as before
// (first window of a 5' round, last windows of a 3' round).
// ------------------------------------------------------------------------------------
constexpr int SEED_SLOTS = 4096;        // 32 KB of shared memory
constexpr int SEED_SHIFT = 20;          // slot = (key * mult) >> SEED_SHIFT
constexpr int SEED_LIST_MAX = 512;
constexpr int SEED_RAW_MAX = 12;        // raw hits kept per direction before giving up (-> scan everything)
constexpr uint32_t SEED_EMPTY = 0xFFFFFFFFu;    // no packed window ever has an all-ones nibble
// A long read is probed in segments of SEED_SEG columns, one thread each, so that a batch of kilobase reads is
// several waves of equal work instead of one wave that lasts as long as its longest read.  Segment g probes
// the window positions [g * SEED_SEG, (g + 1) * SEED_SEG + m_max + kt + 8): two pieces of one occurrence lie
// at most m_max - 8 + kt positions apart, so the segment that holds the first of them sees the second as well
// (the rule that a hit needs a neighbour is applied per segment); hits in the overlap are reported by both
// segments, which only repeats a window.  The last segment (SEED_SEGS_MAX - 1) takes the rest of the read.
constexpr uint32_t SEED_SEG = 1024;
constexpr int SEED_SEGS_MAX = 4;
ORC_HD int seed_segments(uint32_t len)
{
    const uint32_t g = (len + SEED_SEG - 1u) / SEED_SEG;
    return g < 1u ? 1 : (g > (uint32_t)SEED_SEGS_MAX ? SEED_SEGS_MAX : (int)g);
}
ORC_HD void seed_range(uint32_t len, int g, int n_seg, int m_max, int kt, uint32_t &a, uint32_t &b)
{
    a = (uint32_t)g * SEED_SEG;
    b = (g == n_seg - 1) ? len : ((uint32_t)(g + 1) * SEED_SEG + (uint32_t)(m_max + kt + 8));
    if (b > len) b = len;
}


// slot of a key: the top 12 bits of key * mult (multiply-high keeps the index arithmetic off the
// ALU pipe, which the funnel shifts and compares of the probe loop already fill)
ORC_HD uint32_t seed_slot(uint32_t key, uint32_t mult)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(key * mult, 1u << (32 - SEED_SHIFT));
#else
    return (key * mult) >> SEED_SHIFT;
#endif
}

struct SeedTable {
    int32_t on;                     // 0: the round keeps the flank scan
    int32_t need;                   // hits an occurrence is guaranteed to have: 1 or 2
    uint32_t mult;
    int32_t n_list;
    int32_t kt, m_max, pad_[2];
    uint32_t key[SEED_SLOTS];       // SEED_EMPTY: free slot
    uint32_t val[SEED_SLOTS];       // first list entry | count << 16
    uint32_t list[SEED_LIST_MAX];   // piece << 16 | direction << 20
};

struct alignas(16) SeedWins {       // the seed windows of one (read, direction), increasing, disjoint; moved as 16-byte words
    uint32_t n;
    uint32_t s[MAX_WIN], e[MAX_WIN];
    uint32_t all;                   // too many hits to keep apart: this (read, direction) takes the flank scan
};

struct Quad { uint32_t w[4]; };
ORC_HD Quad load_quad(const uint32_t *__restrict__ W, int64_t q)     // words 4q .. 4q+3 (W is 16-byte aligned)
{
    Quad r;
#if defined(__CUDA_ARCH__)
    const uint4 v = *reinterpret_cast<const uint4 *>(W + 4 * q);
    r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
#else
    for (int i = 0; i < 4; i++) r.w[i] = W[4 * q + i];
#endif
    return r;
}

// keys: the SEED_SLOTS keys alone (shared memory on the device); vals / list are only read on a hit
// Probes the window positions [from, to) of the view (a segment, seed_range(); the whole read: 0, n); diagonals
// and windows are in the coordinates of the whole view and clipped to it.
ORC_HD void seed_scan(const uint32_t *__restrict__ W, uint64_t lo, uint32_t n, const uint32_t *keys,
                      const uint32_t *__restrict__ vals, uint32_t mult, const uint32_t *__restrict__ list,
                      int need, int kt, int m_max, SeedWins out[2], uint32_t from = 0u, uint32_t to = 0xFFFFFFFFu)
{
    for (int d = 0; d < 2; d++) {
        out[d].n = 0; out[d].all = 0;
        for (int i = 0; i < MAX_WIN; i++) { out[d].s[i] = 0; out[d].e[i] = 0; }
    }
    if (to > n) to = n;
    if (n < 8u || to < from + 8u) return;   // no piece fits (and no whole adapter either)
    int32_t rc0[2][SEED_RAW_MAX];   // diagonals of the hits, per direction, kept sorted
    int nraw[2] = {0, 0};
    // 32 columns per 16-byte load, the next load in flight while these are probed.  Windows that
    // start before lo or reach past lo + n (the neighbours' codes) are discarded when they hit.
    const int64_t q_first = (int64_t)((lo + from) >> 5), q_last = (int64_t)((lo + to - 8u) >> 5);
    Quad cur = load_quad(W, q_first);
    Quad nxt = load_quad(W, q_first + 1);       // at most one quad past the read: guard words
    for (int64_t q = q_first; q <= q_last; q++) {
        const Quad nn = load_quad(W, q + 2 <= q_last + 1 ? q + 2 : q_last + 1);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t a = cur.w[u], b = u < 3 ? cur.w[u + 1] : nxt.w[0];
            bool any = false;
#pragma unroll
            for (int t = 0; t < 8; t++) {
                const uint32_t x = t ? funnel_r(a, b, 4u * (uint32_t)t) : a;
                any |= (keys[seed_slot(x, mult)] == x);
            }
            if (any) {
                uint32_t mask = 0;
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const uint32_t x = t ? funnel_r(a, b, 4u * (uint32_t)t) : a;
                    mask |= (keys[seed_slot(x, mult)] == x ? 1u : 0u) << t;
                }
                while (mask) {
#if defined(__CUDA_ARCH__)
                    const int t = __ffs((int)mask) - 1;
#else
                    const int t = __builtin_ctz(mask);
#endif
                    mask &= mask - 1u;
                    const uint64_t pos = 32ull * (uint64_t)q + 8ull * (uint64_t)u + (uint64_t)t;
                    if (pos < lo + from || pos + 8u > lo + to) continue; // the piece must lie inside the probed range
                    const uint32_t x = t ? funnel_r(a, b, 4u * (uint32_t)t) : a;
                    const uint32_t val = vals[seed_slot(x, mult)];
                    const uint32_t first = val & 0xFFFFu, cnt = val >> 16;
                    for (uint32_t z = 0; z < cnt; z++) {
                        const uint32_t e = list[first + z];
                        const int d = (int)((e >> 20) & 1u);
                        const int32_t p0 = d ? (int32_t)(lo + n - 8u - pos) : (int32_t)(pos - lo);
                        const int32_t c0 = p0 - 8 * (int32_t)((e >> 16) & 15u);
                        if (nraw[d] < SEED_RAW_MAX) {
                            int i = nraw[d]++;
                            while (i > 0 && rc0[d][i - 1] > c0) { rc0[d][i] = rc0[d][i - 1]; i--; }
                            rc0[d][i] = c0;
                        } else out[d].all = 1u;
                    }
                }
            }
        }
        cur = nxt; nxt = nn;
    }
    // A hit counts if need == 1, or if another hit lies within kt diagonals of it: the two exact
    // pieces an occurrence is guaranteed to have are at most kt diagonals apart, so each of them
    // has a neighbour that close in the sorted list.  (Which pieces and adapters the neighbours
    // belong to is not checked: a superfluous window only costs time.)
    for (int d = 0; d < 2; d++) {
        if (out[d].all) continue;
        SeedWins &o = out[d];
        for (int i = 0; i < nraw[d]; i++) {
            const int32_t c = rc0[d][i];
            const bool ok = need <= 1 || (i > 0 && c - rc0[d][i - 1] <= kt) || (i + 1 < nraw[d] && rc0[d][i + 1] - c <= kt);
            if (!ok) continue;
            const int32_t ws = c - 2 * kt - 1, we = c + m_max + kt;
            const uint32_t s = ws > 0 ? (uint32_t)ws : 0u;
            const uint32_t e = we < (int32_t)n ? (uint32_t)we : n;
            if (o.n > 0 && (s <= o.e[o.n - 1] + 1u || o.n == (uint32_t)MAX_WIN)) {
                if (e > o.e[o.n - 1]) o.e[o.n - 1] = e;
            } else {
                o.s[o.n] = s; o.e[o.n] = e; o.n++;
            }
        }
    }
}

// Stage 1.  peq32_base: table of the shared prefix, entry (code, lane) at code*256 + lane*4,
// row Lp at bit 31.  kt = largest k of the round, ext = m_max - Lp + kt, back = Lp + kt + 1
// (an alignment through (Lp, j') with <= kt errors starts at a column >= j' - Lp - kt).
// `suffix_base` (5' rounds, may be nullptr) is the table of the suffix S all adapters share, Ls
// its length; kmax_any / min_ov_min are the loosest acceptance limits of the round.
ORC_HD void trigger_lane(const uint32_t *__restrict__ W, uint64_t lo, uint32_t len, int dir,
                         const char *peq32_base, int lane, int Lp, int kt, int type,
                         uint32_t ext, uint32_t back, WinList &out,
                         const char *suffix_base = nullptr, int Ls = 0,
                         const uint8_t *kmax_any = nullptr, int min_ov_min = 1, int m_max = 0, int m_min = 0,
                         int sfx_primary = 0, const int8_t *first_lim = nullptr,
                         const uint8_t *lut = nullptr, const SeedWins *seeded = nullptr,
                         int n_seg = 1, size_t seg_stride = 0)
{
    const uint32_t n = len;
    out.n = 0; out.flags = 0;
    for (int i = 0; i < MAX_WIN; i++) { out.s[i] = 0; out.e[i] = 0; }
    bool open = false;
    uint32_t cs = 0, ce = 0;
    // 3' rounds whose adapters share a suffix S longer than their shared prefix scan S instead
    // (sfx_primary): a full-length alignment ends with S fully aligned at <= k errors, so the
    // columns where S's last row is <= k mark where alignments END; the window then reaches
    // m_max + k + 1 columns back.  For the M13 tables that is the 23-nt SP27 flank, which random
    // sequence almost never approaches, against 0.2 % of the columns for the 17-nt prefix.
    // Partial adapters at the read end (no S in them) are found by a prefix scan of the last
    // m_max + 2k + 1 columns only.
    const bool sp = sfx_primary != 0 && type == TYPE_BACK && suffix_base != nullptr;
    const char *main_base = sp ? suffix_base : peq32_base;
    const int Lmain = sp ? Ls : Lp;
    const uint32_t pad = (Lmain == 32) ? 0u : ((1u << (32 - Lmain)) - 1u);
    uint32_t Pv, Mv = 0;
    int D;
    const uint32_t lane4 = (uint32_t)lane * 4u;
    uint32_t sel0, sel1, sel2, sel3;
    if (!dir) { sel0 = 0x5504u; sel1 = 0x5514u; sel2 = 0x5524u; sel3 = 0x5534u; }
    else      { sel0 = 0x5534u; sel1 = 0x5524u; sel2 = 0x5514u; sel3 = 0x5504u; }
    // The prefix scan itself never uses column 0 for free (column costs i, also for 5' adapters):
    // it only has to find alignments that start in row 0, which cost the same in either matrix.
    Pv = ~pad; D = Lmain;
    if (type == TYPE_FRONT) {
        // Alignments of a 5' adapter that start in column 0 (origin < 0: the read begins inside
        // the adapter) end within the first m + k columns and end with (a suffix of) the suffix S
        // all adapters share, so the last row of S's own matrix (column 0 free, REF_START) must
        // pass the candidate condition at the column where they end.  No such column -> no such
        // alignment for any adapter -> the window [0, m_max + k] is not needed.
        const uint32_t wmax = (uint32_t)(m_max + kt);
        const uint32_t w0 = wmax < n ? wmax : n;
        bool need = true;
        if (suffix_base != nullptr && Ls > 0 && first_lim != nullptr) {
            need = false;
            uint32_t sPv = 0, sMv = 0;
            int sD = 0;
            ChunkReader rs;
            rs.init(W, lo, len, dir, 0u);
            for (uint32_t c0 = 0; c0 < w0 && !need; c0 += 8) {
                uint32_t A, B;
                rs.next(A, B);
                const int ncol = imin(8, (int)(w0 - c0));
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    if (t < ncol) {
                        const uint32_t src = (t & 1) ? B : A;
                        const uint32_t sel = (t >> 1) == 0 ? sel0 : (t >> 1) == 1 ? sel1 : (t >> 1) == 2 ? sel2 : sel3;
                        const uint32_t Eq = *reinterpret_cast<const uint32_t *>(suffix_base + byte_perm(src, lane4, sel));
                        const uint32_t Xv = Eq | sMv;
                        const uint32_t Xh = (((Eq & sPv) + sPv) ^ sPv) | Eq;
                        uint32_t Ph = sMv | ~(Xh | sPv);
                        uint32_t Mh = sPv & Xh;
                        sD += (int)(Ph >> 31) - (int)(Mh >> 31);
                        Ph <<= 1; Mh <<= 1;
                        sPv = Mh | ~(Xv | Ph);
                        sMv = Ph & Xv;
                        // an alignment ending here has cost c >= sD and aligns at most
                        // min(m_max, j + c) adapter characters: first_lim[j] is the largest cost that passes
                        if (sD <= (int)first_lim[c0 + t + 1]) need = true;
                    }
                }
            }
        }
        if (need) win_add(out, open, cs, ce, 0u, w0);          // alignments starting in column 0
    }
    // Triggers are collected into clusters (runs with gaps <= 2k).  A cluster of a 3' round is
    // confirmed before it opens a window: a full-length alignment through (Lp, j') ends between
    // j' + (m_min - Lp) - k and j' + (m_max - Lp) + k with the suffix S all adapters share fully
    // aligned, i.e. S's own matrix has a last-row cost <= k at that end column.  One short scan of
    // S over that range decides; random hits of the 17-nt SP27 prefix (0.2 % of the columns) almost
    // never survive it.  A cluster whose window would reach the read end is kept unconditionally
    // (partial adapters at the 3' end do not contain S).
    // The confirming scans run after the main loop, cluster by cluster, so that the lanes of a warp
    // (32 different reads) do them at the same time; during the main loop clusters are only noted.
    const bool confirm = !sp && type == TYPE_BACK && suffix_base != nullptr && Ls > 0;
    constexpr int MAX_CLU = 12;
    uint32_t clu_fs[MAX_CLU], clu_ls[MAX_CLU];
    int n_clu = 0;
    bool noted_all = true;              // false once a cluster had to be added without being noted
    uint32_t clu_f = 0, clu_l = 0;      // first / last trigger column of the open cluster (0: none)
    auto add_cluster = [&](uint32_t f, uint32_t l, bool check) {
        const uint32_t wl_s = f > back ? f - back : 0u;
        const uint32_t wl_e = (l + ext) < n ? (l + ext) : n;
        bool ok = true;
        const uint32_t hi = l + (uint32_t)(m_max - Lp + kt);
        if (check && hi < n) {
            ok = false;
            const int lo_end = imax(1, (int)f + (m_min - Lp) - kt);
            const uint32_t s0 = (uint32_t)imax(0, lo_end - Ls - kt - 1);
            const uint32_t spad = (Ls == 32) ? 0u : ((1u << (32 - Ls)) - 1u);
            uint32_t sPv = ~spad, sMv = 0;
            int sD = Ls;
            ChunkReader rs;
            rs.init(W, lo, len, dir, s0);
            for (uint32_t c0 = s0; c0 < hi && !ok; c0 += 8) {
                uint32_t A, B;
                rs.next(A, B);
                const int ncol = imin(8, (int)(hi - c0));
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    if (t < ncol) {
                        const uint32_t src = (t & 1) ? B : A;
                        const uint32_t sel = (t >> 1) == 0 ? sel0 : (t >> 1) == 1 ? sel1 : (t >> 1) == 2 ? sel2 : sel3;
                        const uint32_t Eq = *reinterpret_cast<const uint32_t *>(suffix_base + byte_perm(src, lane4, sel));
                        const uint32_t Xv = Eq | sMv;
                        const uint32_t Xh = (((Eq & sPv) + sPv) ^ sPv) | Eq;
                        uint32_t Ph = sMv | ~(Xh | sPv);
                        uint32_t Mh = sPv & Xh;
                        sD += (int)(Ph >> 31) - (int)(Mh >> 31);
                        Ph <<= 1; Mh <<= 1;
                        sPv = Mh | ~(Xv | Ph);
                        sMv = Ph & Xv;
                        if (sD <= kt && (int)c0 + t + 1 >= lo_end) ok = true;
                    }
                }
            }
        }
        if (ok) win_add(out, open, cs, ce, wl_s, wl_e);
    };
    auto flush_cluster = [&]() {
        if (clu_f == 0) return;
        if (sp) {                       // triggers are END columns: the window reaches back m_max + k + 1
            const uint32_t rb = (uint32_t)(m_max + kt + 1);
            win_add(out, open, cs, ce, clu_f > rb ? clu_f - rb : 0u, clu_l);
        } else if (!confirm) add_cluster(clu_f, clu_l, false);
        else if (noted_all && n_clu < MAX_CLU) { clu_fs[n_clu] = clu_f; clu_ls[n_clu] = clu_l; n_clu++; }
        else {
            // out of slots: from here on clusters open their windows unconfirmed; the noted ones are
            // earlier in the read and must be added first to keep the windows in order
            for (int c = 0; c < n_clu; c++) add_cluster(clu_fs[c], clu_ls[c], false);
            n_clu = 0;
            noted_all = false;
            add_cluster(clu_f, clu_l, false);
        }
        clu_f = 0;
    };
    // Stage 1s: whole-adapter alignments were found through their seeds (seed_scan); only the
    // windows of the alignments that run off an end of the read are decided here
    // (a read with more key matches than seed_scan keeps apart -- tens of kilobases of sequence --
    // takes the flank scan below instead)
    // (segment g of the read at seeded[g * seg_stride]; in direction 1 the columns run the other way, so the
    // segments come in descending order there)
    bool seeds = seeded != nullptr;
    for (int g = 0; seeds && g < n_seg; g++) seeds = seeded[(size_t)g * seg_stride].all == 0u;
    if (seeds)
        for (int q = 0; q < n_seg; q++) {
            const SeedWins &sw = seeded[(size_t)(dir ? n_seg - 1 - q : q) * seg_stride];
            for (uint32_t i = 0; i < sw.n; i++) win_add(out, open, cs, ce, sw.s[i], sw.e[i]);
        }
    ChunkReader rd;
    rd.init(W, lo, len, dir, 0u);
    const int nchunks = seeds ? 0 : (int)((n + 7u) >> 3);
    for (int q = 0; q < nchunks; q++) {
        uint32_t A, B;
        rd.next(A, B);
        const int ncol = imin(8, (int)n - 8 * q);
        // Row 0 costs 0 in every column, so D[Lmain][j] is the sum of the column's vertical deltas:
        // popc(Pv) - popc(Mv) (the padding rows below bit 32 - Lmain keep both bits clear).  The
        // eight columns of a chunk therefore run without any bookkeeping.  The cost moves by at most
        // one per column, so between two columns c apart with costs Da and Db it stays above
        // (Da + Db - c) / 2: only a (half-)chunk where that can reach the threshold is run again,
        // column by column, from the state saved at its start.
        const uint32_t Pv0 = Pv, Mv0 = Mv;
        uint32_t Pv4 = 0, Mv4 = 0;
        auto column = [&](int t) {
            const uint32_t src = (t & 1) ? B : A;
            const uint32_t sel = (t >> 1) == 0 ? sel0 : (t >> 1) == 1 ? sel1 : (t >> 1) == 2 ? sel2 : sel3;
            const uint32_t Eq = *reinterpret_cast<const uint32_t *>(main_base + byte_perm(src, lane4, sel));
            const uint32_t Xv = Eq | Mv;
            const uint32_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
            uint32_t Ph = Mv | ~(Xh | Pv);
            uint32_t Mh = Pv & Xh;
            Ph <<= 1; Mh <<= 1;
            Pv = Mh | ~(Xv | Ph);
            Mv = Ph & Xv;
        };
        if (ncol == 8) {
#pragma unroll
            for (int t = 0; t < 4; t++) column(t);
            Pv4 = Pv; Mv4 = Mv;
#pragma unroll
            for (int t = 4; t < 8; t++) column(t);
        } else {
#pragma unroll
            for (int t = 0; t < 8; t++) if (t < ncol) column(t);
        }
        const uint32_t Pv8 = Pv, Mv8 = Mv;
        const int Dend = popc32(Pv8) - popc32(Mv8);
        auto replay = [&](int t0, int t1, uint32_t pv, uint32_t mv, int d) {
#pragma unroll 1
            for (int t = t0; t < t1; t++) {
                const uint32_t src = (t & 1) ? B : A;
                const uint32_t sel = (t >> 1) == 0 ? sel0 : (t >> 1) == 1 ? sel1 : (t >> 1) == 2 ? sel2 : sel3;
                const uint32_t Eq = *reinterpret_cast<const uint32_t *>(main_base + byte_perm(src, lane4, sel));
                const uint32_t Xv = Eq | mv;
                const uint32_t Xh = (((Eq & pv) + pv) ^ pv) | Eq;
                uint32_t Ph = mv | ~(Xh | pv);
                uint32_t Mh = pv & Xh;
                d += (int)(Ph >> 31) - (int)(Mh >> 31);
                Ph <<= 1; Mh <<= 1;
                pv = Mh | ~(Xv | Ph);
                mv = Ph & Xv;
                if (d <= kt) {
                    const uint32_t j = (uint32_t)(8 * q + t + 1);
                    if (clu_f != 0 && j - clu_l > (uint32_t)(2 * kt)) flush_cluster();
                    if (clu_f == 0) clu_f = j;
                    clu_l = j;
                }
            }
        };
        if (D + Dend <= ncol + 2 * kt) {
            if (ncol == 8) {
                const int Dmid = popc32(Pv4) - popc32(Mv4);
                if (D + Dmid <= 4 + 2 * kt) replay(0, 4, Pv0, Mv0, D);
                if (Dmid + Dend <= 4 + 2 * kt) replay(4, 8, Pv4, Mv4, Dmid);
            } else {
                replay(0, ncol, Pv0, Mv0, D);
            }
        }
        D = Dend;
    }
    flush_cluster();
    for (int c = 0; c < n_clu; c++) add_cluster(clu_fs[c], clu_ls[c], true);
    if (type == TYPE_BACK) {
        uint32_t ePv = Pv, eMv = Mv;    // prefix rows 1..Lp of the last column
        const bool restart = sp || seeds;   // the main loop did not scan the prefix: do its last columns here
        bool r6_near = !restart;        // prefix-primary rounds: not tracked, assume yes
        if (restart) {
            // prefix scan of the last m_max + 2k + 1 columns (restart: exact for every alignment
            // that can still reach the read end as a partial adapter)
            const uint32_t T = (uint32_t)(m_max + 2 * kt + 1);
            const uint32_t s1 = n > T ? n - T : 0u;
            const uint32_t jmin = n > (uint32_t)(m_max - Lp + kt) ? n - (uint32_t)(m_max - Lp + kt) : 0u;
            const uint32_t ppad = (Lp == 32) ? 0u : ((1u << (32 - Lp)) - 1u);
            ePv = ~ppad; eMv = 0;
            int eD = Lp;
            uint32_t first_trig = 0;
            // a cell (i, n) with Lp < i <= 32 is reached through row Lp within 32 - Lp + k columns of n
            const uint32_t jr6 = n > (uint32_t)(32 - Lp + kt) ? n - (uint32_t)(32 - Lp + kt) : 0u;
            ChunkReader re;
            re.init(W, lo, len, dir, s1);
            for (uint32_t c0 = s1; c0 < n; c0 += 8) {
                uint32_t A, B;
                re.next(A, B);
                const int ncol = imin(8, (int)(n - c0));
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    if (t < ncol) {
                        const uint32_t src = (t & 1) ? B : A;
                        const uint32_t sel = (t >> 1) == 0 ? sel0 : (t >> 1) == 1 ? sel1 : (t >> 1) == 2 ? sel2 : sel3;
                        const uint32_t Eq = *reinterpret_cast<const uint32_t *>(peq32_base + byte_perm(src, lane4, sel));
                        const uint32_t Xv = Eq | eMv;
                        const uint32_t Xh = (((Eq & ePv) + ePv) ^ ePv) | Eq;
                        uint32_t Ph = eMv | ~(Xh | ePv);
                        uint32_t Mh = ePv & Xh;
                        eD += (int)(Ph >> 31) - (int)(Mh >> 31);
                        Ph <<= 1; Mh <<= 1;
                        ePv = Mh | ~(Xv | Ph);
                        eMv = Ph & Xv;
                        const uint32_t j = c0 + (uint32_t)t + 1u;
                        if (eD <= kt && j >= jmin && first_trig == 0) first_trig = j;
                        if (eD <= kt && j >= jr6) r6_near = true;
                    }
                }
            }
            if (first_trig != 0) {
                const uint32_t rb = (uint32_t)(Lp + kt + 1);
                win_add(out, open, cs, ce, first_trig > rb ? first_trig - rb : 0u, n);
            }
        }
        // The last window is for the cells (i, n) with i <= Lp: an adapter prefix of at most Lp
        // characters at the read end.  Those rows are the same for every adapter and the prefix
        // scan has them exactly: if none of them passes R6's necessary condition under the
        // loosest limits of the round, no adapter has such a candidate and the window is moot.
        bool need = kmax_any == nullptr;
        if (!need) {
            int cum = 0;
            for (int i = 1; i <= Lp; i++) {
                const int bit = 32 - Lp + i - 1;
                cum += (int)((ePv >> bit) & 1u) - (int)((eMv >> bit) & 1u);
                if (i >= min_ov_min && cum <= (int)kmax_any[i]) need = true;
            }
        }
        if (need) {
            const uint32_t r = (uint32_t)(Lp + kt + 1);
            win_add(out, open, cs, ce, n > r ? n - r : 0u, n);
        }
        // stage 2a looks at the last-column rows i <= 32 of each adapter only if one of them can be
        // a candidate: a shared row (need), or a row further down reached through a prefix trigger
        // close to the read end
        if (need || r6_near) out.flags |= 1u;
    }
    if (open) { out.s[out.n] = cs; out.e[out.n] = ce; out.n++; }
}

ORC_HD uint32_t win_columns(const WinList &w)
{
    uint32_t c = 0;
    for (uint32_t i = 0; i < w.n; i++) c += w.e[i] - w.s[i];
    return c;
}

// ------------------------------------------------------------------------------------
// Stage 2a (block_test): may this (read, direction, adapter) have a candidate cell inside the
// windows at all?  A 32-bit Myers scan of a block B of Lb = min(32, m) adapter rows over the
// window columns answers with a necessary condition, at well under half the price of the
// 64-bit scan; only the pairs that pass go on to scan_window.  The adapters of a round differ
// in a few rows only, so for the reads that carry one of them the other adapters stop here.
//
//  * 5' adapter, B = its last Lb rows.  An acceptable alignment ending in (m, j) with cost c
//    spends at most c on B, and B's own matrix (row 0 free) has D_B[Lb][j] <= c <= k in that
//    same column.  If the window starts at the true column 0 the alignment may begin inside B
//    (the read starts within the adapter): B's column 0 is free as well (cost 0 in every row),
//    and for the first columns, where nearly any cost is "<= k", the acceptance limit of the
//    alignment's largest possible length decides instead (first_lim, the loosest of the round).
//  * 3' adapter, B = its first Lb rows, the same rows of the same matrix (same start costs).  A
//    path to (m, j) or to a last-column cell (i, n) with i > Lb crosses row Lb inside the window
//    at cost <= k; a last-column cell with i <= Lb is a cell of B's matrix itself and is tested
//    against R6's necessary condition directly.
//
// Away from the true column 0 the columns run without bookkeeping, as in stage 1: the cost is a
// popcount away and moves by at most one per column, so (Da + Db - c) / 2 bounds it between two
// columns c apart.  Passing a pair that has no candidate only costs time.
//
// The block does not need the whole window either.  Every window that starts after column 0 starts at least
// kt + 1 columns before the earliest alignment it was opened for begins (seeds: 2kt + 1 before the seed's
// diagonal; flank triggers: Lp + kt + 1 before the trigger), so a 5' candidate (m, j) lies at j >= s + m + 1
// and its part on B -- the last Lb rows -- begins at or after j - Lb - k: the first m - Lb - kt columns of
// the window cannot matter to B (a restarted scan is exact for what begins after its start).  Likewise every
// window that ends before the read does ends at least m_max - Lp + kt columns after its last trigger, and a
// path of a 3' adapter has crossed row Lb by then with m - Lb - kt columns to spare.  `trim` = m - Lb - kt.
ORC_HD bool block_test(const uint32_t *__restrict__ W, uint64_t lo, uint32_t len, int dir,
                       const WinList *wl, const char *peq32b_base, int lane, int Lb, int k, int type,
                       const uint8_t *kmax, int min_ov, const int8_t *first_lim, int trim = 0)
{
    if (Lb <= 0) return true;
    const uint32_t n = len;
    const uint32_t pad = (Lb == 32) ? 0u : ((1u << (32 - Lb)) - 1u);
    const uint32_t lane4 = (uint32_t)lane * 4u;
    uint32_t sel0, sel1, sel2, sel3;
    if (!dir) { sel0 = 0x5504u; sel1 = 0x5514u; sel2 = 0x5524u; sel3 = 0x5534u; }
    else      { sel0 = 0x5534u; sel1 = 0x5524u; sel2 = 0x5514u; sel3 = 0x5504u; }
    const uint32_t nw = wl ? wl->n : 1u;
    for (uint32_t w = 0; w < nw; w++) {
        uint32_t s = wl ? wl->s[w] : 0u, e = wl ? wl->e[w] : n;
        if (trim > 0) {
            if (type == TYPE_FRONT) { if (s > 0u) s = s + (uint32_t)trim < e ? s + (uint32_t)trim : e; }
            else if (e < n) e = e > s + (uint32_t)trim ? e - (uint32_t)trim : s;
        }
        const bool free0 = s == 0u && type == TYPE_FRONT;
        uint32_t Pv, Mv = 0;
        int D;
        if (free0) { Pv = 0; D = 0; }
        else { Pv = ~pad; D = Lb; }
        ChunkReader rd;
        rd.init(W, lo, len, dir, s);
        const int ncols = (int)(e - s);
        const int nchunks = (ncols + 7) >> 3;
        for (int q = 0; q < nchunks; q++) {
            uint32_t A, B;
            rd.next(A, B);
            const int ncol = imin(8, ncols - 8 * q);
            if (free0 && q < 2) {
                // the first sixteen columns after a free column 0: cost and limit column by column
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    if (t < ncol) {
                        const uint32_t src = (t & 1) ? B : A;
                        const uint32_t sel = (t >> 1) == 0 ? sel0 : (t >> 1) == 1 ? sel1 : (t >> 1) == 2 ? sel2 : sel3;
                        const uint32_t Eq = *reinterpret_cast<const uint32_t *>(peq32b_base + byte_perm(src, lane4, sel));
                        const uint32_t Xv = Eq | Mv;
                        const uint32_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
                        uint32_t Ph = Mv | ~(Xh | Pv);
                        uint32_t Mh = Pv & Xh;
                        D += (int)(Ph >> 31) - (int)(Mh >> 31);
                        Ph <<= 1; Mh <<= 1;
                        Pv = Mh | ~(Xv | Ph);
                        Mv = Ph & Xv;
                        if (D <= imin(k, (int)first_lim[8 * q + t + 1])) return true;
                    }
                }
                continue;
            }
            uint32_t Pv4 = 0, Mv4 = 0;
            auto column = [&](int t) {
                const uint32_t src = (t & 1) ? B : A;
                const uint32_t sel = (t >> 1) == 0 ? sel0 : (t >> 1) == 1 ? sel1 : (t >> 1) == 2 ? sel2 : sel3;
                const uint32_t Eq = *reinterpret_cast<const uint32_t *>(peq32b_base + byte_perm(src, lane4, sel));
                const uint32_t Xv = Eq | Mv;
                const uint32_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
                uint32_t Ph = Mv | ~(Xh | Pv);
                uint32_t Mh = Pv & Xh;
                Ph <<= 1; Mh <<= 1;
                Pv = Mh | ~(Xv | Ph);
                Mv = Ph & Xv;
            };
            if (ncol == 8) {
#pragma unroll
                for (int t = 0; t < 4; t++) column(t);
                Pv4 = Pv; Mv4 = Mv;
#pragma unroll
                for (int t = 4; t < 8; t++) column(t);
            } else {
#pragma unroll
                for (int t = 0; t < 8; t++) if (t < ncol) column(t);
            }
            // the free column 0 leaves set bits of Pv / Mv only in the block's own rows as well
            const int Dend = popc32(Pv) - popc32(Mv);
            if (D + Dend <= ncol + 2 * k) {
                if (ncol < 8) return true;
                const int Dmid = popc32(Pv4) - popc32(Mv4);
                if (D + Dmid <= 4 + 2 * k || Dmid + Dend <= 4 + 2 * k) return true;
            }
            D = Dend;
        }
        if (type == TYPE_BACK && e == n && (wl == nullptr || (wl->flags & 1u))) {
            // R6's necessary condition for the cells (i, n), i <= Lb (as in scan_window)
            int cum = 0;
            for (int i = 1; i <= Lb; i++) {
                const int bit = 32 - Lb + i - 1;
                cum += (int)((Pv >> bit) & 1u) - (int)((Mv >> bit) & 1u);
                if (i >= min_ov && cum <= (int)kmax[i]) return true;
            }
        }
    }
    return false;
}

// Per-pair state of stage 2.  Candidates of cost 0 are settled on the spot: their path is a
// pure diagonal of matches (a zero-cost cell has equal characters and a zero-cost diagonal
// predecessor, all the way to row 0 or column 0), so score = min(j, m) and origin = j - m.
// Only pairs that meet a candidate of cost > 0 (`need`) go to the resolver, which redoes the
// whole pair from its hull.
struct LaneScan {
    ScanHull h;
    Best best;
    int32_t need;       // a candidate with cost > 0 was seen: the resolver decides
    int32_t broke;      // R5 hit its early exit (exact full match) while settling inline
    int32_t c5, c6;     // D[m][jf] and D[i1][n]: the costs the band resolver counts on from
};

// Task.pad_: bit 0 = TASK_WIDE, bits 8..13 = D[m][jf], bits 16..21 = D[i1][n]
ORC_HD int32_t task_anchors(const LaneScan &L) { return (int32_t)(((uint32_t)L.c5 & 63u) << 8 | ((uint32_t)L.c6 & 63u) << 16); }

ORC_HD void lane_scan_init(LaneScan &L, int m, int n)
{
    L.h.jf = 0x7fffffff; L.h.jl = -1; L.h.i1 = 0x7fffffff; L.h.i2 = -1;
    L.best.ref_stop = m; L.best.query_stop = n; L.best.cost = m + n + 1; L.best.origin = 0; L.best.score = 0;
    L.need = 0; L.broke = 0; L.c5 = 0; L.c6 = 0;
}

// Stage 2: columns s+1 .. e of one pair.
ORC_HD void scan_window(const uint32_t *__restrict__ W, uint64_t lo, uint32_t len, int dir,
                        uint32_t s, uint32_t e, const char *peq_base, int lane, uint64_t pv0, int d0,
                        int m, int k, const uint8_t *kmax, int min_ov, int type, LaneScan &L,
                        const uint8_t *lut)
{
    const uint64_t pad = (m == 64) ? 0ull : ((1ull << (64 - m)) - 1ull);
    const int n = (int)len;
    uint64_t Pv, Mv = 0;
    int D;
    if (s == 0) { Pv = pv0; D = d0; }            // R2: the true column 0
    else { Pv = ~pad; D = m; }                   // restart: cost i
    const uint32_t lane8 = (uint32_t)(lane & 31) * 8u;           // peq_base is the lane's bank (peq_bank())
    // PRMT selectors: result byte0 <- lane8.byte0, byte1 <- code byte b, bytes 2,3 <- 0
    uint32_t sel0, sel1, sel2, sel3;
    if (!dir) { sel0 = 0x5504u; sel1 = 0x5514u; sel2 = 0x5524u; sel3 = 0x5534u; }
    else      { sel0 = 0x5534u; sel1 = 0x5524u; sel2 = 0x5514u; sel3 = 0x5504u; }
    ChunkReader rd;
    rd.init(W, lo, len, dir, s);
    const int ncols = (int)(e - s);
    const int nchunks = (ncols + 7) >> 3;
    int last_d = -1;                             // D[m][n] if column n was a candidate
    for (int q = 0; q < nchunks; q++) {
        uint32_t A, B;
        rd.next(A, B);
        const int ncol = imin(8, ncols - 8 * q);
        uint32_t accP = 0, accM = 0;
        auto column = [&](int t) {
            const uint32_t src = (t & 1) ? B : A;
            const uint32_t sel = (t >> 1) == 0 ? sel0 : (t >> 1) == 1 ? sel1 : (t >> 1) == 2 ? sel2 : sel3;
            const uint64_t Eq = *reinterpret_cast<const uint64_t *>(peq_base + byte_perm(src, lane8, sel));
            // Myers 1999 / Hyyro 2003, one column; row 0 never changes (QUERY_START): shift in 0
            const uint64_t Xv = Eq | Mv;
            const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
            uint64_t Ph = Mv | ~(Xh | Pv);
            uint64_t Mh = Pv & Xh;
            accP = funnel_l1(accP, (uint32_t)(Ph >> 32));
            accM = funnel_l1(accM, (uint32_t)(Mh >> 32));
            Ph <<= 1; Mh <<= 1;
            Pv = Mh | ~(Xv | Ph);
            Mv = Ph & Xv;
        };
        if (ncol == 8) {                         // full chunk: no per-column bound test
#pragma unroll
            for (int t = 0; t < 8; t++) column(t);
        } else {
#pragma unroll
            for (int t = 0; t < 8; t++) if (t < ncol) column(t);
        }
        // D[m][j] can only fall by the M bits; a chunk that might reach k is tested exactly, and
        // replayed column by column only if it does
        bool replay = false;
        if (D - popc32(accM) <= k) { int dsum; replay = chunk_min(lut, accP, accM, D, dsum) <= k; }
        if (replay) {
            for (int t = 0; t < ncol; t++) {
                const int b = ncol - 1 - t;
                D += (int)((accP >> b) & 1u) - (int)((accM >> b) & 1u);
                if (D <= k) {
                    // R5 necessary condition: a path with D errors ending in (m, j) aligns at
                    // most min(m, j + D) adapter characters.
                    const int j = (int)s + 8 * q + t + 1;
                    const int lmax = imin(m, j + D);
                    if (lmax >= min_ov && D <= (int)kmax[lmax]) {
                        if (j < L.h.jf) { L.h.jf = j; L.c5 = D; }
                        L.h.jl = j;
                        if (j == n) last_d = D;
                        if (D == 0) {
                            if (!L.need && !L.broke) {
                                Cell c;
                                c.cost = 0; c.score = imin(j, m); c.origin = j - m;
                                if (r5_update(L.best, m, n, c, j, min_ov, kmax)) L.broke = 1;
                            }
                        } else {
                            L.need = 1;
                        }
                    }
                }
            }
        } else {
            D += popc32(accP) - popc32(accM);
        }
    }
    if (e != len) return;
    if (type == TYPE_BACK) {
        // R6 necessary condition for the cells (i, n): origin >= 0 for BACK, so length == i
        int cum = 0;
        uint64_t zero_rows = 0;
        for (int i = 1; i <= m; i++) {
            const int bit = 64 - m + i - 1;
            cum += (int)((Pv >> bit) & 1u) - (int)((Mv >> bit) & 1u);
            if (i >= min_ov && cum <= (int)kmax[i]) {
                if (i < L.h.i1) { L.h.i1 = i; L.c6 = cum; }
                L.h.i2 = i;
                if (cum == 0) zero_rows |= 1ull << (i - 1);
                else L.need = 1;
            }
        }
        if (!L.need && !L.broke) {
            for (int i = m; i >= 1; i--) {       // top row first, like cutadapt
                if (!((zero_rows >> (i - 1)) & 1ull)) continue;
                Cell c;
                c.cost = 0; c.score = i; c.origin = n - i;
                r6_update(L.best, n, c, i, min_ov, kmax);
            }
        }
    } else if (last_d == 0 && !L.need && !L.broke) {
        // FRONT: R6 looks at the single cell (m, n)
        Cell c;
        c.cost = 0; c.score = imin(n, m); c.origin = n - m;
        r6_update(L.best, n, c, m, min_ov, kmax);
    }
}

// Mismatches between adapter positions [a_from, a_to) and the view positions that start at
// view_start, 16 packed codes per step (see trace_back for the two code layouts).
ORC_HD int diag_mismatches(const uint32_t *W, uint64_t lo, uint32_t len, int dir,
                           const uint32_t *code4, const uint32_t *rcode4, int m,
                           int a_from, int a_to, int view_start)
{
    const int cnt = a_to - a_from;
    int matches = 0;
    if (!dir) {
        for (int o = 0; o < cnt; o += 16) {
            const uint64_t r = nib16(W, (int64_t)lo + view_start + o);
            const uint64_t a = nib16(code4, (int64_t)(16 + a_from + o));
            uint64_t x = r & a;
            x |= x >> 1; x |= x >> 2;
            x &= 0x1111111111111111ull;
            const int c = imin(16, cnt - o);
            if (c < 16) x &= (1ull << (4 * c)) - 1ull;
            matches += popc64(x);
        }
    } else {
        const int q_from = m - a_to;
        const int64_t base = (int64_t)lo + (int64_t)len - view_start - m + a_from;
        for (int o = 0; o < cnt; o += 16) {
            const uint64_t r = nib16(W, base + q_from + o);
            const uint64_t a = nib16(rcode4, (int64_t)(q_from + o));
            uint64_t x = r & a;
            x |= x >> 1; x |= x >> 2;
            x &= 0x1111111111111111ull;
            const int c = imin(16, cnt - o);
            if (c < 16) x &= (1ull << (4 * c)) - 1ull;
            matches += popc64(x);
        }
    }
    return cnt - matches;
}

// --no-indels (cutadapt prices indels at 100000): within k errors only diagonal moves remain,
// so D[m][j] is the number of mismatches along the diagonal that ends in (m, j) -- from row 0
// (origin j - m >= 0) or, for a 5' adapter, from column 0 (origin j - m < 0, the adapter's last
// j characters) -- score = matches - mismatches and origin = j - m, all in closed form: every
// candidate is settled here and no pair of such a round ever needs the resolver.
ORC_HD void scan_window_noindel(const uint32_t *__restrict__ W, uint64_t lo, uint32_t len, int dir,
                                uint32_t s, uint32_t e, const uint32_t *code4, const uint32_t *rcode4,
                                int m, int k, const uint8_t *kmax, int min_ov, int type, LaneScan &L)
{
    const int n = (int)len;
    bool last_cand = false;
    Cell last;
    last.cost = last.score = last.origin = 0;
    for (int j = (int)s + 1; j <= (int)e; j++) {
        if (j < m && type == TYPE_BACK) continue;          // column 0 of a 3' adapter costs 100000 * i
        const int ov = imin(j, m);
        const int mis = diag_mismatches(W, lo, len, dir, code4, rcode4, m, m - ov, m, j - ov);
        if (mis <= k && ov >= min_ov && mis <= (int)kmax[ov]) {
            L.h.jf = imin(L.h.jf, j);
            L.h.jl = j;
            Cell c;
            c.cost = mis; c.score = ov - 2 * mis; c.origin = j - m;
            if (j == n) { last_cand = true; last = c; }
            if (!L.broke && r5_update(L.best, m, n, c, j, min_ov, kmax)) L.broke = 1;
        }
    }
    if (e != len || L.broke) return;
    if (type == TYPE_BACK) {
        for (int i = imin(m, n); i >= 1; i--) {            // top row first, like cutadapt
            const int mis = diag_mismatches(W, lo, len, dir, code4, rcode4, m, 0, i, n - i);
            if (mis <= k && i >= min_ov && mis <= (int)kmax[i]) {
                L.h.i1 = imin(L.h.i1, i);
                L.h.i2 = imax(L.h.i2, i);
                Cell c;
                c.cost = mis; c.score = i - 2 * mis; c.origin = n - i;
                r6_update(L.best, n, c, i, min_ov, kmax);
            }
        }
    } else if (last_cand) {
        r6_update(L.best, n, last, m, min_ov, kmax);       // FRONT: R6 looks at the single cell (m, n)
    }
}

// All windows of one pair.  wl == nullptr: no prefix filter, one window over the whole view.
ORC_HD void scan_lane(const uint32_t *__restrict__ W, uint64_t lo, uint32_t len, int dir,
                      const WinList *wl, const char *peq_base, int lane, uint64_t pv0, int d0,
                      int m, int k, const uint8_t *kmax, int min_ov, int type, LaneScan &L,
                      int indels, const uint32_t *code4, const uint32_t *rcode4, const uint8_t *lut)
{
    lane_scan_init(L, m, (int)len);
    const uint32_t nw = wl ? wl->n : 1u;
    for (uint32_t w = 0; w < nw; w++) {
        const uint32_t s = wl ? wl->s[w] : 0u, e = wl ? wl->e[w] : len;
        if (indels) scan_window(W, lo, len, dir, s, e, peq_base, lane, pv0, d0, m, k, kmax, min_ov, type, L, lut);
        else scan_window_noindel(W, lo, len, dir, s, e, code4, rcode4, m, k, kmax, min_ov, type, L);
    }
}

// Aligner.locate's return value (R7) from the running best
ORC_HD void best_to_result(const Best &best, int m, int n, PairResult &res)
{
    res.pad_ = 0;
    if (best.cost == m + n + 1) {
        res.has = 0; res.ref_start = res.ref_stop = res.query_start = res.query_stop = res.score = res.errors = 0;
        return;
    }
    res.has = 1;
    if (best.origin >= 0) { res.ref_start = 0; res.query_start = best.origin; }
    else { res.ref_start = -best.origin; res.query_start = 0; }
    res.ref_stop = best.ref_stop; res.query_stop = best.query_stop;
    res.score = best.score; res.errors = best.cost;
}

// R8 as one unsigned 64-bit maximum per (read, orientation): higher score, then fewer errors,
// then the adapter that comes first in the file; the low word says where the result lives.
ORC_HD uint64_t pack_key(int score, int errors, int adapter, uint32_t slot)
{
    return ((uint64_t)(uint32_t)(score + 512) << 44) | ((uint64_t)(uint32_t)(63 - errors) << 38) |
           ((uint64_t)(uint32_t)(63 - adapter) << 32) | (uint64_t)slot;
}

// ------------------------------------------------------------------------------------
// resolve_pair: exact (cost, score, origin) on a diagonal band, then R5/R6/R7.
// ------------------------------------------------------------------------------------
// The resolver re-runs the bit-parallel scan over the columns that matter for one pair and
// keeps the vertical deltas (Pv, Mv) and D[m][j] of the last RING columns.  From them any
// cost D[i][j'] of those columns is a popcount away, which is all that cutadapt's
// predecessor rule (R3) looks at -- so the path that the forward recurrence would have
// inherited score and origin along can be walked backwards from a candidate end cell
// (DESIGN.md section 4).  A scan restarted at column ws > 0 with column costs D[i][ws] = i
// over-estimates costs, but is exact for every cell whose optimal path starts at a column
// >= ws, and all cells a candidate's walk looks at are of that kind (ws <= j - m - k - 1).
constexpr int RING = 128;            // >= m + k + 2 for every supported adapter (k < m <= 64)
struct ColRing {
    uint64_t pv[RING], mv[RING];
};

// The ring is indexed by the column's distance from the scan start ws: the lanes of a warp
// scan in lockstep, so they touch the same slot at the same time and the lane-interleaved
// local memory sees one coalesced access per column instead of 32 scattered ones.
// D[i][j] of a stored column j, i in 0..m: row 0 costs 0 in every column, so the cost is the sum
// of the vertical deltas of rows 1..i (the padding bits below row 1 are clear in both words).
ORC_HD int ring_cost(const ColRing &R, int m, int i, int j, int ws)
{
    if (i <= 0) return 0;
    const int x = (j - ws) & (RING - 1);
    const int sh = m - i;                           // drops rows i+1..m off the top
    return popc64(R.pv[x] << sh) - popc64(R.mv[x] << sh);
}

// Walk cutadapt's path back from cell (i, j) whose cost is d.  ws is the first stored
// column (the restart column, or 0).  Yields the score and the origin of the cell.
//
// On equal characters cutadapt always takes the diagonal, so most of a path is runs of matches
// along a diagonal -- consecutive characters of both strings.  A run is measured 16 cells at a
// time: 16 packed read codes AND 16 packed adapter codes, first all-zero nibble = first mismatch.
// Only at a mismatch cell (at most k per path) are the neighbour costs looked up in the ring.
ORC_HD void trace_back(const uint32_t *W, uint64_t lo, uint32_t len, int dir, const uint64_t *peq_lane,
                       const uint32_t *code4, const uint32_t *rcode4,
                       int m, int type, int ws, const ColRing &R, int i, int j, int d,
                       int &score_out, int &origin_out)
{
    int score = 0, origin = 0;
    for (;;) {
        if (i == 0) { origin = j; break; }           // row 0: cost 0, score 0, origin = column (R3)
        if (j == 0) {                                // column 0 (R2): FRONT origin -i, BACK origin 0
            origin = (type == TYPE_FRONT) ? -i : 0;
            break;
        }
        if (j <= ws) { origin = j; break; }          // unreachable for a genuine candidate
        // matches along the diagonal from (i, j) up-left, at most min(16, i, j - ws) of them.
        // Forward lane: read codes j-16..j-1 of the view = storage lo+j-16.., adapter codes
        // i-16..i-1 (16 pad nibbles in front of code4), cell (i, j) in the top nibble.  Reverse
        // lane: view position p = storage lo+len-1-p, complemented, so positions j-1, j-2, ... are
        // storage lo+len-j, +1, ... compared raw with the reverse-complemented adapter; cell
        // (i, j) is in the bottom nibble, brought to the top by a bit reversal.  Both kinds of
        // lane run the same instructions.
        const int avail = imin(16, imin(i, j - ws));
        const uint64_t r = nib16(W, dir ? (int64_t)lo + (int64_t)len - j : (int64_t)lo + j - 16);
        const uint64_t a = nib16(dir ? rcode4 : code4, dir ? (int64_t)(m - i) : (int64_t)i);
        uint64_t x = r & a;
        x |= x >> 1; x |= x >> 2;
        uint64_t mis = ~x & 0x1111111111111111ull;
        if (dir) mis = brev64(mis);                              // nibble q -> nibble 15-q
        const int run = imin(clz64(mis) >> 2, avail);
        score += run; i -= run; j -= run;                        // characters equal: diagonal, unconditionally
        if (run == avail) continue;                              // cap or border reached: look again
        const int bit = 64 - m + i - 1;
        const uint64_t pvj = R.pv[(j - ws) & (RING - 1)], mvj = R.mv[(j - ws) & (RING - 1)];
        const uint64_t pvl = R.pv[(j - 1 - ws) & (RING - 1)], mvl = R.mv[(j - 1 - ws) & (RING - 1)];
        const int d_up = d - (int)((pvj >> bit) & 1u) + (int)((mvj >> bit) & 1u);
        const int d_left = ring_cost(R, m, i, j - 1, ws);
        const int d_diag = d_left - (int)((pvl >> bit) & 1u) + (int)((mvl >> bit) & 1u);
        const int c_diag = d_diag + 1, c_del = d_left + 1, c_ins = d_up + 1;
        if (c_diag <= c_del && c_diag <= c_ins) { score -= 1; --i; --j; d = d_diag; }
        else if (c_ins <= c_del) { score -= 2; --i; d = d_up; }
        else { score -= 2; --j; d = d_left; }
    }
    score_out = score;
    origin_out = origin;
}

// trace_back() for a group of lanes that walk at the same time (all lanes of `mask` call this,
// `on` says which of them have a walk to do): every trip of the loop starts with a vote, so the
// lanes take their steps side by side instead of drifting apart at the first branch.
ORC_HD void trace_back_group(uint32_t mask, bool on, const uint32_t *W, uint64_t lo, uint32_t len, int dir,
                             const uint32_t *code4, const uint32_t *rcode4,
                             int m, int type, int ws, const ColRing &R, int i, int j, int d,
                             int &score_out, int &origin_out)
{
    int score = 0, origin = 0;
    while (ORC_ANY(mask, on)) {
        if (on) {
            if (i == 0) { origin = j; on = false; }                  // row 0 (R3)
            else if (j == 0) { origin = (type == TYPE_FRONT) ? -i : 0; on = false; }   // column 0 (R2)
            else if (j <= ws) { origin = j; on = false; }            // unreachable for a genuine candidate
            else {
                const int avail = imin(16, imin(i, j - ws));
                const uint64_t r = nib16(W, dir ? (int64_t)lo + (int64_t)len - j : (int64_t)lo + j - 16);
                const uint64_t a = nib16(dir ? rcode4 : code4, dir ? (int64_t)(m - i) : (int64_t)i);
                uint64_t x = r & a;
                x |= x >> 1; x |= x >> 2;
                uint64_t mis = ~x & 0x1111111111111111ull;
                if (dir) mis = brev64(mis);
                const int run = imin(clz64(mis) >> 2, avail);
                score += run; i -= run; j -= run;
                if (run < avail) {                                   // a mismatch cell inside the matrix
                    const int bit = 64 - m + i - 1;
                    const uint64_t pvj = R.pv[(j - ws) & (RING - 1)], mvj = R.mv[(j - ws) & (RING - 1)];
                    const uint64_t pvl = R.pv[(j - 1 - ws) & (RING - 1)], mvl = R.mv[(j - 1 - ws) & (RING - 1)];
                    const int d_up = d - (int)((pvj >> bit) & 1u) + (int)((mvj >> bit) & 1u);
                    const int d_left = ring_cost(R, m, i, j - 1, ws);
                    const int d_diag = d_left - (int)((pvl >> bit) & 1u) + (int)((mvl >> bit) & 1u);
                    const int c_diag = d_diag + 1, c_del = d_left + 1, c_ins = d_up + 1;
                    if (c_diag <= c_del && c_diag <= c_ins) { score -= 1; --i; --j; d = d_diag; }
                    else if (c_ins <= c_del) { score -= 2; --i; d = d_up; }
                    else { score -= 2; --j; d = d_left; }
                }
            }
        }
    }
    score_out = score;
    origin_out = origin;
}

// The resolver's state for one task.  The work is cut in two so that the kernel can bring the
// lanes of a warp back together in between: resolve_begin() runs the column scan (lanes differ
// in the number of columns), resolve_end() does the walks, which then start together.
struct ResolveCtx {
    const uint64_t *peq_lane;
    const uint32_t *code4, *rcode4;
    const uint8_t *kmax;
    int32_t dir, lane, m, k, min_ov, n, type;
    int32_t has6, join, second;         // second: a separate scan of the last columns follows
    int32_t r6lo, r6hi, ws6;
    // the scan in progress
    int32_t ws, we, jf, jl, r6, narrow, ubw, broke;
    int32_t top_j, traced_j, traced_score, traced_origin;
    Best best;
};

// Columns ws+1..we of one scan: the ring is filled, R5 runs on the candidate columns jf..jl.
// Narrow hull: when all candidate columns lie within m/2 - 2k of each other, so do the origins
// of any two candidates (origin is within k of j - m; a negative origin only ever helps the
// second clause), and R5's update rule reduces to "the first acceptable candidate, then any with
// a strictly higher score": the winner is the acceptable candidate of highest score, leftmost
// among equals, in whatever order they are looked at.  Their walks are left to resolve_finish();
// here only the candidate with the highest score bound is noted.  (The whole scan must still be
// in the ring by then.)  Otherwise the walks happen on the spot, in column order.
ORC_HD void resolve_columns(const uint32_t *W, uint64_t lo, uint32_t len, ResolveCtx &C, ColRing &R)
{
    const int m = C.m, n = C.n, k = C.k, ws = C.ws, we = C.we, jf = C.jf, jl = C.jl, dir = C.dir;
    const uint64_t pad = (m == 64) ? 0ull : ((1ull << (64 - m)) - 1ull);
    uint64_t Pv, Mv = 0;
    if (ws == 0 && C.type == TYPE_FRONT) Pv = 0;                // R2, true column 0 of a 5' adapter
    else Pv = ~pad;                                             // cost i (true for BACK at 0; restart otherwise)
    R.pv[0] = Pv; R.mv[0] = 0;
    C.traced_j = -1; C.traced_score = 0; C.traced_origin = 0;
    C.narrow = (jf <= jl && (jl - jf) + 2 * k <= m / 2 && we - ws < RING) ? 1 : 0;
    C.top_j = -1;
    C.broke = 0;
    int top_ub = -(1 << 20);
    // Upper bound of a candidate's score: every error costs at least 2 (score <= length - 2*cost),
    // except the errors a 3' adapter takes in column 0 (cost i, score 0: R2), reachable only if
    // the scan starts at the true column 0; then only score <= length - cost holds.
    C.ubw = (ws == 0 && C.type == TYPE_BACK) ? 1 : 2;
    const int ubw = C.ubw;
    const bool narrow = C.narrow != 0;
    // eight columns per packed word, table address by PRMT, as in scan_window
    const char *peq_base = reinterpret_cast<const char *>(C.peq_lane - (C.lane & 31));
    const uint32_t lane8 = (uint32_t)(C.lane & 31) * 8u;
    uint32_t sel0, sel1, sel2, sel3;
    if (!dir) { sel0 = 0x5504u; sel1 = 0x5514u; sel2 = 0x5524u; sel3 = 0x5534u; }
    else      { sel0 = 0x5534u; sel1 = 0x5524u; sel2 = 0x5514u; sel3 = 0x5504u; }
    ChunkReader rd;
    rd.init(W, lo, len, dir, (uint32_t)ws);
    const int ncols = we - ws;
    const int nchunks = (ncols + 7) >> 3;
    for (int q = 0; q < nchunks; q++) {
        uint32_t A, B;
        rd.next(A, B);
        const int ncol = imin(8, ncols - 8 * q);
        const int jb = ws + 8 * q;                              // this chunk: columns jb+1 .. jb+ncol
        if (ncol == 8 && (jb + 8 < jf || jb + 1 > jl)) {
            // no candidate column in here: just the recurrence and the ring
#pragma unroll
            for (int t = 0; t < 8; t++) {
                const uint32_t src = (t & 1) ? B : A;
                const uint32_t sel = (t >> 1) == 0 ? sel0 : (t >> 1) == 1 ? sel1 : (t >> 1) == 2 ? sel2 : sel3;
                const uint64_t Eq = *reinterpret_cast<const uint64_t *>(peq_base + byte_perm(src, lane8, sel));
                const uint64_t Xv = Eq | Mv;
                const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
                uint64_t Ph = Mv | ~(Xh | Pv);
                uint64_t Mh = Pv & Xh;
                Ph <<= 1; Mh <<= 1;
                Pv = Mh | ~(Xv | Ph);
                Mv = Ph & Xv;
                const int x = (8 * q + t + 1) & (RING - 1);
                R.pv[x] = Pv; R.mv[x] = Mv;
            }
            continue;
        }
        int D = popc64(Pv) - popc64(Mv);                        // D[m][jb]
#pragma unroll 1
        for (int t = 0; t < ncol; t++) {
            const int j = jb + t + 1;
            const uint32_t src = (t & 1) ? B : A;
            const uint32_t sel = (t >> 1) == 0 ? sel0 : (t >> 1) == 1 ? sel1 : (t >> 1) == 2 ? sel2 : sel3;
            const uint64_t Eq = *reinterpret_cast<const uint64_t *>(peq_base + byte_perm(src, lane8, sel));
            const uint64_t Xv = Eq | Mv;
            const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
            uint64_t Ph = Mv | ~(Xh | Pv);
            uint64_t Mh = Pv & Xh;
            D += (int)(Ph >> 63) - (int)(Mh >> 63);
            Ph <<= 1; Mh <<= 1;
            Pv = Mh | ~(Xv | Ph);
            Mv = Ph & Xv;
            { const int x = (j - ws) & (RING - 1); R.pv[x] = Pv; R.mv[x] = Mv; }
            if (j >= jf && j <= jl && D <= k) {
                const int lmax = imin(m, j + D);
                if (lmax >= C.min_ov && D <= (int)C.kmax[lmax]) {
                    const int ub = lmax - ubw * D;
                    if (narrow) {
                        if (ub > top_ub) { top_ub = ub; C.top_j = j; }     // highest bound, leftmost
                    } else if (C.best.cost == m + n + 1 || ub > C.best.score) {
                        // a candidate that cannot beat the best so far cannot change it (every R5
                        // update after the first needs a strictly higher score)
                        Cell c;
                        c.cost = D;
                        trace_back(W, lo, len, dir, C.peq_lane, C.code4, C.rcode4, m, C.type, ws, R, m, j, D, c.score, c.origin);
                        C.traced_j = j; C.traced_score = c.score; C.traced_origin = c.origin;
                        if (r5_update(C.best, m, n, c, j, C.min_ov, C.kmax)) { C.broke = 1; return; }
                    }
                }
            }
        }
    }
}

// After resolve_columns(): the walks of a narrow hull, then R6 on the rows r6lo..r6hi of column
// n if the scan reached it.
ORC_HD void resolve_finish(const uint32_t *W, uint64_t lo, uint32_t len, ResolveCtx &C, ColRing &R,
                           uint32_t mask, bool grouped)
{
    const int m = C.m, n = C.n, k = C.k, ws = C.ws, dir = C.dir, ubw = C.ubw;
    Best &best = C.best;
    if (grouped) {
        // The walks of a narrow hull, the noted candidate first.  All lanes of `mask` are here;
        // each picks its next candidate worth a walk (none: it only keeps voting), then they walk
        // together, until no lane has one left.
        bool more = !C.broke && C.narrow && C.top_j >= 0;
        int q = C.jf - 1;
        while (ORC_ANY(mask, more)) {
            int j = 0, Dj = 0;
            bool have = false;
            if (more) {
                for (; q <= C.jl && !have; q++) {
                    j = q < C.jf ? C.top_j : q;
                    Dj = ring_cost(R, m, m, j, ws);
                    if (q < C.jf) { have = true; continue; }
                    if (j == C.top_j || Dj > k) continue;
                    const int lmax = imin(m, j + Dj);
                    if (!(lmax >= C.min_ov && Dj <= (int)C.kmax[lmax])) continue;
                    const int ub = lmax - ubw * Dj;
                    have = best.cost == m + n + 1 || ub > best.score || (ub == best.score && j < best.query_stop);
                }
                if (!have) more = false;
            }
            Cell c;
            c.cost = Dj; c.score = 0; c.origin = 0;
            trace_back_group(mask, have, W, lo, len, dir, C.code4, C.rcode4, m, C.type, ws, R, m, j, Dj, c.score, c.origin);
            if (have) {
                if (j == n) { C.traced_j = j; C.traced_score = c.score; C.traced_origin = c.origin; }
                const int length = m + imin(c.origin, 0);
                if (length >= C.min_ov && c.cost <= (int)C.kmax[KMAX_R5 + length] &&
                    (best.cost == m + n + 1 || c.score > best.score || (c.score == best.score && j < best.query_stop))) {
                    best.score = c.score; best.cost = c.cost; best.origin = c.origin; best.ref_stop = m; best.query_stop = j;
                }
            }
        }
        if (!C.broke && C.narrow && best.cost == 0 && best.origin >= 0) C.broke = 1;     // R5's early exit
    }
    if (C.broke) return;
    if (C.r6 && C.we == n && n > ws) {
        // R6: rows r6hi..r6lo of column n, top row first like cutadapt
        for (int i = imin(C.r6hi, m); i >= imax(C.r6lo, 1); i--) {
            const int Di = ring_cost(R, m, i, n, ws);
            if (Di > k) continue;
            const int lmax = (C.type == TYPE_FRONT) ? imin(i, n + Di) : i;
            if (!(lmax >= C.min_ov && Di <= (int)C.kmax[lmax])) continue;
            const int ub = lmax - ubw * Di;
            if (ub < best.score || (ub == best.score && Di >= best.cost)) continue;
            Cell c;
            c.cost = Di;
            if (i == m && C.traced_j == n) { c.score = C.traced_score; c.origin = C.traced_origin; }
            else trace_back(W, lo, len, dir, C.peq_lane, C.code4, C.rcode4, m, C.type, ws, R, i, n, Di, c.score, c.origin);
            r6_update(best, n, c, i, C.min_ov, C.kmax);
        }
    }
}

// First half of a task: set up, run the column scan.
ORC_HD void resolve_begin(const uint32_t *W, const View &v, const RoundTable &T, const Task &t,
                          ResolveCtx &C, ColRing &R)
{
    const int a = (int)t.lane % T.n_adapters;
    C.dir = (int)t.lane / T.n_adapters;
    C.m = T.m[a]; C.k = T.k[a]; C.min_ov = T.min_ov[a];
    C.kmax = T.kmax[a][0];
    C.peq_lane = &T.peq[t.lane >> 5][0][t.lane & 31];
    C.lane = (int)t.lane;
    C.code4 = T.code4[a]; C.rcode4 = T.rcode4[a];
    C.n = (int)v.len;
    C.type = T.type;
    const int m = C.m, n = C.n, k = C.k;
    C.best.ref_stop = m; C.best.query_stop = n; C.best.cost = m + n + 1; C.best.origin = 0; C.best.score = 0;
    const bool has5 = t.jf <= t.jl;
    const bool has6 = (T.type == TYPE_BACK) ? (t.i1 <= t.i2) : (has5 && t.jl == n);   // FRONT: only cell (m, n)
    C.has6 = has6 ? 1 : 0;
    C.r6lo = (T.type == TYPE_BACK) ? t.i1 : m;
    C.r6hi = (T.type == TYPE_BACK) ? t.i2 : m;
    C.ws6 = imax(0, n - C.r6hi - k - 1);
    C.second = 0; C.join = 0; C.broke = 0; C.narrow = 0; C.top_j = -1; C.r6 = 0;
    C.ws = 0; C.we = 0; C.jf = 1; C.jl = 0; C.ubw = 2;
    C.traced_j = -1; C.traced_score = 0; C.traced_origin = 0;
    if (has5) {
        C.ws = imax(0, t.jf - m - k - 1);
        // run on to column n in the same scan when the last-column cells are close enough
        const bool join = has6 && (C.ws6 <= t.jl + 1 || C.ws == 0 && C.ws6 == 0);
        C.join = join ? 1 : 0;
        C.second = (has6 && !join) ? 1 : 0;
        C.we = join ? n : t.jl; C.jf = t.jf; C.jl = t.jl; C.r6 = join ? 1 : 0;
        resolve_columns(W, v.lo, v.len, C, R);
    } else if (has6) {
        C.ws = C.ws6; C.we = n; C.jf = 1; C.jl = 0; C.r6 = 1;
        resolve_columns(W, v.lo, v.len, C, R);
    }
}

// Second half: the walks, the last-column cells, the result.
ORC_HD void resolve_end(const uint32_t *W, const View &v, ResolveCtx &C, PairResult &res, ColRing &R,
                        uint32_t mask = 0xffffffffu)
{
    resolve_finish(W, v.lo, v.len, C, R, mask, true);
    if (C.second && !C.broke) {
        C.ws = C.ws6; C.we = C.n; C.jf = 1; C.jl = 0; C.r6 = 1;
        resolve_columns(W, v.lo, v.len, C, R);
        resolve_finish(W, v.lo, v.len, C, R, mask, false);      // no R5 candidates in this scan
    }
    best_to_result(C.best, C.m, C.n, res);
}

ORC_HD void resolve_pair(const uint32_t *W, const View &v, const RoundTable &T, const Task &t,
                         PairResult &res, ColRing &R)
{
    ResolveCtx C;
    resolve_begin(W, v, T, t, C, R);
    resolve_end(W, v, C, res, R);
}

// ------------------------------------------------------------------------------------
// The band resolver: the same re-scan and the same walks, on 8 bytes per column.
//
// What a walk needs at a cell (i, j) whose characters differ (cost d >= 1, known) is cutadapt's choice
// among its three predecessors (R3).  The cell's cost is the minimum of the three candidates, all of
// which are >= d, so
//   * the diagonal wins (first in cutadapt's order) iff D[i-1][j-1] == d - 1, i.e. iff the cell's
//     diagonal delta is 1 -- the complement of Hyyro's "diagonal zero" vector D0 = Xh | Mv;
//   * otherwise the insertion (up) wins iff D[i-1][j] == d - 1, i.e. iff the vertical delta is +1 -- Pv;
//   * otherwise the deletion (left).
// Every move from such a cell lowers the cost by exactly one and a match keeps it, so no cost is ever
// looked up: two bits per cell decide.  (A scan restarted at ws > 0 over-estimates costs elsewhere, but a
// predecessor with the true cost d - 1 lies on an optimal path to the candidate and is exact, and one
// whose true cost is >= d stays >= d: both bits are those of the true matrix at every cell a walk visits.)
//
// A path with at most k errors stays within k diagonals of where it ends, so per column only the rows
// around the end cells' diagonals matter: with end cells on diagonals dlo..dhi (column minus row), column
// j needs the 32 rows from j - (dhi + k + 1) upwards of Pv and D0, and no costs.
//
// A walk looks at a column only where its path meets a cell whose characters differ -- at most k times --
// so the columns are not kept at all: the ring holds the scan's state (VP, VN on the band, 8 bytes) after
// every BAND_CKPT-th column, and band_step() re-runs the one to BAND_CKPT columns from the checkpoint in
// front of the cell it asks about (a few dozen instructions, a handful of times per task, against the
// ~2000 of the scan).  With a checkpoint every second column that is 320 bytes of ring and 64 bytes of
// read codes per thread in shared memory, 16 warps per SM -- the first band resolver's 640-byte ring (one
// entry per column) held the kernel at 8; the wide resolver's 2 KB ring lives in local memory.  Measured
// per 1 Mi COI reads (rounds 1 + 2): every column 0.64 + 0.46 ms, every 2nd 0.56 + 0.38, every 4th
// 0.61 + 0.42 (the walks of a warp run in lockstep, so a re-run costs a warp instruction stream for the
// few lanes that need it), every 8th 0.73 + 0.49.
// Tasks whose end cells span more than 30 - 2k diagonals, or whose scan is longer than BAND_COLS columns,
// keep the wide resolver above.
// ------------------------------------------------------------------------------------
constexpr int BAND_COLS = 80;           // columns of one scan (column 0 = the scan's start state)
#ifndef ORC_BAND_CKPT_LOG
#define ORC_BAND_CKPT_LOG 1
#endif
constexpr int BAND_CKPT_LOG = ORC_BAND_CKPT_LOG;
constexpr int BAND_CKPT = 1 << BAND_CKPT_LOG;           // the state after columns 0, 2, 4, ... is kept
constexpr int BAND_ENTRIES = BAND_COLS / BAND_CKPT;     // a scan has at most BAND_COLS - 1 columns
static_assert(BAND_CKPT <= 8 && BAND_COLS % BAND_CKPT == 0, "band_columns stores checkpoints inside its chunks of 8 columns");
constexpr int BAND_CODE_WORDS = 16;     // packed read codes a scan and its walks touch: BAND_COLS + 16 before + 24 behind
constexpr uint32_t TASK_WIDE = 1u;      // Task.pad_ bit 0: not eligible for the band resolver

struct alignas(8) BandEntry { uint32_t vp, vn; };      // the band's vertical deltas after a checkpoint column
// What the band resolver reads of one adapter, compact (the kernel keeps one per adapter of the round in
// shared memory; RoundTable spreads the same over arrays sized for MAX_AD adapters)
struct BandAdapter {
    int32_t m, k, min_ov, pad_;
    uint32_t code4[12], rcode4[12];
    uint8_t kmax[3][MAX_M + 8];         // RoundTable.kmax[a]: pruning limits, R5 limits, R6 limits
};
struct BandRing {                       // checkpoint q (the state after column q * BAND_CKPT) of this thread: p[q * stride]
    BandEntry *p;
    int32_t stride;
    // the packed codes of the read around the scan, copied once per scan (16 loads in flight together) so that
    // neither the column loop nor the walks wait for global memory: word w0 + i of the code array at cw[i * stride]
    uint32_t *cw;
    int64_t w0;
};

// 16 consecutive codes starting at code index idx, from the ring's copy
ORC_HD uint64_t band_nib16(const BandRing &R, int64_t idx)
{
    const int w = (int)((idx >> 3) - R.w0);
    const uint32_t sh = (uint32_t)(idx & 7) * 4u;
    const uint32_t a = R.cw[w * R.stride], b = R.cw[(w + 1) * R.stride], c = R.cw[(w + 2) * R.stride];
    return (uint64_t)funnel_r(a, b, sh) | ((uint64_t)funnel_r(b, c, sh) << 32);
}

// ChunkReader over the ring's copy of the codes
struct BandReader {
    const uint32_t *cw;
    int32_t stride, w, dir_;
    uint32_t lo_, hi_, sh, shA, shB;
    ORC_HD void init(const BandRing &R, uint64_t lo, uint32_t len, int dir, uint32_t p0)
    {
        const int64_t s = dir ? (int64_t)lo + (int64_t)len - 8 - (int64_t)p0 : (int64_t)lo + (int64_t)p0;
        cw = R.cw; stride = R.stride; dir_ = dir;
        w = (int)((s >> 3) - R.w0);
        sh = (uint32_t)(s & 7) * 4u;
        shA = dir ? 4u : 0u;
        shB = dir ? 0u : 4u;
        if (!dir) { hi_ = cw[w * stride]; w += 1; lo_ = 0; }
        else      { lo_ = cw[(w + 1) * stride]; hi_ = 0; }
    }
    ORC_HD void next(uint32_t &A, uint32_t &B)
    {
        const uint32_t nw = cw[w * stride];
        if (!dir_) { lo_ = hi_; hi_ = nw; w += 1; }
        else       { hi_ = lo_; lo_ = nw; w -= 1; }
        const uint32_t x = funnel_r(lo_, hi_, sh);
        A = (x >> shA) & 0x0F0F0F0Fu;
        B = (x >> shB) & 0x0F0F0F0Fu;
    }
};

// Copy the code words one scan (columns ws+1..we of the view, lane direction dir) and its walks read.
ORC_HD void band_load_codes(const uint32_t *W, uint64_t lo, uint32_t len, int dir, int ws, int we, BandRing &R)
{
    // lowest code index touched: direction 0 walks look 16 codes back from column ws + 1; direction 1 reads
    // upwards from the code of column we
    const int64_t first = dir ? (int64_t)lo + (int64_t)len - we : (int64_t)lo + ws - 16;
    R.w0 = first >> 3;                  // arithmetic shift: floors also below zero (guard words in front)
#pragma unroll
    for (int i = 0; i < BAND_CODE_WORDS; i++) R.cw[i * R.stride] = W[R.w0 + i];
}

// Geometry of a task's scans (what resolve_begin derives from the hull).
struct ResolveGeo {
    int32_t has5, has6, r6lo, r6hi, ws6;
    int32_t ws, we, jf, jl, r6, join, second;       // the first scan
};

ORC_HD void resolve_geometry(int type, int m, int k, int n, const Task &t, ResolveGeo &G)
{
    const bool has5 = t.jf <= t.jl;
    const bool has6 = (type == TYPE_BACK) ? (t.i1 <= t.i2) : (has5 && t.jl == n);   // FRONT: only cell (m, n)
    G.has5 = has5 ? 1 : 0; G.has6 = has6 ? 1 : 0;
    G.r6lo = (type == TYPE_BACK) ? t.i1 : m;
    G.r6hi = (type == TYPE_BACK) ? t.i2 : m;
    G.ws6 = imax(0, n - G.r6hi - k - 1);
    G.ws = 0; G.we = 0; G.jf = 1; G.jl = 0; G.r6 = 0; G.join = 0; G.second = 0;
    if (has5) {
        G.ws = imax(0, t.jf - m - k - 1);
        const bool join = has6 && (G.ws6 <= t.jl + 1 || (G.ws == 0 && G.ws6 == 0));
        G.join = join ? 1 : 0;
        G.second = (has6 && !join) ? 1 : 0;
        G.we = join ? n : t.jl; G.jf = t.jf; G.jl = t.jl; G.r6 = join ? 1 : 0;
    } else if (has6) {
        G.ws = G.ws6; G.we = n; G.r6 = 1;
    }
}

// May one scan (columns ws+1..we, R5 candidates jf..jl, R6 rows r6lo..r6hi if r6) use the band ring?
// boff = dhi + k + 1: cell (i, j) is bit boff - (j - i) of column j's window.
ORC_HD bool band_scan_ok(int m, int k, int n, int ws, int we, int jf, int jl, bool r6, int r6lo, int r6hi, int &boff)
{
    int dlo = 0x3fffffff, dhi = -0x3fffffff;
    if (jf <= jl) { dlo = jf - m; dhi = jl - m; }
    if (r6 && we == n) {
        const int lo = n - imin(r6hi, m), hi = n - imax(r6lo, 1);
        if (lo <= hi) { dlo = imin(dlo, lo); dhi = imax(dhi, hi); }
    }
    boff = 0;
    if (k > 14) return false;                           // hull costs are kept in four bits, 15 = none
    if (dlo > dhi) return we - ws < BAND_COLS;          // nothing to walk from
    if ((dhi - dlo) + 2 * k + 3 > 32) return false;     // one spare row on either side
    if (we - ws >= BAND_COLS) return false;
    boff = dhi + k + 1;
    return true;
}

ORC_HD bool task_band_ok(int type, int m, int k, int n, const Task &t)
{
    ResolveGeo G;
    resolve_geometry(type, m, k, n, t, G);
    int boff;
    if (!G.has5 && !G.has6) return true;
    if (!band_scan_ok(m, k, n, G.ws, G.we, G.jf, G.jl, G.r6 != 0, G.r6lo, G.r6hi, boff)) return false;
    if (G.second && !band_scan_ok(m, k, n, G.ws6, n, 1, 0, true, G.r6lo, G.r6hi, boff)) return false;
    return true;
}

ORC_HD void band_adapter_fill(const RoundTable &T, int a, BandAdapter &B)
{
    B.m = T.m[a]; B.k = T.k[a]; B.min_ov = T.min_ov[a]; B.pad_ = 0;
    for (int i = 0; i < 12; i++) { B.code4[i] = T.code4[a][i]; B.rcode4[i] = T.rcode4[a][i]; }
    for (int t = 0; t < 3; t++)
        for (int i = 0; i < MAX_M + 8; i++) B.kmax[t][i] = T.kmax[a][t][i];
}

// The band resolver's own match table: per (lane bank, read code, lane % 32) four words {ones, low word,
// high word, zeros} of the lane's 64-bit match vector, so that the 32 band rows of any column -- rows below
// the adapter read as matches (they stand for row 0), rows above it as mismatches -- are two adjacent words
// and one funnel shift away: entry at code * 512 + (lane % 32) * 16 inside a bank of BAND_BANK_BYTES.
constexpr int BAND_BANK_BYTES = 16 * 32 * 16;
ORC_HD void band_table_entry(const RoundTable &T, int bank, int code, int l, uint32_t out[4])
{
    const uint64_t v = T.peq[bank][code][l];
    out[0] = 0xFFFFFFFFu; out[1] = (uint32_t)v; out[2] = (uint32_t)(v >> 32); out[3] = 0u;
}

struct BandCtx {
    const char *peq_base;               // the lane's bank of the band match table (band_table_entry)
    const uint32_t *code4, *rcode4;
    const uint8_t *kmax;
    int32_t dir, lane, m, k, min_ov, n, type;
    int32_t c5, c6;                     // D[m][jf], D[i1][n] (from the scan, exact)
    ResolveGeo G;
    // the scan in progress
    int32_t ws, we, jf, jl, r6, narrow, ubw, broke, boff;
    int32_t top_j, traced_j, traced_score, traced_origin;
    uint64_t hc0, hc1;                  // D[m][j] of the hull columns jf .. jf+31, four bits each (15: above k)
    uint32_t vpn, vnn;                  // vertical deltas of column we in its band rows
    uint32_t hull_p, hull_n;            // horizontal deltas of row m in the hull columns: bit j - jf
    int32_t sp0;                        // band position 0 of column ws + c in the match table: sp0 + c
    Best best;
};

ORC_HD int band_hull_cost(const BandCtx &C, int j)
{
    const int q = j - C.jf;
    return (int)(((q < 16 ? C.hc0 : C.hc1) >> (4 * (q & 15))) & 15ull);
}

// x << s, 0 for s >= 32 (and for "negative" s)
ORC_HD uint32_t shl32(uint32_t x, uint32_t s) { return s < 32u ? x << s : 0u; }

// The 32 band rows of the match vector for read code `code` in column ws + c.
ORC_HD uint32_t band_eq(const char *peq_base, uint32_t code_lane, int sp0, int c)
{
    const int sp = imin(imax(sp0 + c, 0), 95);
    const uint32_t *e = reinterpret_cast<const uint32_t *>(peq_base + 2u * code_lane) + (sp >> 5);
    return funnel_r(e[0], e[1], (uint32_t)sp & 31u);
}

// One column of the band recurrence (see band_columns): VP, VN of the previous column in, of this one out.
ORC_HD void band_advance(uint32_t Eq, uint32_t &VP, uint32_t &VN, uint32_t &D0, uint32_t &HP, uint32_t &HN)
{
    VP = (VP >> 1) | 0x80000000u;
    VN >>= 1;
    D0 = (((Eq & VP) + VP) ^ VP) | Eq | VN;
    HP = VN | ~(D0 | VP);
    HN = D0 & VP;
    const uint32_t X = HP << 1;
    VP = (HN << 1) | ~(D0 | X);
    VN = D0 & X;
}

// One step of the walk at a cell whose characters differ: cutadapt's predecessor from two bits of column j,
// which is re-run from the checkpoint in front of it.
ORC_HD void band_step(const BandCtx &C, const BandRing &R, uint64_t lo, uint32_t len, int &i, int &j, int &d, int &score)
{
    const int c = j - C.ws;                                     // >= 1
    const int base = (c - 1) & ~(BAND_CKPT - 1);
    const BandEntry e = R.p[(base >> BAND_CKPT_LOG) * R.stride];
    uint32_t VP = e.vp, VN = e.vn, D0 = 0, HP, HN;
    const uint32_t lane8 = (uint32_t)(C.lane & 31) * 8u;        // band_eq doubles it: entries of 16 bytes
    for (int cc = base + 1; cc <= c; cc++) {
        const int64_t idx = C.dir ? (int64_t)lo + (int64_t)len - (C.ws + cc) : (int64_t)lo + (C.ws + cc) - 1;
        const uint32_t code = (R.cw[(int)((idx >> 3) - R.w0) * R.stride] >> ((uint32_t)(idx & 7) * 4u)) & 15u;
        band_advance(band_eq(C.peq_base, (code << 8) | lane8, C.sp0, cc), VP, VN, D0, HP, HN);
    }
    const uint32_t b = (uint32_t)(C.boff - (j - i));
    if (!((D0 >> b) & 1u)) { score -= 1; --i; --j; }            // diagonal delta 1: mismatch
    else if ((VP >> b) & 1u) { score -= 2; --i; }               // vertical delta +1: insertion
    else { score -= 2; --j; }                                   // deletion
    --d;
}

ORC_HD void band_trace_back(const BandCtx &C, const uint32_t *W, uint64_t lo, uint32_t len, int dir,
                            const uint32_t *code4, const uint32_t *rcode4, int m, int type, int ws,
                            const BandRing &R, int boff, int i, int j, int d, int &score_out, int &origin_out)
{
    int score = 0, origin = 0;
    for (;;) {
        if (i == 0) { origin = j; break; }
        if (j == 0) { origin = (type == TYPE_FRONT) ? -i : 0; break; }
        if (j <= ws) { origin = j; break; }          // unreachable for a genuine candidate
        const int avail = imin(16, imin(i, j - ws));
        const uint64_t r = band_nib16(R, dir ? (int64_t)lo + (int64_t)len - j : (int64_t)lo + j - 16);
        const uint64_t a = nib16(dir ? rcode4 : code4, dir ? (int64_t)(m - i) : (int64_t)i);
        uint64_t x = r & a;
        x |= x >> 1; x |= x >> 2;
        uint64_t mis = ~x & 0x1111111111111111ull;
        if (dir) mis = brev64(mis);
        const int run = imin(clz64(mis) >> 2, avail);
        score += run; i -= run; j -= run;
        if (run == avail) continue;
        band_step(C, R, lo, len, i, j, d, score);
    }
    score_out = score;
    origin_out = origin;
}

ORC_HD void band_trace_back_group(const BandCtx &C, uint32_t mask, bool on, const uint32_t *W, uint64_t lo, uint32_t len, int dir,
                                  const uint32_t *code4, const uint32_t *rcode4, int m, int type, int ws,
                                  const BandRing &R, int boff, int i, int j, int d, int &score_out, int &origin_out)
{
    int score = 0, origin = 0;
    while (ORC_ANY(mask, on)) {
        if (on) {
            if (i == 0) { origin = j; on = false; }
            else if (j == 0) { origin = (type == TYPE_FRONT) ? -i : 0; on = false; }
            else if (j <= ws) { origin = j; on = false; }
            else {
                const int avail = imin(16, imin(i, j - ws));
                const uint64_t r = band_nib16(R, dir ? (int64_t)lo + (int64_t)len - j : (int64_t)lo + j - 16);
                const uint64_t a = nib16(dir ? rcode4 : code4, dir ? (int64_t)(m - i) : (int64_t)i);
                uint64_t x = r & a;
                x |= x >> 1; x |= x >> 2;
                uint64_t mis = ~x & 0x1111111111111111ull;
                if (dir) mis = brev64(mis);
                const int run = imin(clz64(mis) >> 2, avail);
                score += run; i -= run; j -= run;
                if (run < avail) band_step(C, R, lo, len, i, j, d, score);
            }
        }
    }
    score_out = score;
    origin_out = origin;
}

// Columns ws+1..we of one scan; the state after every BAND_CKPT-th column goes to the ring.
//
// The recurrence itself runs on the band only (Hyyro's diagonal band): the 32-bit vectors hold rows
// j - boff .. j - boff + 31 of column j, so from one column to the next every row moves down one bit
// (VP, VN >>= 1) and a new row enters at the top.  Its vertical delta in the previous column is taken as
// +1, the largest it can be, and the first band row ignores its upper neighbour (no carry into the add):
// both only ever over-estimate costs outside the diagonals a path with <= k errors can use, so every cell
// on such a path -- and both of its decision bits -- is exact, by the argument used for restarted scans.
// Rows below row 1 stand for row 0 (cost 0 in every column): their match bits read as 1 and their deltas
// stay 0, like the padding bits of the 64-bit table.
//
// The column loop is the same for every lane and every column (the lanes of a warp differ only in its
// length).  Costs are not a popcount away here (row 0 is not in the band), and none is needed by the walks;
// for the hull columns' D[m][j] the horizontal deltas of row m in those columns are collected in two words
// (hull_p, hull_n: bit j - jf), and band_hull() counts on from the scan's D[m][jf] (Task anchors) afterwards.
ORC_HD void band_columns(const uint32_t *W, uint64_t lo, uint32_t len, BandCtx &C, BandRing &R)
{
    const int m = C.m, ws = C.ws, we = C.we, dir = C.dir, boff = C.boff;
    uint32_t VP, VN = 0;
    if (ws == 0 && C.type == TYPE_FRONT) VP = 0;                // R2: column 0 of a 5' adapter costs 0 in every row
    else {
        const int z = 1 - (ws - boff);                          // first band position of column ws that is a real row
        VP = z <= 0 ? 0xFFFFFFFFu : (z >= 32 ? 0u : (0xFFFFFFFFu << z));
    }
    const char *peq_base = C.peq_base;
    const uint32_t lane8 = (uint32_t)(C.lane & 31) * 8u;
    uint32_t sel0, sel1, sel2, sel3;
    if (!dir) { sel0 = 0x5504u; sel1 = 0x5514u; sel2 = 0x5524u; sel3 = 0x5534u; }
    else      { sel0 = 0x5534u; sel1 = 0x5524u; sel2 = 0x5514u; sel3 = 0x5504u; }
    band_load_codes(W, lo, len, dir, ws, we, R);
    BandReader rd;
    rd.init(R, lo, len, dir, (uint32_t)ws);
    const int ncols = we - ws;
    const int nchunks = (ncols + 7) >> 3;
    // band position 0 of column ws + c is vector bit 63 - m - boff + ws + c; the table has 32 rows of ones in
    // front of the vector (hence + 32), clamped to [0, 95]: below, everything matches; above, nothing does
    const int sp0 = 63 - m - boff + ws + 32;
    C.sp0 = sp0;
    // row m is band position boff - (j - m) of column j
    const int bm0 = boff + m - ws;
    const int h0 = C.jf - ws;                                   // column c is hull column c - h0 (if C.jf <= C.jl)
    uint32_t hull_p = 0, hull_n = 0;
    {   BandEntry en; en.vp = VP; en.vn = VN; R.p[0] = en; }    // checkpoint 0: the scan's start state
    auto column = [&](uint32_t A, uint32_t B, int t, int c) {
        const uint32_t src = (t & 1) ? B : A;
        const uint32_t sel = (t >> 1) == 0 ? sel0 : (t >> 1) == 1 ? sel1 : (t >> 1) == 2 ? sel2 : sel3;
        uint32_t D0, HP, HN;
        band_advance(band_eq(peq_base, byte_perm(src, lane8, sel), sp0, c), VP, VN, D0, HP, HN);
        const uint32_t bm = (uint32_t)(bm0 - c) & 31u;          // meaningful on the hull columns only
        const uint32_t hq = (uint32_t)(c - h0);                 // 0..31 on the hull columns
        hull_p |= shl32((HP >> bm) & 1u, hq);
        hull_n |= shl32((HN >> bm) & 1u, hq);
        if ((t & (BAND_CKPT - 1)) == BAND_CKPT - 1) {           // c = 8q + t + 1 is a multiple of BAND_CKPT
            BandEntry en;
            en.vp = VP; en.vn = VN;
            R.p[(c >> BAND_CKPT_LOG) * R.stride] = en;
        }
    };
    for (int q = 0; q < nchunks; q++) {
        uint32_t A, B;
        rd.next(A, B);
        const int ncol = imin(8, ncols - 8 * q);
        if (ncol == 8) {
#pragma unroll
            for (int t = 0; t < 8; t++) column(A, B, t, 8 * q + t + 1);
        } else {
#pragma unroll
            for (int t = 0; t < 8; t++) if (t < ncol) column(A, B, t, 8 * q + t + 1);
        }
    }
    C.vpn = VP; C.vnn = VN;             // column we (column n when the scan reached it)
    C.hull_p = hull_p; C.hull_n = hull_n;
}

// R5 over the hull columns jf..jl of the scan just stored, exactly as resolve_columns does it while it
// scans: D[m][j] from the anchor and the stored deltas of row m; narrow hull: note the costs and the
// candidate with the best bound (band_finish walks them); otherwise walk on the spot, in column order.
ORC_HD void band_hull(const uint32_t *W, uint64_t lo, uint32_t len, BandCtx &C, const BandRing &R)
{
    const int m = C.m, n = C.n, k = C.k, ws = C.ws, jf = C.jf, jl = C.jl, dir = C.dir, boff = C.boff;
    C.traced_j = -1; C.traced_score = 0; C.traced_origin = 0;
    C.narrow = (jf <= jl && (jl - jf) + 2 * k <= m / 2) ? 1 : 0;
    C.top_j = -1;
    C.broke = 0;
    C.hc0 = ~0ull; C.hc1 = ~0ull;
    int top_ub = -(1 << 20);
    C.ubw = (ws == 0 && C.type == TYPE_BACK) ? 1 : 2;
    const int ubw = C.ubw;
    const bool narrow = C.narrow != 0;
    int D = C.c5;
    for (int j = jf; j <= jl; j++) {
        if (j > jf) D += (int)((C.hull_p >> (j - jf)) & 1u) - (int)((C.hull_n >> (j - jf)) & 1u);
        if (D > k) continue;
        const int lmax = imin(m, j + D);
        if (!(lmax >= C.min_ov && D <= (int)C.kmax[lmax])) continue;
        const int ub = lmax - ubw * D;
        {   // D[m][j] of the candidate columns: the narrow hull's walks and R6's cell (m, n) read them
            const int hq = j - jf;                              // < 32 for every band-eligible hull
            const uint64_t clr = ~(15ull << (4 * (hq & 15)));
            const uint64_t put = (uint64_t)D << (4 * (hq & 15));
            if (hq < 16) C.hc0 = (C.hc0 & clr) | put; else C.hc1 = (C.hc1 & clr) | put;
        }
        if (narrow) {
            if (ub > top_ub) { top_ub = ub; C.top_j = j; }      // highest bound, leftmost
        } else if (C.best.cost == m + n + 1 || ub > C.best.score) {
            // a candidate that cannot beat the best so far cannot change it (every R5 update after the
            // first needs a strictly higher score)
            Cell c;
            c.cost = D;
            band_trace_back(C, W, lo, len, dir, C.code4, C.rcode4, m, C.type, ws, R, boff, m, j, D, c.score, c.origin);
            C.traced_j = j; C.traced_score = c.score; C.traced_origin = c.origin;
            if (r5_update(C.best, m, n, c, j, C.min_ov, C.kmax)) { C.broke = 1; return; }
        }
    }
}

ORC_HD void band_finish(const uint32_t *W, uint64_t lo, uint32_t len, BandCtx &C, const BandRing &R,
                        uint32_t mask, bool grouped)
{
    const int m = C.m, n = C.n, k = C.k, ws = C.ws, dir = C.dir, ubw = C.ubw;
    Best &best = C.best;
    if (grouped) {
        bool more = !C.broke && C.narrow && C.top_j >= 0;
        int q = C.jf - 1;
        while (ORC_ANY(mask, more)) {
            int j = 0, Dj = 0;
            bool have = false;
            if (more) {
                for (; q <= C.jl && !have; q++) {
                    j = q < C.jf ? C.top_j : q;
                    Dj = band_hull_cost(C, j);
                    if (q < C.jf) { have = true; continue; }
                    if (j == C.top_j || Dj > k) continue;       // 15 marks "not a candidate"
                    const int lmax = imin(m, j + Dj);
                    const int ub = lmax - ubw * Dj;
                    have = best.cost == m + n + 1 || ub > best.score || (ub == best.score && j < best.query_stop);
                }
                if (!have) more = false;
            }
            Cell c;
            c.cost = Dj; c.score = 0; c.origin = 0;
            band_trace_back_group(C, mask, have, W, lo, len, dir, C.code4, C.rcode4, m, C.type, ws, R, C.boff, m, j, Dj,
                                  c.score, c.origin);
            if (have) {
                if (j == n) { C.traced_j = j; C.traced_score = c.score; C.traced_origin = c.origin; }
                const int length = m + imin(c.origin, 0);
                if (length >= C.min_ov && c.cost <= (int)C.kmax[KMAX_R5 + length] &&
                    (best.cost == m + n + 1 || c.score > best.score || (c.score == best.score && j < best.query_stop))) {
                    best.score = c.score; best.cost = c.cost; best.origin = c.origin; best.ref_stop = m; best.query_stop = j;
                }
            }
        }
        if (!C.broke && C.narrow && best.cost == 0 && best.origin >= 0) C.broke = 1;     // R5's early exit
    }
    if (C.broke) return;
    if (C.r6 && C.we == n && n > ws) {
        // R6: rows r6hi..r6lo of column n, top row first like cutadapt.  D[i][n] counts on from the scan's
        // D[r6lo][n] (3' adapters) or is the last hull column's cost (5' adapters: the single cell (m, n)).
        const int ilo = imax(C.G.r6lo, 1), ihi = imin(C.G.r6hi, m);
        int Dtop = 0;
        if (C.type == TYPE_FRONT) Dtop = band_hull_cost(C, n);  // only reached with jl == n
        else {
            Dtop = C.c6;
            for (int i = ilo + 1; i <= ihi; i++) {
                const uint32_t b = (uint32_t)(i - (n - C.boff));
                Dtop += (int)((C.vpn >> b) & 1u) - (int)((C.vnn >> b) & 1u);
            }
        }
        int Di = Dtop;
        for (int i = ihi; i >= ilo; i--) {
            if (i < ihi) {                                      // D[i][n] = D[i+1][n] - the vertical delta of row i+1
                const uint32_t b = (uint32_t)(i + 1 - (n - C.boff));
                Di -= (int)((C.vpn >> b) & 1u) - (int)((C.vnn >> b) & 1u);
            }
            if (Di > k) continue;
            const int lmax = (C.type == TYPE_FRONT) ? imin(i, n + Di) : i;
            if (!(lmax >= C.min_ov && Di <= (int)C.kmax[lmax])) continue;
            const int ub = lmax - ubw * Di;
            if (ub < best.score || (ub == best.score && Di >= best.cost)) continue;
            Cell c;
            c.cost = Di;
            if (i == m && C.traced_j == n) { c.score = C.traced_score; c.origin = C.traced_origin; }
            else band_trace_back(C, W, lo, len, dir, C.code4, C.rcode4, m, C.type, ws, R, C.boff, i, n, Di, c.score, c.origin);
            r6_update(best, n, c, i, C.min_ov, C.kmax);
        }
    }
}

ORC_HD void band_begin(const uint32_t *W, const View &v, int type, int n_adapters, const BandAdapter *ads,
                       const Task &t, BandCtx &C, BandRing &R, const char *peq_base)
{
    const BandAdapter &A = ads[(int)t.lane % n_adapters];
    C.dir = (int)t.lane / n_adapters;
    C.m = A.m; C.k = A.k; C.min_ov = A.min_ov;
    C.kmax = A.kmax[0];
    C.peq_base = peq_base;
    C.lane = (int)t.lane;
    C.code4 = A.code4; C.rcode4 = A.rcode4;
    C.n = (int)v.len;
    C.type = type;
    const int m = C.m, n = C.n;
    C.best.ref_stop = m; C.best.query_stop = n; C.best.cost = m + n + 1; C.best.origin = 0; C.best.score = 0;
    resolve_geometry(type, m, C.k, n, t, C.G);
    C.broke = 0; C.narrow = 0; C.top_j = -1; C.ubw = 2; C.boff = 0;
    C.traced_j = -1; C.traced_score = 0; C.traced_origin = 0;
    C.hc0 = C.hc1 = ~0ull; C.vpn = C.vnn = 0; C.hull_p = C.hull_n = 0; C.sp0 = 0;
    C.c5 = (int)((uint32_t)t.pad_ >> 8) & 63; C.c6 = (int)((uint32_t)t.pad_ >> 16) & 63;
    C.ws = C.G.ws; C.we = C.G.we; C.jf = C.G.jf; C.jl = C.G.jl; C.r6 = C.G.r6;
    if (C.G.has5 || C.G.has6) {
        band_scan_ok(m, C.k, n, C.ws, C.we, C.jf, C.jl, C.r6 != 0, C.G.r6lo, C.G.r6hi, C.boff);
        band_columns(W, v.lo, v.len, C, R);
    }
}

ORC_HD void band_end(const uint32_t *W, const View &v, BandCtx &C, PairResult &res, BandRing &R,
                     uint32_t mask = 0xffffffffu)
{
    band_hull(W, v.lo, v.len, C, R);
    band_finish(W, v.lo, v.len, C, R, mask, true);
    if (C.G.second && !C.broke) {
        C.ws = C.G.ws6; C.we = C.n; C.jf = 1; C.jl = 0; C.r6 = 1;
        band_scan_ok(C.m, C.k, C.n, C.ws, C.we, 1, 0, true, C.G.r6lo, C.G.r6hi, C.boff);
        band_columns(W, v.lo, v.len, C, R);
        band_hull(W, v.lo, v.len, C, R);
        band_finish(W, v.lo, v.len, C, R, mask, false);
    }
    best_to_result(C.best, C.m, C.n, res);
}

// band_table: the table band_table_entry() describes, all banks of the round; ads: one BandAdapter per
// adapter (the host simulation builds both per round)
ORC_HD void band_resolve_pair(const uint32_t *W, const View &v, const RoundTable &T, const BandAdapter *ads, const Task &t,
                              PairResult &res, BandRing &R, const char *band_table)
{
    BandCtx C;
    band_begin(W, v, T.type, T.n_adapters, ads, t, C, R, band_table + (size_t)(t.lane >> 5) * BAND_BANK_BYTES);
    band_end(W, v, C, res, R);
}

// ------------------------------------------------------------------------------------
// select_read: R9, R10 from the per-orientation winners of R8.  key[o] is the maximum of
// pack_key() over the pairs of logical orientation o that matched (0 = none).
// ------------------------------------------------------------------------------------
// action 0: --action=trim (the adapter and what lies beyond it go); 1: --action=retain (the adapter stays:
// AdapterCutter's "retain" keeps read[match.rstart:] of a 5' match and read[:match.rstop] of a 3' match)
ORC_HD void select_read(int type, int revcomp, const View &v, const uint64_t key[2], const PairResult *results,
                        Match &out, View &next, int action = 0)
{
    const int fs = key[0] ? (int)(key[0] >> 44) - 512 : 0;
    const int rs = key[1] ? (int)(key[1] >> 44) - 512 : 0;
    const int o = (revcomp && rs > fs) ? 1 : 0;            // R9: strictly higher score
    next = v;
    const uint32_t eff = (v.rc & 1u) ^ (uint32_t)o;
    if (!key[o]) {
        // no match.  o == 1 here means reverse_score 0 > forward_score < 0 (high error rates):
        // cutadapt then passes on the reverse complement, name + " rc", with no match.
        out.adapter = -1; out.is_rc = o;
        out.ref_start = out.ref_stop = out.query_start = out.query_stop = out.score = out.errors = 0;
        next.rc = ((v.rc & ~1u) | eff) + ((uint32_t)o << 8);
        return;
    }
    const PairResult r = results[(uint32_t)key[o]];
    out.adapter = 63 - (int)((key[o] >> 32) & 63u); out.is_rc = o;
    out.ref_start = r.ref_start; out.ref_stop = r.ref_stop;
    out.query_start = r.query_start; out.query_stop = r.query_stop;
    out.score = r.score; out.errors = r.errors;
    // R10: FRONT keeps [query_stop, n), BACK keeps [0, query_start) of the chosen orientation
    const uint32_t n = v.len;
    uint32_t a0, b0;
    if (type == TYPE_FRONT) { a0 = (uint32_t)(action ? r.query_start : r.query_stop); b0 = n; }
    else { a0 = 0; b0 = (uint32_t)(action ? r.query_stop : r.query_start); }
    if (b0 < a0) b0 = a0;
    next.len = b0 - a0;
    next.lo = eff ? v.lo + (n - b0) : v.lo + a0;
    next.rc = ((v.rc & ~1u) | eff) + ((uint32_t)o << 8);
}

// ------------------------------------------------------------------------------------
// Anchored adapters without indels (-g ^file: / -a file$: with --no-indels; BASELINE config 4).
// cutadapt does not align these: PrefixComparer/SuffixComparer count mismatches of the first /
// last m characters, and >= 2 such adapters with k <= 2 are looked up in a dict of their Hamming
// neighbourhoods (adapters.py IndexedPrefixAdapters).  Only the m characters at
// the anchored end of a read are touched.
// ------------------------------------------------------------------------------------
constexpr int MAX_ANCH = 64;
struct AnchoredTable {
    int32_t n_adapters;
    int32_t suffix;                 // 0: 5' anchored (prefix), 1: 3' anchored (suffix)
    int32_t revcomp;
    int32_t indexed;                // dict semantics (>= 2 adapters, one length, every k <= 2)
    int32_t one_length;             // every adapter has m[0] characters: the packed comparison below applies
    int32_t pad_;
    uint64_t nib[MAX_ANCH][4];      // the adapter as one-hot nibbles (A1 C2 G4 T8), 16 characters per word, 0 behind m
    int32_t m[MAX_ANCH];
    int32_t k[MAX_ANCH];            // int(rate * m)
    uint8_t seq[MAX_ANCH][MAX_M];   // upper-case ASCII
};

// Best adapter for logical orientation o of the view.  seq: ASCII bases; comp: complement LUT.
// Returns the adapter index or -1 and fills res.
ORC_HD int anchored_match(const uint8_t *seq, const uint8_t *comp, const View &v, int o,
                          const AnchoredTable &T, PairResult &res)
{
    const int n = (int)v.len;
    const uint32_t eff = (v.rc & 1u) ^ (uint32_t)o;
    res.has = 0; res.pad_ = 0;
    res.ref_start = res.ref_stop = res.query_start = res.query_stop = res.score = res.errors = 0;
    int best = -1, best_score = 0, best_e = 0, best_m = 0;
    bool use_index = T.indexed != 0;            // (build_anchored_table: indexed implies one_length)
    // One length for all adapters (the M13 indices): the anchored end of the read is gathered ONCE as one-hot
    // nibbles, and every adapter is one AND + fold + popcount per 16 characters away -- a character matches iff
    // the nibbles share their bit (anything but ACGT has none, like the ASCII comparison against ACGT adapters).
    uint64_t rn0 = 0, rn1 = 0, rn2 = 0, rn3 = 0;
    if (T.one_length) {
        const int m = T.m[0];
        if (n < m) return -1;
        bool has_n = false, other = false;
        for (int i = 0; i < m; i++) {
            const int p = T.suffix ? n - m + i : i;
            uint8_t c = eff ? comp[seq[v.lo + (uint64_t)(n - 1 - p)]] : seq[v.lo + (uint64_t)p];
            if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
            const uint64_t code = c == 'A' ? 1u : c == 'C' ? 2u : c == 'G' ? 4u : c == 'T' ? 8u : 0u;
            if (c == 'N') has_n = true;
            else if (code == 0u) other = true;
            const uint64_t put = code << (4 * (i & 15));
            if (i < 16) rn0 |= put; else if (i < 32) rn1 |= put; else if (i < 48) rn2 |= put; else rn3 |= put;
        }
        if (use_index) {
            if (has_n) use_index = false;       // "N" in the affix: the plain comparer loop decides
            else if (other) return -1;          // not a key of the dict
        }
    }
    for (int a = 0; a < T.n_adapters; a++) {
        const int m = T.m[a];
        if (n < m) continue;                    // anchored adapters need the whole adapter (min_overlap = m)
        int e = 0;
        if (T.one_length) {
            int matches = 0;
            const uint64_t rn[4] = {rn0, rn1, rn2, rn3};
#pragma unroll
            for (int w = 0; w < 4; w++) {
                if (16 * w >= m) break;
                uint64_t x = rn[w] & T.nib[a][w];
                x |= x >> 1; x |= x >> 2;
                matches += popc64(x & 0x1111111111111111ull);
            }
            e = m - matches;
        } else {
            for (int i = 0; i < m; i++) {
                const int p = T.suffix ? n - m + i : i;
                uint8_t c = eff ? comp[seq[v.lo + (uint64_t)(n - 1 - p)]] : seq[v.lo + (uint64_t)p];
                if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
                e += (c != T.seq[a][i]);
            }
        }
        if (e > T.k[a]) continue;
        bool take;
        int score;
        if (use_index) {                        // most matches; the later adapter wins equal matches
            score = m - e;
            take = best < 0 || !(m - e < best_m - best_e);
        } else {                                // R8 over the comparers: score, errors, file order
            score = (m - e) - e;
            take = best < 0 || score > best_score || (score == best_score && e < best_e);
        }
        if (take) { best = a; best_score = score; best_e = e; best_m = m; }
    }
    if (best < 0) return -1;
    res.has = 1;
    res.ref_start = 0; res.ref_stop = best_m;
    res.query_start = T.suffix ? n - best_m : 0;
    res.query_stop = T.suffix ? n : best_m;
    res.score = best_score; res.errors = best_e;
    return best;
}

// ------------------------------------------------------------------------------------
// Adapters longer than one 64-bit word ("adapters > 64 nt"): a round that holds one takes this
// path instead of the bit-parallel scan.  It is cutadapt's own recurrence, cell by cell (_align.pyx
// Aligner.locate: R2 first column, R3 cell rule with its tie order match / mismatch / insertion / deletion,
// R4 Ukkonen's `last`, R5 last-row test with the early stop, R6 last-column test over whatever the column
// holds above `last`), one column of (cost, score, origin) entries per thread -- exact by construction and
// slow (m x n cells per pair where the bit-parallel path skips 96 % of them): long adapters are rare, and
// a pipeline that meets one should not have to leave the GPU.
// ------------------------------------------------------------------------------------
constexpr int MAX_M_LONG = 256, MAX_AD_LONG = 16;
struct LongTable {
    int32_t n_adapters, type, revcomp, indels;
    int32_t m[MAX_AD_LONG], k[MAX_AD_LONG], min_ov[MAX_AD_LONG];
    uint8_t code[MAX_AD_LONG][MAX_M_LONG];          // 4-bit IUPAC masks (A1 C2 G4 T8)
    // largest cost that is acceptable for an overlap of L adapter characters, (double)cost <= effective length *
    // rate taken in fp64 on the host: kmax5 for the last-row test (R5), kmax6 for the last-column test (R6) --
    // they differ only for 5' adapters with N wildcards (see build_round_table)
    uint8_t kmax5[MAX_AD_LONG][MAX_M_LONG + 4];
    uint8_t kmax6[MAX_AD_LONG][MAX_M_LONG + 4];
};
struct LongCell { int32_t cost, score, origin; };

// Aligner.locate of adapter a against the view in storage direction dir.  col: m + 1 entries of scratch.
ORC_HD void long_locate(const uint32_t *W, const View &v, int dir, const LongTable &T, int a, LongCell *col,
                        PairResult &res)
{
    const int m = T.m[a], k = T.k[a], n = (int)v.len, min_ov = T.min_ov[a];
    const int ic = T.indels ? 1 : 100000;                   // --no-indels prices them out (_align.pyx)
    const bool front = T.type == TYPE_FRONT;
    const uint8_t *code = T.code[a], *kmax5 = T.kmax5[a], *kmax6 = T.kmax6[a];
    for (int i = 0; i <= m; i++) {                          // R2: both adapter types may start anywhere in the read
        col[i].score = 0;
        if (front) { col[i].cost = 0; col[i].origin = -i; }       // ... and a 5' adapter anywhere in itself
        else { col[i].cost = i * ic; col[i].origin = 0; }
    }
    Best best;
    best.ref_stop = m; best.query_stop = n; best.cost = m + n + 1; best.origin = 0; best.score = 0;
    int last = front ? m : imin(m, k + 1);                  // R4
    for (int j = 1; j <= n; j++) {
        LongCell diag = col[0];
        col[0].origin = j;
        LongCell up = col[0];
        const uint32_t rc = lane_code(W, v.lo, v.len, dir, j - 1);
        for (int i = 1; i <= last; i++) {
            const LongCell cur = col[i];
            LongCell nw;
            if (code[i - 1] & rc) {
                nw.cost = diag.cost; nw.origin = diag.origin; nw.score = diag.score + 1;
            } else {
                const int c_diag = diag.cost + 1, c_del = cur.cost + ic, c_ins = up.cost + ic;
                if (c_diag <= c_del && c_diag <= c_ins) { nw.cost = c_diag; nw.origin = diag.origin; nw.score = diag.score - 1; }
                else if (c_ins <= c_del) { nw.cost = c_ins; nw.origin = up.origin; nw.score = up.score - 2; }
                else { nw.cost = c_del; nw.origin = cur.origin; nw.score = cur.score - 2; }
            }
            diag = cur;
            col[i] = nw;
            up = nw;
        }
        while (last >= 0 && col[last].cost > k) last--;
        if (last < m) { last++; continue; }
        // R5: the last row holds a cell within k errors
        const LongCell c = col[m];
        const int length = m + imin(c.origin, 0);
        if (!(length >= min_ov && c.cost <= (int)kmax5[length])) continue;
        const int best_length = m + imin(best.origin, 0);
        if (best.cost == m + n + 1 || (c.origin <= best.origin + m / 2 && c.score > best.score) ||
            (length > best_length && c.score > best.score)) {
            best.score = c.score; best.cost = c.cost; best.origin = c.origin; best.ref_stop = m; best.query_stop = j;
            if (c.cost == 0 && c.origin >= 0) break;        // exact full match: cutadapt stops scanning
        }
    }
    // R6: the last column -- the column as it stands (after an early stop too, as in _align.pyx; rows above
    // `last` hold what an earlier column left there), every row of a 3' adapter, row m of a 5' adapter
    for (int i = m; i >= (front ? m : 0); i--) {
        const LongCell c = col[i];
        const int length = i + imin(c.origin, 0);
        if (!(length >= min_ov && c.cost <= (int)kmax6[length])) continue;
        if (c.score > best.score || (c.score == best.score && c.cost < best.cost)) {
            best.score = c.score; best.cost = c.cost; best.origin = c.origin; best.ref_stop = i; best.query_stop = n;
        }
    }
    best_to_result(best, m, n, res);
}

// Best adapter (R8: score, then fewer errors, then file order) for logical orientation o of the view.
// Returns the adapter index or -1 and fills res.
ORC_HD int long_match(const uint32_t *W, const View &v, int o, const LongTable &T, LongCell *col, PairResult &res)
{
    const int dir = (int)((v.rc & 1u) ^ (uint32_t)o);
    int best = -1;
    res.has = 0; res.pad_ = 0;
    res.ref_start = res.ref_stop = res.query_start = res.query_stop = res.score = res.errors = 0;
    for (int a = 0; a < T.n_adapters; a++) {
        PairResult r;
        long_locate(W, v, dir, T, a, col, r);
        if (!r.has) continue;
        if (best < 0 || r.score > res.score || (r.score == res.score && r.errors < res.errors)) { best = a; res = r; }
    }
    return best;
}

}  // namespace orc
