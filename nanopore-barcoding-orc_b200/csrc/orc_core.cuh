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
//  resolve_pair()  For a pair that has candidate cells: cutadapt's own recurrence
//                  (cost, score, origin per cell, R2/R3, same tie-breaks) restricted to
//                  the band of diagonals within k of a candidate, then R5/R6/R7 verbatim.
//                  DESIGN.md section 4 proves this reproduces the unrestricted DP exactly.
//  select_read()   R8 (best of the adapters) and R9 (--rc: strictly higher score wins),
//                  R10 (trim -> the next view of the read).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ORC_HD __host__ __device__ __forceinline__
#else
#define ORC_HD inline
#endif

namespace orc {

constexpr int MAX_AD = 16;          // adapters per round
constexpr int MAX_M = 64;           // adapter length (one 64-bit word)
constexpr int MAX_LANES = 32;       // 2 orientations x MAX_AD
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
    int32_t pad_[3];
    int32_t m[MAX_AD];              // adapter length
    int32_t k[MAX_AD];              // int(max_error_rate * m)
    int32_t min_ov[MAX_AD];         // min(min_overlap, m)
    uint8_t kmax[MAX_AD][MAX_M + 8];// kmax[a][L] = max cost with cost <= L * rate (fp64, R5/R6)
    uint8_t code[MAX_AD][MAX_M];    // adapter as 4-bit IUPAC masks (A1 C2 G4 T8)
    uint64_t peq[16][MAX_LANES];    // [read code][lane]: match bits, row i at bit 64-m+i-1,
                                    // the 64-m low padding bits always 1
    uint64_t pv0[MAX_LANES];        // vertical deltas of column 0 (R2)
    int32_t d0[MAX_LANES];          // D[m][0]
};

// A read (or what a previous round left of it) as a window of the packed code array:
// element p of the view is code[lo+p], or, if rc is set, comp(code[lo+len-1-p]).
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
    int32_t pad_[2];
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

// ------------------------------------------------------------------------------------
// scan_lane: the hot loop.  W: packed codes (8 per word) with guard words on both sides.
// peq_base: byte address of RoundTable::peq (shared memory on the device); the entry of
// (code c, lane l) lives at c*256 + l*8, which one PRMT builds from a code byte and l*8.
// ------------------------------------------------------------------------------------
struct ScanHull { int32_t jf, jl, i1, i2; };

ORC_HD void scan_lane(const uint32_t *__restrict__ W, uint64_t lo, uint32_t len, int dir,
                      const char *peq_base, int lane, uint64_t Pv, int D, int m, int k,
                      const uint8_t *kmax, int min_ov, int type, ScanHull &out)
{
    uint64_t Mv = 0;
    int jf = 0x7fffffff, jl = -1;
    const uint32_t lane8 = (uint32_t)lane * 8u;
    // PRMT selectors: result byte0 <- lane8.byte0, byte1 <- code byte b, bytes 2,3 <- 0
    uint32_t sel0, sel1, sel2, sel3;
    if (!dir) { sel0 = 0x5504u; sel1 = 0x5514u; sel2 = 0x5524u; sel3 = 0x5534u; }
    else      { sel0 = 0x5534u; sel1 = 0x5524u; sel2 = 0x5514u; sel3 = 0x5504u; }
    const uint32_t shA = dir ? 4u : 0u, shB = dir ? 0u : 4u;
    const int nchunks = (int)((len + 7u) >> 3);
    // storage index of the lowest-addressed code of the current chunk
    int64_t s = dir ? (int64_t)lo + (int64_t)len - 8 : (int64_t)lo;
    const int64_t step = dir ? -8 : 8;
    int j = 0;
    for (int q = 0; q < nchunks; q++, s += step) {
        const int64_t wi = s >> 3;
        const uint32_t x = funnel_r(W[wi], W[wi + 1], (uint32_t)(s & 7) * 4u);
        const uint32_t A = (x >> shA) & 0x0F0F0F0Fu;   // codes at even view positions of the chunk
        const uint32_t B = (x >> shB) & 0x0F0F0F0Fu;   // codes at odd view positions
        const int ncol = imin(8, (int)len - 8 * q);
#pragma unroll
        for (int t = 0; t < 8; t++) {
            if (t >= ncol) break;
            const uint32_t src = (t & 1) ? B : A;
            const uint32_t sel = (t >> 1) == 0 ? sel0 : (t >> 1) == 1 ? sel1 : (t >> 1) == 2 ? sel2 : sel3;
            const uint32_t addr = byte_perm(src, lane8, sel);
            const uint64_t Eq = *reinterpret_cast<const uint64_t *>(peq_base + addr);
            // Myers 1999 / Hyyro 2003, one column
            const uint64_t Xv = Eq | Mv;
            const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
            uint64_t Ph = Mv | ~(Xh | Pv);
            uint64_t Mh = Pv & Xh;
            D += (int)(Ph >> 63) - (int)(Mh >> 63);
            Ph <<= 1;            // row 0 never changes (QUERY_START): shift in 0
            Mh <<= 1;
            Pv = Mh | ~(Xv | Ph);
            Mv = Ph & Xv;
            ++j;
            if (D <= k) {
                // R5 necessary condition: a path with D errors ending in (m, j) aligns at most
                // min(m, j + D) adapter characters.
                const int lmax = imin(m, j + D);
                if (lmax >= min_ov && D <= (int)kmax[lmax]) {
                    jf = imin(jf, j);
                    jl = j;
                }
            }
        }
    }
    out.jf = jf; out.jl = jl;
    int i1 = 0x7fffffff, i2 = -1;
    if (type == TYPE_BACK) {
        // R6 necessary condition for the cells (i, n): origin >= 0 for BACK, so length == i
        int cum = 0;
        for (int i = 1; i <= m; i++) {
            const int bit = 64 - m + i - 1;
            cum += (int)((Pv >> bit) & 1u) - (int)((Mv >> bit) & 1u);
            if (i >= min_ov && cum <= (int)kmax[i]) { i1 = imin(i1, i); i2 = i; }
        }
    }
    out.i1 = i1; out.i2 = i2;
}

// ------------------------------------------------------------------------------------
// resolve_pair: exact (cost, score, origin) on a diagonal band, then R5/R6/R7.
// ------------------------------------------------------------------------------------
struct Best { int32_t score, cost, origin, ref_stop, query_stop; };
struct Cell { int32_t cost, score, origin; };

// R5 update (last row).  Returns true when cutadapt stops scanning (exact full match).
ORC_HD bool r5_update(Best &b, int m, int n, const Cell &c, int j, int min_ov, const uint8_t *kmax)
{
    const int length = m + imin(c.origin, 0);
    if (!(length >= min_ov && c.cost <= (int)kmax[length])) return false;
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
    if (!(length >= min_ov && c.cost <= (int)kmax[length])) return;
    if (c.score > b.score || (c.score == b.score && c.cost < b.cost)) {
        b.score = c.score; b.cost = c.cost; b.origin = c.origin; b.ref_stop = i; b.query_stop = n;
    }
}

// One pass of cutadapt's recurrence over the diagonals dlo..dhi (d = j - i).  Cells outside
// the band count as cost INF.  R5 is evaluated for columns jf..jl, R6 for rows r6lo..r6hi of
// column n.  col[] is scratch for rows 0..m.  Returns true if R5 hit the early exit.
ORC_HD bool band_pass(const uint32_t *W, uint64_t lo, uint32_t len, int dir,
                      const uint8_t *acode, int m, int type, int k, const uint8_t *kmax, int min_ov,
                      int dlo, int dhi, int jf, int jl, int r6lo, int r6hi, Best &best, Cell *col)
{
    const int n = (int)len;
    const int j0 = imax(0, dlo);
    const bool want_r6 = r6lo <= r6hi;
    int jend = want_r6 ? n : imin(n, jl);
    if (j0 == 0) {
        // R2, column min_n = 0: FRONT (REF_START|QUERY_START): cost 0, origin -i;
        //                       BACK  (QUERY_START only):       cost i, origin 0
        const int ia = imax(0, -dhi), ib = imin(m, -dlo);
        for (int i = ia; i <= ib; i++) {
            col[i].cost = (type == TYPE_FRONT) ? 0 : i;
            col[i].score = 0;
            col[i].origin = (type == TYPE_FRONT) ? -i : 0;
        }
    }
    int j = j0;
    int ilo = 1, ihi = 0;
    for (j = j0 + 1; j <= jend; j++) {
        ilo = imax(1, j - dhi);
        ihi = imin(m, j - dlo);
        if (ilo > ihi) break;                       // the band has left the matrix (j - dhi > m)
        const uint32_t cj = lane_code(W, lo, len, dir, j - 1);
        Cell dg, up;
        if (ilo == 1) {                             // row 0: cost 0, score 0, origin = column
            dg.cost = 0; dg.score = 0; dg.origin = j - 1;
            up.cost = 0; up.score = 0; up.origin = j;
        } else {
            dg = col[ilo - 1];
            up.cost = INF_COST; up.score = 0; up.origin = 0;
        }
        const int left_max = j - 1 - dlo;           // rows that were inside the band in column j-1
        for (int i = ilo; i <= ihi; i++) {
            Cell lf;
            if (i <= left_max) lf = col[i];
            else { lf.cost = INF_COST; lf.score = 0; lf.origin = 0; }
            Cell nw;
            if ((acode[i - 1] & cj) != 0) {         // R3: characters equal -> diagonal, always
                nw.cost = dg.cost; nw.origin = dg.origin; nw.score = dg.score + 1;
            } else {
                const int c_diag = dg.cost + 1, c_del = lf.cost + 1, c_ins = up.cost + 1;
                if (c_diag <= c_del && c_diag <= c_ins) {
                    nw.cost = c_diag; nw.origin = dg.origin; nw.score = dg.score - 1;
                } else if (c_ins <= c_del) {
                    nw.cost = c_ins; nw.origin = up.origin; nw.score = up.score - 2;
                } else {
                    nw.cost = c_del; nw.origin = lf.origin; nw.score = lf.score - 2;
                }
            }
            dg = lf;
            col[i] = nw;
            up = nw;
        }
        if (j >= jf && j <= jl && ihi == m && col[m].cost <= k) {
            if (r5_update(best, m, n, col[m], j, min_ov, kmax)) return true;
        }
    }
    if (want_r6 && j == n + 1 && n > j0) {
        // column n is complete; rows inside the band are ilo..ihi
        for (int i = imin(r6hi, ihi); i >= imax(r6lo, ilo); i--)
            r6_update(best, n, col[i], i, min_ov, kmax);
    }
    return false;
}

ORC_HD void resolve_pair(const uint32_t *W, const View &v, const RoundTable &T, const Task &t,
                         PairResult &res, Cell *col)
{
    const int a = (int)t.lane % T.n_adapters;
    const int dir = (int)t.lane / T.n_adapters;
    const int m = T.m[a], k = T.k[a], min_ov = T.min_ov[a];
    const uint8_t *kmax = T.kmax[a];
    const uint8_t *acode = T.code[a];
    const int n = (int)v.len;
    Best best;
    best.ref_stop = m; best.query_stop = n; best.cost = m + n + 1; best.origin = 0; best.score = 0;
    bool broke = false;
    if (t.jf <= t.jl) {
        const bool r6 = (T.type == TYPE_FRONT) && (t.jl == n);   // FRONT: only cell (m, n)
        broke = band_pass(W, v.lo, v.len, dir, acode, m, T.type, k, kmax, min_ov,
                          t.jf - m - k, t.jl - m + k, t.jf, t.jl, r6 ? m : 1, r6 ? m : 0, best, col);
    }
    if (!broke && T.type == TYPE_BACK && t.i1 <= t.i2) {
        band_pass(W, v.lo, v.len, dir, acode, m, T.type, k, kmax, min_ov,
                  n - t.i2 - k, n - t.i1 + k, 1, 0, t.i1, t.i2, best, col);
    }
    if (best.cost == m + n + 1) { res.has = 0; return; }
    res.has = 1;
    if (best.origin >= 0) { res.ref_start = 0; res.query_start = best.origin; }
    else { res.ref_start = -best.origin; res.query_start = 0; }
    res.ref_stop = best.ref_stop; res.query_stop = best.query_stop;
    res.score = best.score; res.errors = best.cost;
}

// ------------------------------------------------------------------------------------
// select_read: R8, R9, R10.  `mask` has one bit per lane that produced a task; the
// results of those lanes are res[0..popc(mask)) in lane order.
// ------------------------------------------------------------------------------------
ORC_HD void select_read(const RoundTable &T, const View &v, uint32_t mask, const PairResult *res,
                        Match &out, View &next)
{
    const int na = T.n_adapters;
    int best_a[2] = {-1, -1};
    PairResult best_r[2];
    best_r[0].score = best_r[1].score = 0;
    best_r[0].errors = best_r[1].errors = 0;
    int idx = 0;
    for (int lane = 0; lane < T.n_lanes; lane++) {
        if (!((mask >> lane) & 1u)) continue;
        const PairResult r = res[idx++];
        if (!r.has) continue;
        const int a = lane % na;
        const int o = (lane / na) ^ (int)(v.rc & 1u);     // logical orientation searched by the lane
        // R8: score, then errors, then file order (lanes of one direction ascend with a)
        if (best_a[o] < 0 || r.score > best_r[o].score ||
            (r.score == best_r[o].score && r.errors < best_r[o].errors)) {
            best_a[o] = a;
            best_r[o] = r;
        }
    }
    const int fs = best_a[0] >= 0 ? best_r[0].score : 0;
    const int rs = best_a[1] >= 0 ? best_r[1].score : 0;
    const int o = (T.revcomp && rs > fs) ? 1 : 0;          // R9: strictly higher score
    next = v;
    if (best_a[o] < 0) {
        out.adapter = -1; out.is_rc = 0;
        out.ref_start = out.ref_stop = out.query_start = out.query_stop = out.score = out.errors = 0;
        return;
    }
    const PairResult &r = best_r[o];
    out.adapter = best_a[o]; out.is_rc = o;
    out.ref_start = r.ref_start; out.ref_stop = r.ref_stop;
    out.query_start = r.query_start; out.query_stop = r.query_stop;
    out.score = r.score; out.errors = r.errors;
    // R10: FRONT keeps [query_stop, n), BACK keeps [0, query_start) of the chosen orientation
    const uint32_t n = v.len;
    uint32_t a0, b0;
    if (T.type == TYPE_FRONT) { a0 = (uint32_t)r.query_stop; b0 = n; }
    else { a0 = 0; b0 = (uint32_t)r.query_start; }
    if (b0 < a0) b0 = a0;
    const uint32_t eff = (v.rc & 1u) ^ (uint32_t)o;
    next.len = b0 - a0;
    next.lo = eff ? v.lo + (n - b0) : v.lo + a0;
    next.rc = ((v.rc & ~1u) | eff) + ((uint32_t)o << 8);
}

}  // namespace orc
