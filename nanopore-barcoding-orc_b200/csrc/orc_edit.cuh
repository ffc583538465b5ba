// orc_edit.cuh -- pairwise unit-cost edit distance of whole reads, the arithmetic amplicon_sorter
// delegates to edlib ("next" row N3):
//
//   /root/reference/scripts/auxiliary_code/amplicon_sorter.py:225-235   distance()          mode NW
//   /root/reference/scripts/auxiliary_code/amplicon_sorter.py:838-849   distance_finetune() mode HW
//       edlib.align(shorter, longer, task='distance', mode=...)['editDistance']
//
// Myers' bit-vector algorithm in blocks of 64 query rows (the formulation edlib itself uses): a
// block takes the horizontal delta hin in {-1, 0, +1} that enters its top row in a column and
// hands the one leaving its bottom row to the block below.  One warp works on one pair: lane l
// owns WB consecutive blocks and, at step t, is in column t - l, so the delta a lane needs from
// the lane above was produced one step earlier and travels by a single shuffle -- the warp sweeps
// the matrix as a skewed wavefront, every lane busy except while the pipeline fills and drains.
// NW: row 0 costs j (hin = +1 into block 0), the answer is D[m][n].  HW: row 0 is free (hin = 0),
// the answer is the minimum of D[m][j] over the columns.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ORC_EHD __host__ __device__ __forceinline__
#else
#define ORC_EHD inline
#endif

namespace orc {

constexpr int EDIT_SYMS = 8;            // distinct bytes a batch may contain (ACGT, N, three more)
constexpr int EDIT_MAX_WB = 4;          // blocks per lane: queries up to 32 * 4 * 64 = 8192 rows

// One block, one column.  Returns hout; hrow = the horizontal delta at bit `rbit` (the query's
// last row when this is its last block).
template <bool ROW = true>
ORC_EHD int myers_block_step(uint64_t &Pv, uint64_t &Mv, uint64_t Eq, int hin, int rbit, int &hrow)
{
    const uint64_t neg = hin < 0 ? 1ull : 0ull;
    const uint64_t Xv = Eq | Mv;
    Eq |= neg;
    const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
    uint64_t Ph = Mv | ~(Xh | Pv);
    uint64_t Mh = Pv & Xh;
    const int hout = (int)(Ph >> 63) - (int)(Mh >> 63);
    if (ROW) hrow = (int)((Ph >> rbit) & 1ull) - (int)((Mh >> rbit) & 1ull);
    else hrow = 0;
    Ph <<= 1; Mh <<= 1;
    Mh |= neg;
    Ph |= hin > 0 ? 1ull : 0ull;
    Pv = Mh | ~(Xv | Ph);
    Mv = Ph & Xv;
    return hout;
}

// match bits of query rows 64 blk .. 64 blk + 63 against symbol s
ORC_EHD uint64_t edit_eq_word(const uint8_t *q, uint32_t m, uint32_t blk, uint32_t s)
{
    uint64_t w = 0;
    const uint32_t base = 64u * blk;
    for (uint32_t i = 0; i < 64u && base + i < m; i++)
        if (q[base + i] == s) w |= 1ull << i;
    return w;
}

// The same computation block after block, column after column (no skew): what the kernel must
// produce, used by the host simulation.
inline uint32_t edit_distance_blocks(const uint8_t *q, uint32_t m, const uint8_t *t, uint32_t n, int mode)
{
    if (m == 0) return mode ? 0u : n;
    const uint32_t nb = (m + 63u) / 64u;
    uint64_t *Pv = new uint64_t[nb], *Mv = new uint64_t[nb];
    for (uint32_t b = 0; b < nb; b++) { Pv[b] = ~0ull; Mv[b] = 0; }
    const int rbit = (int)((m - 1u) & 63u);
    uint32_t score = m, best = m;
    for (uint32_t j = 0; j < n; j++) {
        int hin = mode ? 0 : 1;
        for (uint32_t b = 0; b < nb; b++) {
            int hrow;
            hin = myers_block_step(Pv[b], Mv[b], edit_eq_word(q, m, b, t[j]), hin, rbit, hrow);
            if (b == nb - 1u) score = (uint32_t)((int)score + hrow);
        }
        if (score < best) best = score;
    }
    delete[] Pv;
    delete[] Mv;
    return mode ? best : score;
}

#if defined(__CUDACC__)
// sym: the batch's bytes mapped to 0 .. EDIT_SYMS-1.  G lanes work on one pair (G = 5 .. 32, any
// size: short queries leave most of a warp idle otherwise), so a warp sweeps 32 / G pairs side by side;
// todo lists the pairs of this (WB, G) class, longest target first, so that the pairs of a warp
// take about the same number of steps.  Dynamic shared memory: per warp EDIT_SYMS * WB * 32 match
// words, laid out [symbol][k][lane]; a lane only ever reads its own.
template <int WB, int G>
__global__ void __launch_bounds__(128, 8)
edit_kernel(const uint8_t *__restrict__ sym, const uint64_t *__restrict__ off, const uint32_t *__restrict__ len,
            const uint32_t *__restrict__ pair_a, const uint32_t *__restrict__ pair_b,
            const uint32_t *__restrict__ todo, uint32_t n_todo, int mode, uint32_t *__restrict__ out)
{
    extern __shared__ uint64_t s_eq_all[];
    constexpr uint32_t PER_WARP = 32u / G;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gl = lane % G, grp = lane / G;
    uint64_t *s_eq = s_eq_all + (size_t)(threadIdx.x >> 5) * (EDIT_SYMS * WB * 32);
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n_rounds = (n_todo + PER_WARP - 1u) / PER_WARP;
    for (uint32_t w = warp; w < n_rounds; w += n_warps) {
        const uint32_t slot = w * PER_WARP + grp;
        const bool have = grp < PER_WARP && slot < n_todo;       // lanes past the last whole group idle
        const uint32_t pair = have ? todo[slot] : 0u;
        uint32_t a = have ? pair_a[pair] : 0u, b = have ? pair_b[pair] : 0u;
        uint32_t m = 0, n = 0;
        const uint8_t *q = sym, *t = sym;
        if (have) {
            if (len[a] > len[b]) { const uint32_t x = a; a = b; b = x; }   // the longer one is the target
            m = len[a]; n = len[b];
            q = sym + off[a]; t = sym + off[b];
        }
        const bool work = have && m > 0;
        if (have && m == 0 && gl == 0) out[pair] = mode ? 0u : n;
        const uint32_t nb = (m + 63u) >> 6;
        const uint32_t lanes_used = (nb + WB - 1) / WB;
        __syncwarp();
        // the lane's match words: every query byte is read once and sets its bit in one of the
        // EDIT_SYMS words of its block (two 32-bit halves)
#pragma unroll
        for (int k = 0; k < WB; k++) {
            const uint32_t base = 64u * (gl * WB + k);
            uint32_t lo[EDIT_SYMS], hi[EDIT_SYMS];
#pragma unroll
            for (int s = 0; s < EDIT_SYMS; s++) { lo[s] = 0; hi[s] = 0; }
            if (base < m) {
#pragma unroll 2
                for (int i = 0; i < 32; i++) {
                    const uint32_t c0 = base + i < m ? q[base + i] : 0xFFu;
                    const uint32_t c1 = base + 32u + i < m ? q[base + 32u + i] : 0xFFu;
#pragma unroll
                    for (int s = 0; s < EDIT_SYMS; s++) {
                        lo[s] |= (c0 == (uint32_t)s ? 1u : 0u) << i;
                        hi[s] |= (c1 == (uint32_t)s ? 1u : 0u) << i;
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < EDIT_SYMS; s++)
                s_eq[(s * WB + k) * 32 + lane] = ((uint64_t)hi[s] << 32) | lo[s];
        }
        __syncwarp();
        uint64_t Pv[WB], Mv[WB];
#pragma unroll
        for (int k = 0; k < WB; k++) { Pv[k] = ~0ull; Mv[k] = 0; }
        const int rbit = (int)((m - 1u) & 63u);
        const uint32_t last_gl = work ? (nb - 1u) / WB : 0u;
        const int last_k = work ? (int)((nb - 1u) % WB) : 0;
        int score = (int)m, best = (int)m;
        int carry = 0;                                  // hout of this lane's last block, previous step
        const uint32_t my_steps = work ? n + lanes_used - 1u : 0u;
        const uint32_t steps = __reduce_max_sync(0xffffffffu, my_steps);
        const bool mine = work && gl < lanes_used;
        // column of this lane at step s: s - gl; its symbol is loaded a step ahead
        const uint8_t *tp = t - gl;
        uint32_t c_next = (mine && gl == 0 && n > 0) ? t[0] : 0u;
        if (mode) {
            // HW: D[m][j] of every column, through the horizontal delta at the query's last row
            for (uint32_t step = 0; step < steps; step++) {
                const int from_above = __shfl_up_sync(0xffffffffu, carry, 1);      // lane 0 of a group ignores it
                const uint32_t j = step - gl;                   // wraps below column 0: >= n
                const uint32_t c = c_next;
                if (mine && j + 1u < n) c_next = tp[step + 1u];
                if (mine && j < n) {
                    int hin = gl == 0 ? 0 : from_above;
#pragma unroll
                    for (int k = 0; k < WB; k++) {
                        if (gl * WB + k < nb) {
                            int hrow;
                            hin = myers_block_step<true>(Pv[k], Mv[k], s_eq[(c * WB + k) * 32 + lane], hin, rbit, hrow);
                            if (gl == last_gl && k == last_k) {
                                score += hrow;
                                best = min(best, score);
                            }
                        }
                    }
                    carry = hin;
                }
            }
        } else {
            // NW: only D[m][n] is wanted, and that is n plus the vertical deltas of the last column
            for (uint32_t step = 0; step < steps; step++) {
                const int from_above = __shfl_up_sync(0xffffffffu, carry, 1);      // lane 0 of a group ignores it
                const uint32_t j = step - gl;
                const uint32_t c = c_next;
                if (mine && j + 1u < n) c_next = tp[step + 1u];
                if (mine && j < n) {
                    int hin = gl == 0 ? 1 : from_above;
#pragma unroll
                    for (int k = 0; k < WB; k++) {
                        if (gl * WB + k < nb) {
                            int hrow;
                            hin = myers_block_step<false>(Pv[k], Mv[k], s_eq[(c * WB + k) * 32 + lane], hin, rbit, hrow);
                        }
                    }
                    carry = hin;
                }
            }
            int sum = 0;
#pragma unroll
            for (int k = 0; k < WB; k++) {
                const uint32_t blk = gl * WB + k;
                if (mine && blk < nb) {
                    const uint64_t keep = (blk == nb - 1u && rbit < 63) ? ((1ull << (rbit + 1)) - 1ull) : ~0ull;
                    sum += __popcll(Pv[k] & keep) - __popcll(Mv[k] & keep);
                }
            }
            // inclusive prefix sum inside the group: the lane that writes the result (the last one
            // with a block) then holds the total
#pragma unroll
            for (int d = 1; d < G; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, sum, d);
                if (gl >= (uint32_t)d) sum += v;
            }
            score = (int)n + sum;
        }
        if (work && gl == last_gl) out[pair] = (uint32_t)(mode ? best : score);
    }
}
#endif

}  // namespace orc
