// orc_kernels.cuh -- the sm_100a kernels of the demultiplexer (included by orc_api.cu).
//
//   pack_kernel        ASCII bases -> 4-bit codes, flat (code index == byte index).  HBM-bound.
//   init_views_kernel  view 0 of every read = the whole read, forward; length classes for the ordering.
//   bucket_scatter_kernel  counting sort by size class (reads by view length, items by window columns).
//   per round:
//   seed_kernel        stage 1s, one thread per read: exact 8-mer pieces of the adapters looked up in a
//                      perfect hash in shared memory -> windows of the whole-adapter alignments.
//   trigger_kernel     stage 1, one thread per (read, direction): the windows at the read's ends (or the
//                      whole flank scan when the round has no seed table), merged with the seed windows.
//   filter_kernel      stage 2a, one lane per (read, direction, adapter): 32-row Myers block test over the
//                      window columns; survivors appended to a job list.  INT32-ALU-bound, the largest
//                      share of the step.
//   scan_kernel        stage 2b, one lane per surviving pair: the 64-bit Myers/Hyyro semi-global scan
//                      (orc_core.cuh scan_lane), table in shared memory; pairs whose candidates all cost 0
//                      are settled here, the others become resolver tasks.
//   resolve_band_kernel  one thread per task: the diagonals around the candidates, scan checkpoints (8 bytes
//                      every 2nd column) in shared memory (orc_core.cuh band_*); cutadapt's score/origin
//                      by walking back, the column a walk asks about re-run from its checkpoint.
//   resolve_kernel     the same for the few tasks whose candidates spread over too many diagonals
//                      (128-bit column ring in local memory), on a side stream next to the band resolver.
//   anchored_kernel    anchored --no-indels rounds instead of all of the above.
//   long_kernel        rounds with adapters over 64 nt instead of all of the above (cutadapt's recurrence as it is).
//   select_kernel      best of the adapters, --rc choice, trim -> next view, bin id.
//   bin_count/scan/offsets/place, emit_kernel
//                      stable multi-way partition of the trimmed reads into their
//                      SP5 x SP27 bins and assembly of the FASTQ records.  HBM-bound.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>
#include <stdint.h>

#include "orc_core.cuh"

namespace orc {

constexpr int GUARD_WORDS = 4;         // zero words before and after the code array
constexpr int SCAN_THREADS = 128;
constexpr int BIN_CHUNK = 256;         // reads per warp in the partition kernels
constexpr int MAX_BINS = 512;           // bins = product of (adapters + 1) over the rounds

__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_kernel(const uint8_t *__restrict__ seq, uint32_t *__restrict__ codes, uint64_t n16,
            const uint8_t *__restrict__ lut_g)
{
    __shared__ uint8_t lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = lut_g[i];
    __syncthreads();
    const uint4 *in = reinterpret_cast<const uint4 *>(seq);
    uint2 *out = reinterpret_cast<uint2 *>(codes);
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n16;
         g += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 v = in[g];
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t r[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t acc = 0;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const uint32_t byte = (w[2 * h + (b >> 2)] >> (8 * (b & 3))) & 0xffu;
                acc |= (uint32_t)lut[byte] << (4 * b);
            }
            r[h] = acc;
        }
        out[g] = make_uint2(r[0], r[1]);
    }
}

// ------------------------------------------------------------------------------------
// Ordering by size.  Stage 1 visits the reads in order of decreasing view length and stage 2 the
// (read, direction) items in order of decreasing window columns, so that the lanes of a warp run loops of
// about the same length.  Neither needs a total order: a counting sort over SORT_BUCKETS size classes
// (exact below 1024, steps of 64 above) does, with the histogram taken by the kernel that produces the
// sizes (init_views_kernel / select_kernel for the lengths, trigger_kernel for the columns) and one
// scatter kernel per ordering.
constexpr int SORT_BUCKETS = 2048;
__device__ __forceinline__ uint32_t sort_bucket(uint32_t key)
{
    return key < 1024u ? key : 1024u + min((key - 1024u) >> 6, 1023u);
}

// Read lengths only need classes of similar loop length: steps of 16 up to 4096, steps of 256 above.  Few
// classes let a block count in shared memory and add its totals once at its end (LEN_CLASSES atomics per
// block at most), instead of a million atomics on a few hundred hot addresses.
constexpr int LEN_CLASSES = 512;
__device__ __forceinline__ uint32_t len_bucket(uint32_t len)
{
    return len < 4096u ? len >> 4 : 256u + min((len - 4096u) >> 8, 255u);
}
template <bool LEN>
__device__ __forceinline__ uint32_t class_of(uint32_t key) { return LEN ? len_bucket(key) : sort_bucket(key); }

// hist[bucket] += 1 for every lane with valid set: one atomic per distinct bucket of the warp.
// Every lane of the warp must call it.
__device__ __forceinline__ void hist_add(uint32_t *__restrict__ hist, uint32_t bucket, bool valid)
{
    const uint32_t act = __ballot_sync(0xffffffffu, valid);
    if (valid) {
        const uint32_t peers = __match_any_sync(act, bucket);
        if ((threadIdx.x & 31u) == (uint32_t)(__ffs((int)peers) - 1)) atomicAdd(hist + bucket, (uint32_t)__popc(peers));
    }
}

__global__ void __launch_bounds__(256)
init_views_kernel(const uint64_t *__restrict__ offsets, const uint32_t *__restrict__ lengths,
                  uint32_t n_reads, View *__restrict__ views, uint32_t *__restrict__ len_hist)
{
    __shared__ uint32_t s_hist[LEN_CLASSES];
    for (int i = threadIdx.x; i < LEN_CLASSES; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += gridDim.x * blockDim.x) {
        View v;
        v.lo = offsets[r];
        v.len = lengths[r];
        v.rc = 0;
        views[r] = v;
        atomicAdd(&s_hist[len_bucket(v.len)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LEN_CLASSES; i += blockDim.x)
        if (s_hist[i]) atomicAdd(len_hist + i, s_hist[i]);
}

// Element e goes to position (elements of larger classes) + (its rank inside its class); the rank comes from
// a block-local count plus one reservation per (block, class) on the class cursor, so the order inside a
// class is arbitrary -- it only decides which thread works on which read, never a result.
// FROM_VIEWS: the key of read e is its view length, 0 if the previous round left it unassigned.
template <bool FROM_VIEWS>
__global__ void __launch_bounds__(256)
bucket_scatter_kernel(const View *__restrict__ views, const Match *__restrict__ prev,
                      const uint32_t *__restrict__ keys, uint32_t n, const uint32_t *__restrict__ hist,
                      uint32_t *__restrict__ cursor, uint32_t *__restrict__ order_out,
                      uint32_t *__restrict__ keys_out)
{
    constexpr int PER = SORT_BUCKETS / 256;
    __shared__ uint32_t s_base[SORT_BUCKETS], s_cnt[SORT_BUCKETS], s_blk[SORT_BUCKETS];
    __shared__ uint32_t s_warp[8];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // s_base[b] = number of elements in classes above b (descending order)
    uint32_t h[PER], sum = 0;
#pragma unroll
    for (int i = 0; i < PER; i++) { h[i] = hist[tid * PER + i]; sum += h[i]; s_cnt[tid * PER + i] = 0; }
    uint32_t inc = sum;                      // inclusive suffix sum over the threads
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_down_sync(0xffffffffu, inc, d);
        if (lane + d < 32) inc += t;
    }
    if (lane == 0) s_warp[w] = inc;
    __syncthreads();
    uint32_t above = inc - sum;              // classes of higher threads of this warp
    for (int i = w + 1; i < 8; i++) above += s_warp[i];
#pragma unroll
    for (int i = PER - 1; i >= 0; i--) { s_base[tid * PER + i] = above; above += h[i]; }
    __syncthreads();
    uint32_t key[PER], rank[PER];
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const uint32_t e = blockIdx.x * SORT_BUCKETS + i * 256 + tid;
        key[i] = 0; rank[i] = 0;
        if (e < n) {
            if (FROM_VIEWS) key[i] = (prev != nullptr && prev[e].adapter < 0) ? 0u : views[e].len;
            else key[i] = keys[e];
            rank[i] = atomicAdd(&s_cnt[class_of<FROM_VIEWS>(key[i])], 1u);
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const uint32_t c = s_cnt[tid * PER + i];
        if (c) s_blk[tid * PER + i] = atomicAdd(cursor + tid * PER + i, c);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const uint32_t e = blockIdx.x * SORT_BUCKETS + i * 256 + tid;
        if (e < n) {
            const uint32_t b = class_of<FROM_VIEWS>(key[i]);
            const uint32_t pos = s_base[b] + s_blk[b] + rank[i];
            order_out[pos] = e;
            if (keys_out) keys_out[pos] = key[i];
        }
    }
}

// Stage 1s: one thread per read (in length order), both directions in one pass over the packed
// codes: every 8-code window is looked up in the round's seed table (orc_core.cuh seed_scan).
// The result, the windows of the whole-adapter alignments, goes to trigger_kernel, which then
// only decides the windows at the read's ends.
#ifndef SEED_BLOCKS
#define SEED_BLOCKS 6
#endif
__global__ void __launch_bounds__(256, SEED_BLOCKS)
seed_kernel(const SeedTable *__restrict__ st, const uint32_t *__restrict__ W, const View *__restrict__ views,
            const Match *__restrict__ prev, const uint32_t *__restrict__ order, uint32_t n_reads,
            SeedWins *__restrict__ out, size_t seg_stride)
{
    __shared__ __align__(16) uint32_t s_key[SEED_SLOTS];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(&st->key[0]);
        uint4 *dst = reinterpret_cast<uint4 *>(&s_key[0]);
        for (int i = threadIdx.x; i < SEED_SLOTS / 4; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_reads) return;
    const uint32_t r = order ? order[p] : p;
    if (prev != nullptr && prev[r].adapter < 0) return;     // trigger_kernel skips these reads, too
    const View v = views[r];
    // blockIdx.y = the segment of the read this thread probes (orc_core.cuh seed_segments): reads come in
    // length order, so the threads of a block have the same number of segments, give or take one
    const int g = (int)blockIdx.y, n_seg = seed_segments(v.len);
    if (g >= n_seg) return;
    uint32_t a, b;
    seed_range(v.len, g, n_seg, st->m_max, st->kt, a, b);
    SeedWins sw[2];
    seed_scan(W, v.lo, v.len, s_key, st->val, st->mult, st->list, st->need, st->kt, st->m_max, sw, a, b);
    uint4 *dst = reinterpret_cast<uint4 *>(out + (size_t)g * seg_stride + 2u * r);
    const uint4 *src = reinterpret_cast<const uint4 *>(&sw[0]);
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
}

__global__ void __launch_bounds__(128, 10)      // 48 registers instead of 56: 5 % faster; 40 (12 blocks) slower
trigger_kernel(const RoundTable *__restrict__ tab, const uint32_t *__restrict__ W,
               const View *__restrict__ views, const Match *__restrict__ prev,
               const uint32_t *__restrict__ order, uint32_t n_reads, WinList *__restrict__ wins,
               uint32_t *__restrict__ wcols,
               unsigned long long *__restrict__ col_sum, const SeedWins *__restrict__ seedwins,
               uint32_t *__restrict__ col_hist, size_t seg_stride)
{
    __shared__ __align__(16) uint32_t s_peq32[16][64];
    __shared__ __align__(16) uint32_t s_peq32s[16][64];
    __shared__ uint8_t s_kmax_any[MAX_M + 8];
    __shared__ int8_t s_first_lim[MAX_M + 32];
    __shared__ uint8_t s_lut[256];
    __shared__ int s_par[8];
    __shared__ int s_mmin, s_sfxp;
    if (threadIdx.x == 0) { s_mmin = tab->m_min; s_sfxp = tab->sfx_primary; }
    for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) {
        (&s_peq32[0][0])[i] = (&tab->peq32[0][0])[i];
        (&s_peq32s[0][0])[i] = (&tab->peq32s[0][0])[i];
    }
    for (int i = threadIdx.x; i < MAX_M + 8; i += blockDim.x) s_kmax_any[i] = tab->kmax_any[i];
    for (int i = threadIdx.x; i < MAX_M + 32; i += blockDim.x) s_first_lim[i] = tab->first_lim[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = tab->chunk_lut[i];
    if (threadIdx.x == 0) {
        s_par[0] = tab->lcp; s_par[1] = tab->k_max; s_par[2] = tab->m_max; s_par[3] = tab->type;
        s_par[4] = tab->revcomp; s_par[5] = tab->use_filter; s_par[6] = tab->lcs; s_par[7] = tab->min_ov_min;
    }
    __syncthreads();
    // threads 0..n-1 take direction 0 of the reads (in length order), threads n..2n-1 direction 1:
    // the lanes of a warp then all look at the same strand, so the ones that meet the adapter (and
    // leave the fast path to replay a chunk) do so together
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < 2u * n_reads;
    const int dir = (valid && p >= n_reads) ? 1 : 0;
    const uint32_t pr = dir ? p - n_reads : p;
    const uint32_t r = valid ? (order ? order[pr] : pr) : 0u;
    const int Lp = s_par[0], kt = s_par[1], m_max = s_par[2], type = s_par[3];
    WinList wl;
    wl.n = 0; wl.flags = 0;
    for (int i = 0; i < MAX_WIN; i++) { wl.s[i] = 0; wl.e[i] = 0; }
    uint32_t cols = 0;
    const bool skip = !valid || (prev != nullptr && prev[r].adapter < 0);
    if (!skip) {
        const View v = views[r];
        if (s_par[4] || ((dir ^ (int)(v.rc & 1u)) == 0)) {
            if (s_par[5]) {
                trigger_lane(W, v.lo, v.len, dir, reinterpret_cast<const char *>(&s_peq32[0][0]),
                             (int)(2u * (threadIdx.x & 31u)) + dir, Lp, kt, type, (uint32_t)(m_max - Lp + kt),
                             (uint32_t)(Lp + kt + 1), wl,
                             s_par[6] > 0 ? reinterpret_cast<const char *>(&s_peq32s[0][0]) : nullptr,
                             s_par[6], s_kmax_any, s_par[7], m_max, s_mmin, s_sfxp, s_first_lim, s_lut,
                             seedwins ? seedwins + (2u * r + (uint32_t)dir) : nullptr,
                             seed_segments(v.len), seg_stride);
                cols = win_columns(wl);
            } else {
                wl.n = 1; wl.s[0] = 0; wl.e[0] = v.len;     // no usable shared prefix: scan everything
                wl.flags = 1u;                               // ... and test the last-column rows too
                cols = v.len;
            }
            // an item without any column to scan has no candidate cell (R6's last-column cells are
            // only ever reached through a window that ends at n): cols == 0 drops it from stage 2
        }
    }
    {   // columns stage 2 will scan (for the executed-cells figure of the roofline)
        const uint32_t sum = __reduce_add_sync(0xffffffffu, cols);
        if ((threadIdx.x & 31) == 0 && sum) atomicAdd(col_sum, (unsigned long long)sum);
    }
    hist_add(col_hist, sort_bucket(cols), valid);      // the size classes of the stage-2 ordering
    if (!valid) return;
    const uint32_t item = r * 2u + (uint32_t)dir;
    uint4 *dst = reinterpret_cast<uint4 *>(wins + item);
    const uint4 *src = reinterpret_cast<const uint4 *>(&wl);
    dst[0] = src[0]; dst[1] = src[1];
    wcols[item] = cols;
}

// ------------------------------------------------------------------------------------
// Stage 2a.  A job is one (read, direction, adapter) pair, numbered item * n_adapters + adapter
// over the items in sorted order.  block_test() drops the pairs that cannot have a candidate;
// the survivors' job numbers go to `jobs` (warp-aggregated append), which stage 2b consumes.
__global__ void __launch_bounds__(SCAN_THREADS)
filter_kernel(const RoundTable *__restrict__ tab, const uint32_t *__restrict__ W,
              const View *__restrict__ views, const WinList *__restrict__ wins,
              const uint32_t *__restrict__ wcols_sorted, const uint32_t *__restrict__ item_order,
              uint32_t n_items, uint32_t *__restrict__ jobs, uint32_t *__restrict__ counters,
              unsigned long long *__restrict__ cells_2b, int trim_on)
{
    __shared__ __align__(16) uint32_t s_peq32b[16][64];
    __shared__ int8_t s_first_lim[MAX_M + 32];
    __shared__ uint8_t s_kmax[MAX_AD][MAX_M + 8];     // the pruning limits (kmax[a][0])
    __shared__ int s_k[MAX_AD], s_min_ov[MAX_AD], s_lb[MAX_AD], s_m[MAX_AD];
    __shared__ int s_na, s_type, s_trim;      // s_trim: k_max + 1 when the windows come from stage 1 (0: whole reads)
    for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) (&s_peq32b[0][0])[i] = (&tab->peq32b[0][0])[i];
    for (int i = threadIdx.x; i < MAX_M + 32; i += blockDim.x) s_first_lim[i] = tab->first_lim[i];
    for (int i = threadIdx.x; i < MAX_AD * (MAX_M + 8); i += blockDim.x)
        (&s_kmax[0][0])[i] = tab->kmax[i / (MAX_M + 8)][0][i % (MAX_M + 8)];
    if (threadIdx.x < MAX_AD) {
        s_k[threadIdx.x] = tab->k[threadIdx.x];
        s_min_ov[threadIdx.x] = tab->min_ov[threadIdx.x];
        s_lb[threadIdx.x] = tab->block_len[threadIdx.x];
        s_m[threadIdx.x] = tab->m[threadIdx.x];
    }
    if (threadIdx.x == 0) { s_na = tab->n_adapters; s_type = tab->type; s_trim = (tab->use_filter && trim_on) ? tab->k_max + 1 : 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t na = (uint32_t)s_na;
    const uint32_t n_jobs = n_items * na;
    const int type = s_type;
    const char *base_b = reinterpret_cast<const char *>(&s_peq32b[0][0]);
    uint32_t *job_counter = counters + 4, *n_out = counters + 5;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(job_counter, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_jobs) break;
        const uint32_t p = base + (uint32_t)lane;
        bool active = p < n_jobs;
        uint32_t it = 0, item = 0;
        int a = 0;
        if (active) {
            it = p / na;
            a = (int)(p - it * na);
            active = wcols_sorted[it] != 0u;     // sorted descending: inactive items are at the end
            item = item_order[it];
        }
        if (__ballot_sync(0xffffffffu, active) == 0u) break;
        bool keep = false;
        if (active) {
            const int dir = (int)(item & 1u);
            const View v = views[item >> 1];
            WinList wl;
            {
                const uint4 *src = reinterpret_cast<const uint4 *>(wins + item);
                uint4 *dst = reinterpret_cast<uint4 *>(&wl);
                dst[0] = src[0]; dst[1] = src[1];
            }
            keep = block_test(W, v.lo, v.len, dir, &wl, base_b, a + (int)na * dir, s_lb[a], s_k[a], type,
                              s_kmax[a], s_min_ov[a], s_first_lim, s_trim ? s_m[a] - s_lb[a] - s_trim + 1 : 0);
        }
        const uint32_t mk = __ballot_sync(0xffffffffu, keep);
        if (mk) {
            // cells stage 2b will update for these pairs (the roofline's executed-cells figure)
            const uint32_t c2 = __reduce_add_sync(0xffffffffu, keep ? wcols_sorted[it] * (uint32_t)s_m[a] : 0u);
            if (lane == 0) atomicAdd(cells_2b, (unsigned long long)c2);
            uint32_t ob = 0;
            if (lane == 0) ob = atomicAdd(n_out, (uint32_t)__popc(mk));
            ob = __shfl_sync(0xffffffffu, ob, 0);
            if (keep) jobs[ob + (uint32_t)__popc(mk & lanemask_lt())] = p;
        }
    }
}

// ------------------------------------------------------------------------------------
// Stage 2 of the scan.  A job is one (read, direction, adapter) pair; the (read, direction)
// items come sorted by the number of window columns they have to scan, so the 32 lanes of a
// warp -- 32 consecutive jobs, i.e. the adapters of two or three items -- run the same number
// of columns.  Persistent warps pull 32 jobs at a time from a global counter.  A pair whose
// candidates all have cost 0 is finished here; the others go to the resolver's work list.
// (six blocks per SM is what the 34 KB round table in shared memory allows; capping the registers at 80 to get
// there measured 4 % faster than the 94 the compiler takes unasked, a cap of 64 slower)
__global__ void __launch_bounds__(SCAN_THREADS, 6)
scan_kernel(const RoundTable *__restrict__ tab, const uint32_t *__restrict__ W,
            const View *__restrict__ views, const WinList *__restrict__ wins,
            const uint32_t *__restrict__ wcols_sorted, const uint32_t *__restrict__ item_order,
            uint32_t n_items, PairResult *__restrict__ results, Task *__restrict__ work,
            unsigned long long *__restrict__ best_key, uint32_t *__restrict__ counters,
            const uint32_t *__restrict__ jobs, uint32_t cap_pairs)
{
    __shared__ __align__(16) RoundTable T;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(tab);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&T);
        for (int i = threadIdx.x; i < (int)(sizeof(RoundTable) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t na = (uint32_t)T.n_adapters;
    // with a job list (stage 2a ran) the jobs are its entries, else every pair of every item
    const uint32_t n_jobs = jobs ? counters[5] : n_items * na;
    const int type = T.type;
    uint32_t *job_counter = counters + 0, *res_count = counters + 1, *work_count = counters + 2, *wide_count = counters + 3;

    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(job_counter, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_jobs) break;
        uint32_t p = base + (uint32_t)lane;
        bool active = p < n_jobs;
        if (active && jobs) p = jobs[p];
        uint32_t it = 0, item = 0;
        int a = 0;
        if (active) {
            it = p / na;
            a = (int)(p - it * na);
            active = wcols_sorted[it] != 0u;     // sorted descending: inactive items are at the end
            item = item_order[it];
        }
        if (!jobs && __ballot_sync(0xffffffffu, active) == 0u) break;
        bool has = false, need = false;
        LaneScan L;
        uint32_t r = 0;
        int o = 0, m = 0;
        uint32_t n = 0;
        if (active) {
            r = item >> 1;
            const int dir = (int)(item & 1u);
            const View v = views[r];
            o = dir ^ (int)(v.rc & 1u);
            n = v.len;
            const int tl = a + (int)na * dir;
            m = T.m[a];
            WinList wl;
            {
                const uint4 *src = reinterpret_cast<const uint4 *>(wins + item);
                uint4 *dst = reinterpret_cast<uint4 *>(&wl);
                dst[0] = src[0]; dst[1] = src[1];
            }
            scan_lane(W, v.lo, v.len, dir, &wl, peq_bank(T, tl), tl, T.pv0[tl], T.d0[tl], m, T.k[a], T.kmax[a][0],
                      T.min_ov[a], type, L, T.indels, T.code4[a], T.rcode4[a], T.chunk_lut);
            has = L.h.jf <= L.h.jl || L.h.i1 <= L.h.i2;
            need = has && L.need != 0;
        }
        const uint32_t mh = __ballot_sync(0xffffffffu, has);
        if (mh) {
            uint32_t rb = 0;
            if (lane == 0) rb = atomicAdd(res_count, (uint32_t)__popc(mh));
            rb = __shfl_sync(0xffffffffu, rb, 0);
            const uint32_t slot = rb + (uint32_t)__popc(mh & lanemask_lt());
            // tasks for the band resolver go to the front half of `work`, the few it cannot take (end
            // cells spread over too many diagonals, very long scans) to the back half
            const bool wide = need && !task_band_ok(type, m, T.k[a], (int)n,
                                                    Task{r, 0u, L.h.jf, L.h.jl, L.h.i1, L.h.i2, 0u, 0});
            const uint32_t mn = __ballot_sync(0xffffffffu, need && !wide);
            const uint32_t mw = __ballot_sync(0xffffffffu, wide);
            uint32_t wb = 0, xb = 0;
            if (mn) {
                if (lane == 0) wb = atomicAdd(work_count, (uint32_t)__popc(mn));
                wb = __shfl_sync(0xffffffffu, wb, 0);
            }
            if (mw) {
                if (lane == 0) xb = atomicAdd(wide_count, (uint32_t)__popc(mw));
                xb = __shfl_sync(0xffffffffu, xb, 0);
            }
            // A pair beyond the arenas is dropped: the final res_count (>= the task counts) says so to the
            // host, which then runs the batch again with worst-case arenas (orc_api.cu grow_pair_arenas).  The
            // work position and the result slot come from counters that other warps interleave, so each
            // is checked on its own; a task whose result has no slot is marked so that the resolver skips it.
            const uint32_t wpos = wide ? cap_pairs + xb + (uint32_t)__popc(mw & lanemask_lt())
                                       : wb + (uint32_t)__popc(mn & lanemask_lt());
            if (need) {
                if (wpos < (wide ? 2u * cap_pairs : cap_pairs)) {
                    Task t;
                    t.read = r; t.lane = (uint32_t)(a + (int)na * (int)(item & 1u));
                    t.jf = L.h.jf; t.jl = L.h.jl; t.i1 = L.h.i1; t.i2 = L.h.i2;
                    t.slot = slot < cap_pairs ? slot : 0xFFFFFFFFu;
                    t.pad_ = task_anchors(L) | (wide ? (int32_t)TASK_WIDE : 0);
                    work[wpos] = t;
                }
            } else if (has && slot < cap_pairs) {
                PairResult res;
                best_to_result(L.best, m, (int)n, res);
                results[slot] = res;
                if (res.has)
                    atomicMax(best_key + (size_t)r * 2 + o, (unsigned long long)pack_key(res.score, res.errors, a, slot));
            }
        }
    }
}

// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
resolve_kernel(const RoundTable *__restrict__ tab, const uint32_t *__restrict__ W,
               const View *__restrict__ views, const Task *__restrict__ work,
               const uint32_t *__restrict__ work_count, PairResult *__restrict__ results,
               unsigned long long *__restrict__ best_key, uint32_t cap_pairs)
{
    // Only the head of the table (everything up to and including peq) is used here; the rest
    // stays out of shared memory so that L1 keeps more room for the local-memory column rings.
    constexpr int HEAD = (int)offsetof(RoundTable, pv0);
    __shared__ __align__(16) unsigned char s_tab[HEAD];
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(tab);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_tab);
        for (int i = threadIdx.x; i < HEAD / 4; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const RoundTable &T = *reinterpret_cast<const RoundTable *>(s_tab);
    const uint32_t n = min(*work_count, cap_pairs);
    ColRing ring;
    // every lane of a warp makes the same number of trips, so the warp can be brought back together
    // between the column scan (lanes differ in length) and the walks (which then start together)
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t first = blockIdx.x * blockDim.x + (threadIdx.x & ~31u);
    for (uint32_t base = first; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t t = base + lane;
        bool act = t < n;
        Task task;
        View v;
        ResolveCtx C;
        if (act) {
            task = work[t];
            act = task.slot != 0xFFFFFFFFu;          // its result had no slot (arena overflow): skipped
        }
        if (act) {
            v = views[task.read];
            resolve_begin(W, v, T, task, C, ring);
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, act);      // also the meeting point
        if (act) {
            PairResult res;
            res.has = 0; res.ref_start = res.ref_stop = res.query_start = res.query_stop = 0;
            res.score = res.errors = 0; res.pad_ = 0;
            resolve_end(W, v, C, res, ring, mask);
            results[task.slot] = res;
            if (res.has) {
                const int a = (int)task.lane % T.n_adapters;
                const int o = ((int)task.lane / T.n_adapters) ^ (int)(v.rc & 1u);
                atomicMax(best_key + (size_t)task.read * 2 + o,
                          (unsigned long long)pack_key(res.score, res.errors, a, task.slot));
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// The band resolver (orc_core.cuh band_*): one thread per task, its ring of scan checkpoints (8 bytes after
// every BAND_CKPT-th column) in shared memory, entry q of thread t at ring[q * BAND_THREADS + t]: the lanes of
// a warp scan in lockstep and store one contiguous 256-byte row per checkpoint, and a 64-bit access of any 16
// lanes to any rows hits 16 different bank pairs, so the walks read without conflicts as well.  Only the banks of the 64-bit match
// table that the round uses are staged next to the ring; the acceptance limits and the packed adapter
// codes, touched a few times per task, are read through L1 from the table in global memory.
#ifndef ORC_BAND_THREADS
#define ORC_BAND_THREADS 256
#endif
constexpr int BAND_THREADS = ORC_BAND_THREADS;
inline size_t band_smem_bytes(int n_lanes, int n_adapters)
{
    return (size_t)((n_lanes + 31) / 32) * BAND_BANK_BYTES + (size_t)n_adapters * sizeof(BandAdapter) +
           (size_t)BAND_ENTRIES * BAND_THREADS * sizeof(BandEntry) + (size_t)BAND_CODE_WORDS * BAND_THREADS * sizeof(uint32_t);
}
static_assert(sizeof(BandAdapter) % 8 == 0, "the ring behind the adapter table must stay 8-byte aligned");

#ifdef ORC_BAND_MINBLOCKS
__global__ void __launch_bounds__(BAND_THREADS, ORC_BAND_MINBLOCKS)
#else
__global__ void __launch_bounds__(BAND_THREADS)
#endif
resolve_band_kernel(const RoundTable *__restrict__ tab, const uint32_t *__restrict__ W,
                    const View *__restrict__ views, const Task *__restrict__ work,
                    const uint32_t *__restrict__ work_count, PairResult *__restrict__ results,
                    unsigned long long *__restrict__ best_key, uint32_t cap_pairs)
{
    extern __shared__ __align__(16) unsigned char s_band[];
    const int n_banks = (tab->n_lanes + 31) / 32, na = tab->n_adapters, type = tab->type;
    for (int i = threadIdx.x; i < n_banks * 16 * 32; i += blockDim.x) {     // the band's own match table
        uint32_t e[4];
        band_table_entry(*tab, i >> 9, (i >> 5) & 15, i & 31, e);
        reinterpret_cast<uint4 *>(s_band)[i] = make_uint4(e[0], e[1], e[2], e[3]);
    }
    BandAdapter *s_ads = reinterpret_cast<BandAdapter *>(s_band + (size_t)n_banks * BAND_BANK_BYTES);
    for (int a = threadIdx.x; a < na; a += blockDim.x) band_adapter_fill(*tab, a, s_ads[a]);
    __syncthreads();
    BandRing ring;
    ring.p = reinterpret_cast<BandEntry *>(s_ads + na) + threadIdx.x;
    ring.stride = BAND_THREADS;
    ring.cw = reinterpret_cast<uint32_t *>(reinterpret_cast<BandEntry *>(s_ads + na) + BAND_ENTRIES * BAND_THREADS) + threadIdx.x;
    ring.w0 = 0;
    const uint32_t n = min(*work_count, cap_pairs);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t first = blockIdx.x * blockDim.x + (threadIdx.x & ~31u);
    for (uint32_t base = first; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t t = base + lane;
        bool act = t < n;
        Task task;
        View v;
        BandCtx C;
        if (act) {
            task = work[t];
            act = task.slot != 0xFFFFFFFFu;          // its result had no slot (arena overflow): skipped
        }
        if (act) {
            v = views[task.read];
            band_begin(W, v, type, na, s_ads, task, C, ring,
                       reinterpret_cast<const char *>(s_band) + (size_t)(task.lane >> 5) * BAND_BANK_BYTES);
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, act);      // also the meeting point
        if (act) {
            PairResult res;
            res.has = 0; res.ref_start = res.ref_stop = res.query_start = res.query_stop = 0;
            res.score = res.errors = 0; res.pad_ = 0;
            band_end(W, v, C, res, ring, mask);
            results[task.slot] = res;
            if (res.has) {
                const int a = (int)task.lane % na;
                const int o = ((int)task.lane / na) ^ (int)(v.rc & 1u);
                atomicMax(best_key + (size_t)task.read * 2 + o,
                          (unsigned long long)pack_key(res.score, res.errors, a, task.slot));
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// Anchored no-indel round: one thread per (read, orientation); touches m bytes per read.
__global__ void __launch_bounds__(128)
anchored_kernel(const AnchoredTable *__restrict__ tab, const uint8_t *__restrict__ seq,
                const uint8_t *__restrict__ comp_lut_g, const View *__restrict__ views,
                const Match *__restrict__ prev, uint32_t n_reads, PairResult *__restrict__ results,
                unsigned long long *__restrict__ best_key)
{
    __shared__ __align__(16) AnchoredTable T;
    __shared__ uint8_t comp[256];
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(tab);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&T);
        for (int i = threadIdx.x; i < (int)(sizeof(AnchoredTable) / 4); i += blockDim.x) dst[i] = src[i];
        for (int i = threadIdx.x; i < 256; i += blockDim.x) comp[i] = comp_lut_g[i];
    }
    __syncthreads();
    // grid-stride: the table (6.7 KB) is staged once per block, not once per 128 items
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < 2u * n_reads; p += gridDim.x * blockDim.x) {
        const uint32_t r = p >> 1;
        const int o = (int)(p & 1u);
        if (prev != nullptr && prev[r].adapter < 0) continue;
        if (o == 1 && !T.revcomp) continue;
        const View v = views[r];
        PairResult res;
        const int a = anchored_match(seq, comp, v, o, T, res);
        if (a >= 0) {
            results[p] = res;
            best_key[p] = (unsigned long long)pack_key(res.score, res.errors, a, p);
        }
    }
}

// ------------------------------------------------------------------------------------
// Is there a U among the bases of the batch?  One warp per read.  Only launched when plain and IUPAC adapters stand
// side by side (unify_wildcards): such a read is the one case in which the masks and cutadapt's ASCII comparison of
// the plain adapters disagree.
__global__ void __launch_bounds__(256)
u_scan_kernel(const uint8_t *__restrict__ seq, const uint64_t *__restrict__ offsets, const uint32_t *__restrict__ lengths,
              uint32_t n_reads, uint32_t *__restrict__ flag)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    bool seen = false;
    for (uint32_t r = warp; r < n_reads; r += n_warps) {
        const uint8_t *s = seq + offsets[r];
        const uint32_t len = lengths[r];
        for (uint32_t i = lane; i < len; i += 32u) seen = seen || ((s[i] & 0xDFu) == (uint8_t)'U');
    }
    if (__any_sync(0xffffffffu, seen) && lane == 0) atomicOr(flag, 1u);
}

// ------------------------------------------------------------------------------------
// A round with adapters over 64 nt (LongTable): one thread per (read, orientation) runs cutadapt's recurrence
// for every adapter, its column in local memory.
__global__ void __launch_bounds__(128)
long_kernel(const LongTable *__restrict__ tab, const uint32_t *__restrict__ W, const View *__restrict__ views,
            const Match *__restrict__ prev, uint32_t n_reads, PairResult *__restrict__ results,
            unsigned long long *__restrict__ best_key)
{
    __shared__ __align__(16) LongTable T;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(tab);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&T);
        for (int i = threadIdx.x; i < (int)(sizeof(LongTable) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    LongCell col[MAX_M_LONG + 1];
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < 2u * n_reads; p += gridDim.x * blockDim.x) {
        const uint32_t r = p >> 1;
        const int o = (int)(p & 1u);
        if (prev != nullptr && prev[r].adapter < 0) continue;
        if (o == 1 && !T.revcomp) continue;
        const View v = views[r];
        PairResult res;
        const int a = long_match(W, v, o, T, col, res);
        if (a >= 0) {
            results[p] = res;
            best_key[p] = (unsigned long long)pack_key(res.score, res.errors, a, p);
        }
    }
}

// ------------------------------------------------------------------------------------
// select: also derives, after the last round, the bin id and the FASTQ record size.
struct SelectArgs {
    int type, revcomp, action;  // of the round (RoundTable.type / .revcomp / .action)
    const View *views_in;
    View *views_out;
    const Match *prev;          // matches of the previous round (nullptr in round 1)
    Match *out;
    const unsigned long long *best_key;   // [n_reads][2]
    const PairResult *results;
    uint32_t n_reads;
    // last round only:
    int last_round, round_index, n_ad0;
    const Match *match0;        // round-1 matches (== out when n_rounds == 1)
    const uint8_t *drop_bins;
    const uint64_t *name_offsets;   // may be nullptr (no FASTQ emission)
    const uint32_t *name_lengths;   // may be nullptr (then name r ends where name r+1 starts)
    int32_t *bin;
    uint32_t *out_len, *rec_bytes;
    unsigned long long *next_bases;   // sum of the view lengths that enter the next round (or nullptr)
    uint32_t *next_len_hist;          // size classes of those lengths for the next round's ordering (or nullptr)
};

__global__ void __launch_bounds__(128) select_kernel(SelectArgs A)
{
    __shared__ uint32_t s_hist[LEN_CLASSES];
    if (A.next_len_hist != nullptr) {
        for (int i = threadIdx.x; i < LEN_CLASSES; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
    }
    const uint32_t n_pad = (A.n_reads + 31u) & ~31u;        // whole warps make every trip (warp reductions inside)
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_pad; r += gridDim.x * blockDim.x) {
        const bool valid = r < A.n_reads;
        View v;
        v.lo = 0; v.len = 0; v.rc = 0;
        if (valid) v = A.views_in[r];
        Match mt;
        View next = v;
        if (!valid || (A.prev != nullptr && A.prev[r].adapter < 0)) {
            mt.adapter = -1; mt.is_rc = 0;
            mt.ref_start = mt.ref_stop = mt.query_start = mt.query_stop = mt.score = mt.errors = 0;
        } else {
            uint64_t key[2];
            key[0] = A.best_key[(size_t)r * 2];
            key[1] = A.best_key[(size_t)r * 2 + 1];
            select_read(A.type, A.revcomp, v, key, A.results, mt, next, A.action);
        }
        if (A.next_bases != nullptr) {
            const uint32_t add = (valid && mt.adapter >= 0) ? next.len : 0u;
            const uint32_t sum = __reduce_add_sync(0xffffffffu, add);
            if ((threadIdx.x & 31) == 0 && sum) atomicAdd(A.next_bases, (unsigned long long)sum);
        }
        if (!valid) continue;
        if (A.next_len_hist != nullptr)     // what bucket_scatter_kernel<true> will read as this read's key
            atomicAdd(&s_hist[len_bucket(mt.adapter >= 0 ? next.len : 0u)], 1u);
        A.out[r] = mt;
        A.views_out[r] = next;
        if (A.last_round) {
            int b;
            if (A.round_index == 0) b = mt.adapter + 1;
            else b = (A.match0[r].adapter + 1) + (A.n_ad0 + 1) * (mt.adapter + 1);
            if (A.drop_bins[b]) b = -1;
            A.bin[r] = b;
            A.out_len[r] = next.len;
            uint32_t rb = 0;
            if (b >= 0 && A.name_offsets != nullptr) {
                const uint32_t nl = A.name_lengths ? A.name_lengths[r]
                                                   : (uint32_t)(A.name_offsets[r + 1] - A.name_offsets[r]);
                rb = 1u + nl + 3u * (next.rc >> 8) + 1u + next.len + 3u + next.len + 1u;   // @name[ rc]*\nSEQ\n+\nQUAL\n
            }
            A.rec_bytes[r] = rb;
        }
    }
    if (A.next_len_hist != nullptr) {
        __syncthreads();
        for (int i = threadIdx.x; i < LEN_CLASSES; i += blockDim.x)
            if (s_hist[i]) atomicAdd(A.next_len_hist + i, s_hist[i]);
    }
}

// ------------------------------------------------------------------------------------
// Stable multi-way partition.  Chunk c = reads [c*BIN_CHUNK, (c+1)*BIN_CHUNK), one warp.
// hist layout: [bin][chunk] so that the scan over chunks is contiguous.
__global__ void __launch_bounds__(128)
bin_count_kernel(const int32_t *__restrict__ bin, const uint32_t *__restrict__ rec_bytes, uint32_t n_reads,
                 int n_bins, uint32_t n_chunks, uint32_t *__restrict__ hist_cnt, uint64_t *__restrict__ hist_bytes)
{
    __shared__ uint32_t s_cnt[4][MAX_BINS];
    __shared__ unsigned long long s_bytes[4][MAX_BINS];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t chunk = blockIdx.x * 4 + w;
    for (int b = lane; b < n_bins; b += 32) { s_cnt[w][b] = 0; s_bytes[w][b] = 0; }
    __syncwarp();
    if (chunk < n_chunks) {
        const uint32_t r0 = chunk * BIN_CHUNK;
        for (int it = 0; it < BIN_CHUNK / 32; it++) {
            const uint32_t r = r0 + it * 32 + lane;
            if (r < n_reads) {
                const int b = bin[r];
                if (b >= 0) {
                    atomicAdd(&s_cnt[w][b], 1u);
                    atomicAdd(&s_bytes[w][b], (unsigned long long)rec_bytes[r]);
                }
            }
        }
        __syncwarp();
        for (int b = lane; b < n_bins; b += 32) {
            hist_cnt[(size_t)b * n_chunks + chunk] = s_cnt[w][b];
            hist_bytes[(size_t)b * n_chunks + chunk] = s_bytes[w][b];
        }
    }
}

// One block per bin: exclusive scan over the chunks of the bin (in place) and the bin's totals.
__global__ void __launch_bounds__(256)
bin_scan_kernel(uint32_t n_chunks, uint32_t *__restrict__ hist_cnt, uint64_t *__restrict__ hist_bytes,
                uint64_t *__restrict__ bin_counts, uint64_t *__restrict__ bin_bytes)
{
    __shared__ uint32_t s_c[8];
    __shared__ unsigned long long s_b[8];
    __shared__ uint32_t s_run_c;
    __shared__ unsigned long long s_run_b;
    const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t *hc = hist_cnt + (size_t)b * n_chunks;
    uint64_t *hb = hist_bytes + (size_t)b * n_chunks;
    if (threadIdx.x == 0) { s_run_c = 0; s_run_b = 0; }
    __syncthreads();
    for (uint32_t c0 = 0; c0 < n_chunks; c0 += 256) {
        const uint32_t c = c0 + threadIdx.x;
        const uint32_t vc = c < n_chunks ? hc[c] : 0u;
        const unsigned long long vb = c < n_chunks ? hb[c] : 0ull;
        uint32_t ic = vc;
        unsigned long long ib = vb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t tc = __shfl_up_sync(0xffffffffu, ic, d);
            const unsigned long long tb = __shfl_up_sync(0xffffffffu, ib, d);
            if (lane >= d) { ic += tc; ib += tb; }
        }
        if (lane == 31) { s_c[w] = ic; s_b[w] = ib; }
        __syncthreads();
        uint32_t oc = s_run_c;
        unsigned long long ob = s_run_b;
        for (int i = 0; i < w; i++) { oc += s_c[i]; ob += s_b[i]; }
        if (c < n_chunks) { hc[c] = oc + ic - vc; hb[c] = ob + ib - vb; }
        __syncthreads();
        if (threadIdx.x == 255) { s_run_c = oc + ic; s_run_b = ob + ib; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { bin_counts[b] = s_run_c; bin_bytes[b] = s_run_b; }
}

// exclusive scan of the bin byte totals -> where each bin starts in the FASTQ output
__global__ void bin_offsets_kernel(int n_bins, const uint64_t *__restrict__ bin_bytes, uint64_t *__restrict__ bin_offsets)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long acc = 0;
        for (int b = 0; b < n_bins; b++) { bin_offsets[b] = acc; acc += bin_bytes[b]; }
        bin_offsets[n_bins] = acc;
    }
}

__global__ void __launch_bounds__(128)
bin_place_kernel(const int32_t *__restrict__ bin, const uint32_t *__restrict__ rec_bytes, uint32_t n_reads,
                 int n_bins, uint32_t n_chunks, const uint64_t *__restrict__ hist_bytes,
                 const uint64_t *__restrict__ bin_offsets, uint64_t *__restrict__ dest)
{
    __shared__ unsigned long long s_cur[4][MAX_BINS];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t chunk = blockIdx.x * 4 + w;
    if (chunk >= n_chunks) return;
    for (int b = lane; b < n_bins; b += 32)
        s_cur[w][b] = bin_offsets[b] + hist_bytes[(size_t)b * n_chunks + chunk];
    __syncwarp();
    const uint32_t r0 = chunk * BIN_CHUNK;
    for (int it = 0; it < BIN_CHUNK / 32; it++) {
        const uint32_t r = r0 + it * 32 + lane;
        const int b = r < n_reads ? bin[r] : -1;
        const unsigned long long nb = (b >= 0) ? rec_bytes[r] : 0ull;
        unsigned long long d = ~0ull;
        // input order inside a bin: the lanes that hold reads of the same bin (match-any ballot) take their
        // places by a prefix sum of the record sizes of the peers in front of them; the last peer moves the
        // bin's cursor on for the next 32 reads
        const uint32_t have = __ballot_sync(0xffffffffu, b >= 0);
        if (b >= 0) {
            const uint32_t peers = __match_any_sync(have, b);
            unsigned long long before = 0, total = 0;
            for (uint32_t m = peers; m; m &= m - 1u) {
                const int src = __ffs((int)m) - 1;
                const unsigned long long v = __shfl_sync(peers, nb, src);
                if (src < lane) before += v;
                total += v;
            }
            const unsigned long long cur = s_cur[w][b];
            d = cur + before;
            __syncwarp(peers);
            if (lane == 31 - __clz((int)peers)) s_cur[w][b] = cur + total;
        }
        __syncwarp();
        if (r < n_reads) dest[r] = d;
    }
}

// Warp-cooperative copy of n bytes with arbitrary source and destination alignment: the
// destination is written in aligned 32-bit words assembled from two aligned source words by a
// funnel shift (head and tail bytes singly).  The loads of two rounds (256 bytes per warp) are
// issued before the first store, so that a warp keeps several requests in flight.  src may be
// over-read by up to 7 bytes (the blobs have slack).
__device__ __forceinline__ void warp_copy(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src,
                                          uint32_t n, int lane)
{
    const uint32_t head = min(n, (uint32_t)((4u - ((uintptr_t)dst & 3u)) & 3u));
    if ((uint32_t)lane < head) dst[lane] = src[lane];
    const uint32_t nw = (n - head) >> 2;
    const uint8_t *sp = src + head;
    const uint32_t sh = (uint32_t)((uintptr_t)sp & 3u) * 8u;
    const uint32_t *sa = reinterpret_cast<const uint32_t *>((uintptr_t)sp & ~(uintptr_t)3);
    uint32_t *da = reinterpret_cast<uint32_t *>(dst + head);
    for (uint32_t base = lane; base < nw; base += 64) {
        const uint32_t w1 = base + 32u;
        const bool two = w1 < nw;
        const uint32_t a0 = sa[base], a1 = sa[base + 1];
        uint32_t b0 = 0, b1 = 0;
        if (two) { b0 = sa[w1]; b1 = sa[w1 + 1]; }
        da[base] = __funnelshift_r(a0, a1, sh);
        if (two) da[w1] = __funnelshift_r(b0, b1, sh);
    }
    const uint32_t done = head + 4u * nw;
    if (done + (uint32_t)lane < n) dst[done + lane] = src[done + lane];
}

// The same with 16-byte stores: the destination is written in aligned 16-byte words, each assembled from the
// two aligned source words it straddles by four funnel shifts; which four of the eight loaded 32-bit words
// are involved depends on the copy's source/destination phase only, so the choice is the same for every
// lane.  Head and tail bytes singly.  src may be over-read by up to 31 bytes (the blobs have 64 bytes of slack).
__device__ __forceinline__ uint4 shift16(const uint4 lo, const uint4 hi, uint32_t a, uint32_t b)
{
    uint4 r;
    switch (a) {                    // a = source phase / 4 (warp-uniform), b = (source phase % 4) * 8
    case 0: r.x = __funnelshift_r(lo.x, lo.y, b); r.y = __funnelshift_r(lo.y, lo.z, b);
            r.z = __funnelshift_r(lo.z, lo.w, b); r.w = __funnelshift_r(lo.w, hi.x, b); break;
    case 1: r.x = __funnelshift_r(lo.y, lo.z, b); r.y = __funnelshift_r(lo.z, lo.w, b);
            r.z = __funnelshift_r(lo.w, hi.x, b); r.w = __funnelshift_r(hi.x, hi.y, b); break;
    case 2: r.x = __funnelshift_r(lo.z, lo.w, b); r.y = __funnelshift_r(lo.w, hi.x, b);
            r.z = __funnelshift_r(hi.x, hi.y, b); r.w = __funnelshift_r(hi.y, hi.z, b); break;
    default: r.x = __funnelshift_r(lo.w, hi.x, b); r.y = __funnelshift_r(hi.x, hi.y, b);
            r.z = __funnelshift_r(hi.y, hi.z, b); r.w = __funnelshift_r(hi.z, hi.w, b); break;
    }
    return r;
}

__device__ __forceinline__ void warp_copy16(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src,
                                            uint32_t n, int lane)
{
    const uint32_t head = min(n, (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u));
    if ((uint32_t)lane < head) dst[lane] = src[lane];
    const uint32_t nq = (n - head) >> 4;
    const uint8_t *sp = src + head;
    const uint32_t ph = (uint32_t)((uintptr_t)sp & 15u);
    const uint32_t a = ph >> 2, b = (ph & 3u) * 8u;
    const uint4 *sa = reinterpret_cast<const uint4 *>((uintptr_t)sp & ~(uintptr_t)15);
    uint4 *da = reinterpret_cast<uint4 *>(dst + head);
    for (uint32_t q = lane; q < nq; q += 64) {
        const uint32_t q1 = q + 32u;
        const bool two = q1 < nq;
        const uint4 l0 = sa[q], h0 = sa[q + 1];
        uint4 l1 = make_uint4(0, 0, 0, 0), h1 = l1;
        if (two) { l1 = sa[q1]; h1 = sa[q1 + 1]; }
        da[q] = shift16(l0, h0, a, b);
        if (two) da[q1] = shift16(l1, h1, a, b);
    }
    const uint32_t done = head + 16u * nq;
    if (done + (uint32_t)lane < n) dst[done + lane] = src[done + lane];
}

// The same in reverse: dst[i] = f(src[n - 1 - i]) with f = the complement table (bases) or the
// identity (qualities, comp == nullptr).  Output word w holds the source bytes e, e-1, e-2, e-3
// (e = n - 1 - head - 4w): an unaligned 4-byte window read as above, bytes swapped.
__device__ __forceinline__ void warp_copy_rev(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src,
                                              uint32_t n, int lane, const uint8_t *comp)
{
    const uint32_t head = min(n, (uint32_t)((4u - ((uintptr_t)dst & 3u)) & 3u));
    if ((uint32_t)lane < head) {
        const uint8_t b = src[n - 1 - lane];
        dst[lane] = comp ? comp[b] : b;
    }
    const uint32_t nw = (n - head) >> 2;
    uint32_t *da = reinterpret_cast<uint32_t *>(dst + head);
    const uint8_t *top = src + (n - head);           // one past the source byte of output byte `head`
    auto put = [&](uint32_t w, uint32_t x) {            // x = bytes src[e-3] .. src[e] of output word w
        uint32_t y;
        if (comp) y = (uint32_t)comp[x >> 24] | ((uint32_t)comp[(x >> 16) & 255u] << 8) |
                      ((uint32_t)comp[(x >> 8) & 255u] << 16) | ((uint32_t)comp[x & 255u] << 24);
        else y = __byte_perm(x, 0u, 0x0123u);
        da[w] = y;
    };
    // all output words read the source at the same byte phase: (top - 4(w+1)) & 3 == top & 3
    const uint32_t sh = (uint32_t)((uintptr_t)top & 3u) * 8u;
    const uint32_t *ta = reinterpret_cast<const uint32_t *>((uintptr_t)top & ~(uintptr_t)3);
    for (uint32_t base = lane; base < nw; base += 64) {
        const uint32_t w1 = base + 32u;
        const bool two = w1 < nw;
        const uint32_t *pa = ta - (base + 1u);
        const uint32_t a0 = pa[0], a1 = pa[1];
        uint32_t b0 = 0, b1 = 0;
        if (two) { b0 = pa[-32]; b1 = pa[-31]; }
        put(base, __funnelshift_r(a0, a1, sh));
        if (two) put(w1, __funnelshift_r(b0, b1, sh));
    }
    const uint32_t done = head + 4u * nw;
    if (done + (uint32_t)lane < n) {
        const uint8_t b = src[n - 1 - (done + lane)];
        dst[done + lane] = comp ? comp[b] : b;
    }
}

// '@' name [' rc']* '\n' seq '\n' '+' '\n' qual '\n' of every read that has a destination.
// A warp takes 32 consecutive reads at a time: every lane loads the placement and the view of one of them
// (coalesced, one round of latency for 32 reads, and reads without a destination -- dropped bins -- cost
// nothing more), then the warp copies the records one after the other, the fields handed round by shuffles.
__global__ void __launch_bounds__(256, 8)
emit_kernel(const uint8_t *__restrict__ seq, const uint8_t *__restrict__ qual,
            const uint8_t *__restrict__ names, const uint64_t *__restrict__ name_offsets,
            const uint32_t *__restrict__ name_lengths, const uint64_t *__restrict__ offsets,
            const uint64_t *__restrict__ qual_offsets, const View *__restrict__ views,
            const uint64_t *__restrict__ dest, uint32_t n_reads,
            const uint8_t *__restrict__ comp_lut_g, uint8_t *__restrict__ out)
{
    __shared__ uint8_t comp[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) comp[i] = comp_lut_g[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t base = warp * 32u; base < n_reads; base += n_warps * 32u) {
        const uint32_t r = base + (uint32_t)lane;
        unsigned long long my_d = ~0ull, my_lo = 0, my_n0 = 0, my_qd = 0;
        uint32_t my_len = 0, my_rc = 0, my_nl = 0;
        if (r < n_reads) {
            my_d = dest[r];
            if (my_d != ~0ull) {
                const View v = views[r];
                my_lo = v.lo; my_len = v.len; my_rc = v.rc;
                my_n0 = name_offsets[r];
                my_nl = name_lengths ? name_lengths[r] : (uint32_t)(name_offsets[r + 1] - my_n0);
                my_qd = qual_offsets ? qual_offsets[r] - offsets[r] : 0ull;
            }
        }
        uint32_t todo = __ballot_sync(0xffffffffu, my_d != ~0ull);
        while (todo) {
            const int src_lane = __ffs((int)todo) - 1;
            todo &= todo - 1u;
            const unsigned long long d = __shfl_sync(0xffffffffu, my_d, src_lane);
            const unsigned long long lo = __shfl_sync(0xffffffffu, my_lo, src_lane);
            const unsigned long long n0 = __shfl_sync(0xffffffffu, my_n0, src_lane);
            const unsigned long long qd = __shfl_sync(0xffffffffu, my_qd, src_lane);
            const uint32_t L = __shfl_sync(0xffffffffu, my_len, src_lane);
            const uint32_t rc = __shfl_sync(0xffffffffu, my_rc, src_lane);
            const uint32_t nl = __shfl_sync(0xffffffffu, my_nl, src_lane);
            const uint32_t nrc = rc >> 8;
            uint8_t *o = out + d;
            if (lane == 0) o[0] = '@';
            warp_copy(o + 1, names + n0, nl, lane);
            o += 1 + nl;
            for (uint32_t i = lane; i < 3 * nrc; i += 32) o[i] = (i % 3 == 0) ? ' ' : (i % 3 == 1) ? 'r' : 'c';
            o += 3 * nrc;
            if (lane == 0) o[0] = '\n';
            o += 1;
            const uint8_t *s = seq + lo;
            const uint8_t *q = qual + lo + qd;
            if (rc & 1u) {
                warp_copy_rev(o, s, L, lane, comp);
                warp_copy_rev(o + L + 3, q, L, lane, nullptr);
            } else {
                warp_copy16(o, s, L, lane);
                warp_copy16(o + L + 3, q, L, lane);
                if (nrc >= 2u) {
                    // reverse-complemented in both rounds: dnaio's table sends U to A and A to T, so a read
                    // that is back in its own orientation has T where it had U
                    __syncwarp();
                    for (uint32_t i = lane; i < L; i += 32) {
                        const uint8_t c = s[i];
                        if (c == 'U' || c == 'u') o[i] = (uint8_t)(c - 1);
                    }
                }
            }
            if (lane == 0) { o[L] = '\n'; o[L + 1] = '+'; o[L + 2] = '\n'; o[2 * L + 3] = '\n'; }
        }
    }
}

// ------------------------------------------------------------------------------------
// INT32 issue-rate micro-benchmark (roofline denominator of the scan kernel): 8 independent
// chains per thread.  MODE 0: LOP3 only (ALU pipe).  MODE 1: 4 LOP3 + 4 IMAD (ALU + FMA pipes).
template <int MODE>
__global__ void __launch_bounds__(256) int32_peak_kernel(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t x0 = seed + threadIdx.x, x1 = x0 * 3u, x2 = x0 * 5u, x3 = x0 * 7u;
    uint32_t x4 = x0 * 11u, x5 = x0 * 13u, x6 = x0 * 17u, x7 = x0 * 19u;
    const uint32_t y = seed ^ 0x9e3779b9u, z = seed * 0x85ebca6bu + 1u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0) {
                asm volatile(
                    "lop3.b32 %0, %0, %8, %9, 0x96;\n\t"
                    "lop3.b32 %1, %1, %8, %9, 0x96;\n\t"
                    "lop3.b32 %2, %2, %8, %9, 0x96;\n\t"
                    "lop3.b32 %3, %3, %8, %9, 0x96;\n\t"
                    "lop3.b32 %4, %4, %8, %9, 0x96;\n\t"
                    "lop3.b32 %5, %5, %8, %9, 0x96;\n\t"
                    "lop3.b32 %6, %6, %8, %9, 0x96;\n\t"
                    "lop3.b32 %7, %7, %8, %9, 0x96;\n\t"
                    : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7)
                    : "r"(y), "r"(z));
            } else {
                asm volatile(
                    "lop3.b32 %0, %0, %8, %9, 0x96;\n\t"
                    "mad.lo.u32 %4, %4, %8, %9;\n\t"
                    "lop3.b32 %1, %1, %8, %9, 0x96;\n\t"
                    "mad.lo.u32 %5, %5, %8, %9;\n\t"
                    "lop3.b32 %2, %2, %8, %9, 0x96;\n\t"
                    "mad.lo.u32 %6, %6, %8, %9;\n\t"
                    "lop3.b32 %3, %3, %8, %9, 0x96;\n\t"
                    "mad.lo.u32 %7, %7, %8, %9;\n\t"
                    : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7)
                    : "r"(y), "r"(z));
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}

}  // namespace orc
