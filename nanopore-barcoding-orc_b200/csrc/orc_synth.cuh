// orc_synth.cuh -- synthetic ONT-like reads made on the GPU, straight into a slot's device buffers.
//
// The reference ships no reads, and BASELINE configs[4] asks for 100 M distinct ones: at
// about 1.3 KB of sequence, qualities and name per read that is 130 GB of input, far more than host-side
// generators can make (orcdemux/synth.py: about a minute of CPU per million reads) or PCIe should carry for a
// device-resident measurement.  This is the same read model as a kernel:
//
//   template = SP5[x] + insert + SP27rc[y]     x, y uniform over the round-1 / round-2 adapters of the ctx;
//                                              total length uniform in [len_min, len_max]
//   iid errors over the whole template         substitution 1.5 %, insertion 1.0 %, deletion 1.5 %
//   p = 0.15 each                              5' / 3' truncation by 1..40 nt
//   p = 0.05 each                              no 5' / no 3' adapter
//   p = 0.10                                   the read is reverse-complemented
//   p = 0.005                                  one base is N
//   qualities uniform Phred 5..40; names r<index>, every 7th with a comment
//
// Every random draw is a hash of (seed, read, purpose, position), so the two passes (lengths, then
// bytes) see the same reads and a shard depends on nothing but its seed.  It is NOT the numpy generator's
// stream: parity tests that use these reads export them (orc_export) and hand the same bytes to the oracle.
// Benchmark / test input only; nothing on the matching path depends on it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "orc_core.cuh"

namespace orc {

struct SynthTable {
    int32_t n5, n27;                // adapters of the two rounds
    int32_t m5[MAX_AD], m27[MAX_AD];
    uint8_t a5[MAX_AD][MAX_M];      // 0..3 = A C G T
    uint8_t a27[MAX_AD][MAX_M];
};

struct SynthArgs {
    uint64_t seed;
    uint32_t n_reads, len_min, len_max;
};

__device__ __forceinline__ uint64_t synth_mix(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// one 64-bit draw for (read, purpose, index)
__device__ __forceinline__ uint64_t synth_draw(uint64_t key, uint32_t purpose, uint32_t i)
{
    return synth_mix(key ^ ((uint64_t)purpose << 40) ^ (uint64_t)i * 0xD1342543DE82EF95ull);
}
__device__ __forceinline__ float synth_unit(uint64_t h) { return (float)(h >> 40) * (1.0f / 16777216.0f); }

struct SynthRead {
    uint64_t key;
    uint32_t total;                 // template length
    int32_t x, y;                   // adapter of each end, -1: none
    uint32_t t5, t3;                // truncation asked for (before clamping to half the read)
    uint32_t rc, has_n;
};

__device__ __forceinline__ SynthRead synth_read(const SynthTable &T, const SynthArgs &A, uint32_t r)
{
    SynthRead R;
    R.key = synth_mix(A.seed * 0x2545F4914F6CDD1Dull + r);
    R.total = A.len_min + (uint32_t)(synth_draw(R.key, 0, 0) % (uint64_t)(A.len_max - A.len_min + 1u));
    R.x = (int32_t)(synth_draw(R.key, 0, 1) % (uint64_t)T.n5);
    R.y = (int32_t)(synth_draw(R.key, 0, 2) % (uint64_t)T.n27);
    if (synth_unit(synth_draw(R.key, 0, 3)) < 0.05f) R.x = -1;
    if (synth_unit(synth_draw(R.key, 0, 4)) < 0.05f) R.y = -1;
    R.t5 = synth_unit(synth_draw(R.key, 0, 5)) < 0.15f ? 1u + (uint32_t)(synth_draw(R.key, 0, 6) % 40ull) : 0u;
    R.t3 = synth_unit(synth_draw(R.key, 0, 7)) < 0.15f ? 1u + (uint32_t)(synth_draw(R.key, 0, 8) % 40ull) : 0u;
    R.rc = synth_unit(synth_draw(R.key, 0, 9)) < 0.10f ? 1u : 0u;
    R.has_n = synth_unit(synth_draw(R.key, 0, 10)) < 0.005f ? 1u : 0u;
    return R;
}

// template position p -> what it emits: n in {0, 1, 2} bases b0 (b1)
__device__ __forceinline__ uint32_t synth_emit(const SynthTable &T, const SynthRead &R, uint32_t p, uint32_t &b0, uint32_t &b1)
{
    uint32_t base;
    const uint32_t L5 = R.x >= 0 ? (uint32_t)T.m5[R.x] : 0u, L27 = R.y >= 0 ? (uint32_t)T.m27[R.y] : 0u;
    if (L27 && p + L27 >= R.total && R.total >= L27) base = T.a27[R.y][p - (R.total - L27)];
    else if (p < L5) base = T.a5[R.x][p];
    else base = (uint32_t)(synth_draw(R.key, 1, p) & 3ull);
    const uint64_t h = synth_draw(R.key, 2, p);
    const float u = synth_unit(h);
    if (u < 0.015f) { b0 = (base + 1u + (uint32_t)((h & 0xFFFFull) % 3ull)) & 3u; return 1u; }      // substitution
    if (u < 0.025f) { b0 = base; b1 = (uint32_t)(h & 3ull); return 2u; }                            // insertion behind it
    if (u < 0.040f) return 0u;                                                                      // deletion
    b0 = base;
    return 1u;
}

__device__ __forceinline__ uint32_t synth_name_len(uint32_t g)
{
    uint32_t d = 1, v = g;
    while (v >= 10u) { v /= 10u; d++; }
    uint32_t n = 1u + d;                                    // "r" + digits
    if (g % 7u == 0u) {
        uint32_t c = g % 512u, dc = 1;
        while (c >= 10u) { c /= 10u; dc++; }
        n += 4u + dc;                                       // " ch=" + digits
    }
    return n;
}

// pass 1: one warp per read -> length and name length
__global__ void __launch_bounds__(256)
synth_lengths_kernel(const SynthTable *__restrict__ tab, SynthArgs A, uint32_t *__restrict__ lengths,
                     uint32_t *__restrict__ raw_lengths, uint32_t *__restrict__ name_lengths)
{
    __shared__ SynthTable T;
    for (int i = threadIdx.x; i < (int)(sizeof(SynthTable) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t *>(&T)[i] = reinterpret_cast<const uint32_t *>(tab)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= A.n_reads) return;
    const SynthRead R = synth_read(T, A, r);
    uint32_t cnt = 0;
    for (uint32_t p = lane; p < R.total; p += 32) { uint32_t b0, b1; cnt += synth_emit(T, R, p, b0, b1); }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) {
        const uint32_t t5 = min(R.t5, cnt / 2u), t3 = min(R.t3, cnt / 2u);
        raw_lengths[r] = cnt;
        lengths[r] = cnt - t5 - t3;
        name_lengths[r] = synth_name_len(r);
    }
}

// pass 2: one warp per read -> bases, qualities, name
__global__ void __launch_bounds__(256)
synth_write_kernel(const SynthTable *__restrict__ tab, SynthArgs A, const uint32_t *__restrict__ lengths,
                   const uint32_t *__restrict__ raw_lengths, const uint64_t *__restrict__ offsets,
                   const uint64_t *__restrict__ name_offsets, uint8_t *__restrict__ seq, uint8_t *__restrict__ qual,
                   uint8_t *__restrict__ names)
{
    __shared__ SynthTable T;
    for (int i = threadIdx.x; i < (int)(sizeof(SynthTable) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t *>(&T)[i] = reinterpret_cast<const uint32_t *>(tab)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= A.n_reads) return;
    const SynthRead R = synth_read(T, A, r);
    const uint32_t raw = raw_lengths[r], len = lengths[r];
    const uint32_t t5 = min(R.t5, raw / 2u);
    uint8_t *s = seq + offsets[r], *q = qual + offsets[r];
    const uint32_t n_pos = len ? (uint32_t)(synth_draw(R.key, 4, 0) % (uint64_t)len) : 0u;
    uint32_t done = 0;                                      // bases emitted by earlier template positions
    for (uint32_t p0 = 0; p0 < R.total; p0 += 32) {
        const uint32_t p = p0 + (uint32_t)lane;
        uint32_t b0 = 0, b1 = 0, c = 0;
        if (p < R.total) c = synth_emit(T, R, p, b0, b1);
        uint32_t inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        const uint32_t first = done + inc - c;
        for (uint32_t e = 0; e < c; e++) {
            const uint32_t o = first + e;                   // index in the untruncated read
            if (o < t5 || o - t5 >= len) continue;
            uint32_t pos = o - t5, b = e ? b1 : b0;
            if (R.rc) { pos = len - 1u - pos; b = 3u - b; }
            uint8_t ch = (uint8_t)("ACGT"[b]);
            if (R.has_n && pos == n_pos) ch = 'N';
            s[pos] = ch;
            q[pos] = (uint8_t)(33u + 5u + (uint32_t)(synth_draw(R.key, 3, pos) % 36ull));
        }
        done += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) {                                        // "r<index>" (+ " ch=<index % 512>")
        uint8_t *nm = names + name_offsets[r];
        char buf[24];
        int k = 0;
        uint32_t v = r;
        do { buf[k++] = (char)('0' + v % 10u); v /= 10u; } while (v);
        *nm++ = 'r';
        while (k) *nm++ = (uint8_t)buf[--k];
        if (r % 7u == 0u) {
            *nm++ = ' '; *nm++ = 'c'; *nm++ = 'h'; *nm++ = '=';
            v = r % 512u;
            do { buf[k++] = (char)('0' + v % 10u); v /= 10u; } while (v);
            while (k) *nm++ = (uint8_t)buf[--k];
        }
    }
}

// exclusive prefix sums of the read lengths (-> offsets) and the name lengths (-> name offsets, n + 1
// entries): one block walks the array in tiles; the generator is not a hot path
__global__ void __launch_bounds__(1024)
synth_offsets_kernel(const uint32_t *__restrict__ lengths, const uint32_t *__restrict__ name_lengths, uint32_t n,
                     uint64_t *__restrict__ offsets, uint64_t *__restrict__ name_offsets, uint64_t *__restrict__ totals)
{
    __shared__ unsigned long long s_w[2][32];
    __shared__ unsigned long long s_run[2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_run[0] = 0; s_run[1] = 0; }
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long a = i < n ? lengths[i] : 0ull, b = i < n ? name_lengths[i] : 0ull;
        unsigned long long ia = a, ib = b;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long ta = __shfl_up_sync(0xffffffffu, ia, d), tb = __shfl_up_sync(0xffffffffu, ib, d);
            if (lane >= d) { ia += ta; ib += tb; }
        }
        if (lane == 31) { s_w[0][w] = ia; s_w[1][w] = ib; }
        __syncthreads();
        unsigned long long oa = s_run[0], ob = s_run[1];
        for (int k = 0; k < w; k++) { oa += s_w[0][k]; ob += s_w[1][k]; }
        if (i < n) { offsets[i] = oa + ia - a; name_offsets[i] = ob + ib - b; }
        __syncthreads();
        if (threadIdx.x == 1023) { s_run[0] = oa + ia; s_run[1] = ob + ib; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { name_offsets[n] = s_run[1]; totals[0] = s_run[0]; totals[1] = s_run[1]; }
}

}  // namespace orc
