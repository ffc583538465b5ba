// orc_gz.cuh -- gzip members made on the device (orc_params.emit_gzip).
//
// What the reference leaves on disk are .fastq.gz files (02_cutadapt_loop.sh:64-72, :94-102: `-o ...fastq.gz`);
// the bin-major FASTQ text of a batch becomes one gzip member per bin here, so that only the compressed bytes
// cross PCIe and the host appends them to the bin files as they are (a concatenation of gzip members is a
// gzip file; orc_io.cpp writes the same member format, with the "OC" size field its reader uses to inflate
// members in parallel).
//
// A member holds ONE dynamic-Huffman DEFLATE block of literals only (no LZ77 matches: bases and qualities of
// nanopore reads hardly repeat, zlib's own matches gain about a tenth on them): the code is built per batch
// from a sampled byte histogram of the batch's FASTQ text, so every member of the batch carries the same block
// header.  Encoding is then a table look-up per byte; where a byte's bits go is a prefix sum of code lengths,
// taken per chunk of GZ_CHUNK bytes (one thread), per tile of GZ_TILE chunks (one block), per batch.  CRC-32 of a
// member = the xor of its chunks' CRCs, each moved past the bytes behind it with the "append n zero bytes"
// operator (zlib's crc32_combine; its matrices for n = 2^k come from the host).
//
// Everything that computes is ORC_HD so that tests/gzsim.cpp can run the same code on the CPU and hand the
// members to zlib (tests/test_gz.py); the kernels at the end only distribute the work.
#pragma once
#include <stdint.h>

#include "orc_core.cuh"

namespace orc {

#ifndef GZ_CHUNK_BYTES
#define GZ_CHUNK_BYTES 1024
#endif
constexpr int GZ_CHUNK = GZ_CHUNK_BYTES;    // bytes of text per encoding thread (chunks are GZ_CHUNK-aligned in the text)
constexpr int GZ_TILE = 256;                // chunks per block-wide prefix sum (gz_measure_kernel)
constexpr int GZ_HEADER_BYTES = 20;         // 10 + XLEN + "OC" subfield with the member's size (orc_io.cpp GZ_SIZE_AT)
constexpr int GZ_TRAILER_BYTES = 8;         // CRC-32, ISIZE
constexpr int GZ_HDR_WORDS = 44;            // block header: 3 + 5 + 5 + 4 + 19 * 3 + 258 * 5 = 1364 bits at most
constexpr int GZ_EOB = 256;
constexpr int MAX_BINS_GZ = 512;            // == MAX_BINS (orc_kernels.cuh)

struct GzTable {
    uint16_t code[257];     // Huffman code of byte b (256: end of block), bit-reversed: DEFLATE packs codes MSB first
    uint8_t len[257];       // its length in bits, 1..15 (every byte value has a code: gz_build_from_sorted)
    uint8_t pad_;
    uint32_t hdr_nbits;
    uint32_t hdr[GZ_HDR_WORDS];     // BFINAL BTYPE HLIT HDIST HCLEN, the code-length code, the 258 code lengths
    uint32_t sym[257];              // code[b] | len[b] << 16: one look-up per byte in the kernels
    uint32_t crc_tab[16][256];      // CRC-32 (0xEDB88320): [0] the byte table, [k] = [0] moved past k zero bytes
                                    // (slicing-by-16: of the sixteen look-ups per 16 bytes only the four of the first
                                    // word depend on the running CRC, the other twelve are indexed by text bytes --
                                    // few distinct values in FASTQ text, so few shared-memory bank conflicts)
    uint32_t crc_pow[32][32];       // crc_pow[k]: the operator "append 2^k zero bytes" on a CRC, as a 32 x 32 bit matrix
};

// ---- tables that do not depend on the data (host, once per ctx)
inline uint32_t gz_gf2_times(const uint32_t *mat, uint32_t vec)
{
    uint32_t sum = 0;
    for (int i = 0; vec; vec >>= 1, i++)
        if (vec & 1u) sum ^= mat[i];
    return sum;
}
inline void gz_fill_crc_tables(GzTable &T)
{
    for (uint32_t n = 0; n < 256; n++) {
        uint32_t c = n;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        T.crc_tab[0][n] = c;
    }
    for (uint32_t n = 0; n < 256; n++)
        for (int k = 1; k < 16; k++) T.crc_tab[k][n] = T.crc_tab[0][T.crc_tab[k - 1][n] & 255u] ^ (T.crc_tab[k - 1][n] >> 8);
    // the operator for one zero BIT, squared three times = one zero byte, squared on = 2^k bytes
    uint32_t odd[32], even[32];
    odd[0] = 0xEDB88320u;
    for (int n = 1; n < 32; n++) odd[n] = 1u << (n - 1);
    auto square = [](uint32_t *sq, const uint32_t *m) { for (int n = 0; n < 32; n++) sq[n] = gz_gf2_times(m, m[n]); };
    square(even, odd);      // 2 bits
    square(odd, even);      // 4 bits
    square(even, odd);      // 8 bits = 1 byte
    for (int n = 0; n < 32; n++) T.crc_pow[0][n] = even[n];
    for (int k = 1; k < 32; k++) square(T.crc_pow[k], T.crc_pow[k - 1]);
}

// CRC of A || B from crc(A), crc(B) and |B|: crc(A) moved past |B| zero bytes, xor crc(B)
ORC_HD uint32_t gz_crc_shift(const uint32_t (*pow)[32], uint32_t crc, uint64_t nbytes)
{
    for (int k = 0; nbytes && k < 32; k++, nbytes >>= 1) {
        if (!(nbytes & 1u)) continue;
        uint32_t sum = 0;
        const uint32_t *m = pow[k];
        for (int i = 0; crc; crc >>= 1, i++)
            if (crc & 1u) sum ^= m[i];
        crc = sum;
    }
    return crc;
}

ORC_HD uint32_t gz_rev(uint32_t code, int len)
{
    uint32_t r = 0;
    for (int i = 0; i < len; i++) { r = (r << 1) | (code & 1u); code >>= 1; }
    return r;
}

ORC_HD void gz_put(uint32_t *w, uint32_t &pos, uint32_t val, int nbits)
{
    const uint32_t i = pos >> 5, sh = pos & 31u;
    w[i] |= val << sh;
    if (sh + (uint32_t)nbits > 32u) w[i + 1] |= val >> (32u - sh);
    pos += (uint32_t)nbits;
}

// The batch's code from a byte histogram.  The histogram is a SAMPLE of the text (gz_hist_kernel), so every byte
// value gets a code whether it was seen or not: weight = count + 1.  The unseen ones end up with 14- or 15-bit
// codes and take about 0.01 of the code space; any table made this way codes any text, and the code is complete
// (inflate rejects codes that are not).
struct GzWork {             // scratch of the table build (shared memory on the device)
    unsigned long long w[257];          // weights by symbol
    unsigned long long nw[2 * 257];     // weights by node: the leaves in ascending order, then the inner nodes as made
    int order[257];                     // symbols by ascending (weight, symbol)
    int parent[2 * 257], node_depth[2 * 257], depth[257];
    int num[16], next_code[16], first_rank[16];     // codes per length, first code of a length, see gz_build_from_sorted
    uint32_t hdr[GZ_HDR_WORDS];
};
// weight of symbol s: sampled count + 1 (end of block: 1)
ORC_HD unsigned long long gz_weight(const unsigned long long *hist, int s) { return s < 256 ? hist[s] + 1ull : 1ull; }
// where symbol s stands among the 257 by ascending (weight, symbol): one thread per symbol on the device
ORC_HD int gz_rank(const unsigned long long *w, int s)
{
    int r = 0;
    const unsigned long long ws = w[s];
    for (int t = 0; t < 257; t++) r += (w[t] < ws || (w[t] == ws && t < s)) ? 1 : 0;
    return r;
}
// or a 32-bit word (words of the output at the borders of a chunk are shared with its neighbours; so are the words
// of the block header while the table is built)
ORC_HD void gz_or(uint32_t *p, uint32_t v)
{
#if defined(__CUDA_ARCH__)
    if (v) atomicOr(p, v);
#else
    *p |= v;
#endif
}
ORC_HD void gz_or_bits(uint32_t *w, uint32_t pos, uint32_t val, int nbits)
{
    const uint32_t i = pos >> 5, sh = pos & 31u;
    gz_or(w + i, val << sh);
    if (sh + (uint32_t)nbits > 32u) gz_or(w + i + 1, val >> (32u - sh));
}
ORC_HD void gz_count(int *p)
{
#if defined(__CUDA_ARCH__)
    atomicAdd(p, 1);
#else
    ++*p;
#endif
}

// K.w and K.order filled.  Written for `nt` threads (tid = 0 .. nt-1) that meet at sync(): a block on the device,
// one thread with a sync that does nothing on the host.  Huffman over the sorted leaves with two queues (the
// inner nodes come out in ascending weight, so the two lightest nodes are always at the heads; this part is
// serial), depths top-down.  Then only the NUMBER of codes of every length is kept: lengths above 15 count as 15,
// and while the code is over-subscribed one 15-bit code is taken away and a shorter code is split into two longer
// ones (which pays for one 15-bit code, as in zlib / miniz); the lengths go back to the symbols longest first in
// ascending weight.  Canonical codes (RFC 1951, 3.2.2) and the block header: the code-length code is a fixed
// complete code -- the lengths 0..12 take four bits (codes 0..12), the lengths 13..15 and the three repeat
// symbols five bits (codes 26..31) -- and the lengths are written one by one, without the repeat symbols (about
// 150 bytes per member).
constexpr uint32_t GZ_HDR_FIXED_BITS = 3 + 5 + 5 + 4 + 19 * 3;
struct GzNoSync { ORC_HD void operator()() const {} };
#if defined(__CUDACC__)
struct GzBlockSync { __device__ __forceinline__ void operator()() const { __syncthreads(); } };
#endif
template <typename Sync>
ORC_HD void gz_build_from_sorted(GzTable &T, GzWork &K, int tid, int nt, Sync sync)
{
    const int n = 257;
    int *num = K.num, *next_code = K.next_code, *depth = K.depth;
    for (int i = tid; i < n; i += nt) K.nw[i] = K.w[K.order[i]];
    for (int i = tid; i < GZ_HDR_WORDS; i += nt) K.hdr[i] = 0;
    for (int l = tid; l < 16; l += nt) num[l] = 0;
    sync();
    if (tid == 0) {
        unsigned long long *nw = K.nw;
        int *parent = K.parent;
        int leaf = 0, inner = n, made = n;              // heads of the two queues, next node to make
        while (made < 2 * n - 1) {
            int pick[2];
            for (int j = 0; j < 2; j++) {
                const bool take_leaf = leaf < n && (inner >= made || nw[leaf] <= nw[inner]);
                pick[j] = take_leaf ? leaf++ : inner++;
            }
            nw[made] = nw[pick[0]] + nw[pick[1]];
            parent[pick[0]] = parent[pick[1]] = made;
            made++;
        }
        K.node_depth[2 * n - 2] = 0;
        for (int i = 2 * n - 3; i >= n; i--) K.node_depth[i] = K.node_depth[parent[i]] + 1;
    }
    sync();
    for (int i = tid; i < n; i += nt) {
        const int d = K.node_depth[K.parent[i]] + 1;
        gz_count(num + (d > 15 ? 15 : d));
    }
    sync();
    if (tid == 0) {
        uint32_t kraft = 0;                             // in units of 2^-15
        for (int l = 1; l <= 15; l++) kraft += (uint32_t)num[l] << (15 - l);
        while (kraft > (1u << 15)) {
            num[15]--;
            for (int l = 14; l >= 1; l--)
                if (num[l]) { num[l]--; num[l + 1] += 2; break; }
            kraft--;
        }
        // K.first_rank[l]: the first of the symbols, in ascending weight, that gets length l
        int acc = 0;
        for (int l = 15; l >= 1; l--) { K.first_rank[l] = acc; acc += num[l]; }
        int code = 0;
        next_code[0] = 0;
        num[0] = 0;
        for (int bits = 1; bits < 16; bits++) { code = (code + num[bits - 1]) << 1; next_code[bits] = code; }
        uint32_t pos = 0;
        gz_put(K.hdr, pos, 1u, 1);              // BFINAL
        gz_put(K.hdr, pos, 2u, 2);              // BTYPE = dynamic Huffman
        gz_put(K.hdr, pos, 0u, 5);              // HLIT: 257 literal/length codes
        gz_put(K.hdr, pos, 0u, 5);              // HDIST: 1 distance code (of length 0: there are no matches)
        gz_put(K.hdr, pos, 15u, 4);             // HCLEN: all 19 code-length codes
        const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        for (int i = 0; i < 19; i++) gz_put(K.hdr, pos, order[i] <= 12 ? 4u : 5u, 3);
    }
    sync();
    for (int i = tid; i < n; i += nt) {
        int l = 15;
        while (l > 1 && !(i >= K.first_rank[l] && i < K.first_rank[l] + num[l])) l--;
        depth[K.order[i]] = l;
    }
    sync();
    for (int s = tid; s < 258; s += nt) {
        const int d = s < 257 ? depth[s] : 0;           // s == 257: the one distance code, unused
        int same = 0;
        uint32_t pos = GZ_HDR_FIXED_BITS;
        for (int t = 0; t < s; t++) {
            const int dt = depth[t];
            same += dt == d ? 1 : 0;
            pos += dt <= 12 ? 4u : 5u;
        }
        if (s < 257) {
            const uint32_t c = gz_rev((uint32_t)(next_code[d] + same), d);
            T.len[s] = (uint8_t)d;
            T.code[s] = (uint16_t)c;
            T.sym[s] = c | ((uint32_t)d << 16);
        }
        if (d <= 12) gz_or_bits(K.hdr, pos, gz_rev((uint32_t)d, 4), 4);
        else gz_or_bits(K.hdr, pos, gz_rev((uint32_t)(26 + (d - 13)), 5), 5);
        if (s == 257) { T.hdr_nbits = pos + 4u; T.pad_ = 0; }
    }
    sync();
    for (int i = tid; i < GZ_HDR_WORDS; i += nt) T.hdr[i] = K.hdr[i];
}
// the whole build on one thread (host simulation)
inline void gz_build_table(const unsigned long long *hist, GzTable &T, GzWork &K)
{
    for (int s = 0; s < 257; s++) K.w[s] = gz_weight(hist, s);
    for (int s = 0; s < 257; s++) K.order[gz_rank(K.w, s)] = s;
    gz_build_from_sorted(T, K, 0, 1, GzNoSync());
}

// Chunk k of the member whose text is [start, end): the part of it inside the GZ_CHUNK-aligned block k of the text
// counted from the block that holds `start`.
ORC_HD uint32_t gz_member_chunks(uint64_t start, uint64_t end)
{
    if (end <= start) return 0u;
    return (uint32_t)((end - (start & ~(uint64_t)(GZ_CHUNK - 1)) + GZ_CHUNK - 1) / GZ_CHUNK);
}
ORC_HD void gz_chunk_range(uint64_t start, uint64_t end, uint32_t k, uint64_t &lo, uint64_t &hi)
{
    const uint64_t a = (start & ~(uint64_t)(GZ_CHUNK - 1)) + (uint64_t)k * GZ_CHUNK;
    lo = a < start ? start : a;
    hi = a + GZ_CHUNK < end ? a + GZ_CHUNK : end;
}

// The histogram looks at one 16-byte vector of the text in GZ_SAMPLE: the first 128 bytes (GZ_RUN vectors, one DRAM
// line) of every 128 * GZ_SAMPLE bytes, and the tail behind the last vector.
constexpr int GZ_SAMPLE = 16, GZ_RUN = 8;
// the i-th sampled vector (i < gz_n_samples(n16)) of a text of n16 whole vectors
ORC_HD uint64_t gz_n_samples(uint64_t n16)
{
    const uint64_t period = (uint64_t)GZ_RUN * GZ_SAMPLE, rest = n16 % period;
    return n16 / period * GZ_RUN + (rest < (uint64_t)GZ_RUN ? rest : (uint64_t)GZ_RUN);
}
ORC_HD uint64_t gz_sample_at(uint64_t i) { return i / GZ_RUN * ((uint64_t)GZ_RUN * GZ_SAMPLE) + i % GZ_RUN; }

#if defined(__CUDACC__)
typedef uint4 gz_vec16;
#else
struct alignas(16) gz_vec16 { uint32_t x, y, z, w; };     // the host simulation's stand-in for uint4
#endif

// byte k of a word, zero-extended: one PRMT on the device
ORC_HD uint32_t gz_byte(uint32_t x, int k) { return byte_perm(x, 0u, 0x4440u + (uint32_t)k); }

// Bits the chunk's bytes take, and their CRC-32 (slicing-by-16).  len: GzTable.len, crc_tab: GzTable.crc_tab (both in
// shared memory on the device).  16 bytes at a time between the first and the last 16-byte boundary of the chunk (the
// text starts at a 16-byte-aligned address).
ORC_HD void gz_chunk_measure(const uint8_t *__restrict__ text, uint64_t lo, uint64_t hi, const uint8_t *len,
                             const uint32_t (*crc_tab)[256], uint32_t &bits, uint32_t &crc_out)
{
    uint32_t nb = 0, crc = 0xFFFFFFFFu;
    uint64_t p = lo;
    for (; p < hi && (p & 15u); p++) {
        const uint8_t c = text[p];
        nb += len[c];
        crc = crc_tab[0][(crc ^ c) & 255u] ^ (crc >> 8);
    }
#define GZ_MEASURE16(q)                                                                                           \
    {                                                                                                             \
        const uint32_t w[4] = {(q).x, (q).y, (q).z, (q).w};                                                        \
        for (int k = 0; k < 4; k++)                                                                               \
            nb += (uint32_t)len[gz_byte(w[k], 0)] + len[gz_byte(w[k], 1)] + len[gz_byte(w[k], 2)] + len[gz_byte(w[k], 3)]; \
        const uint32_t x0 = crc ^ w[0];                                                                           \
        crc = crc_tab[15][gz_byte(x0, 0)] ^ crc_tab[14][gz_byte(x0, 1)] ^ crc_tab[13][gz_byte(x0, 2)] ^ crc_tab[12][gz_byte(x0, 3)] ^ \
              crc_tab[11][gz_byte(w[1], 0)] ^ crc_tab[10][gz_byte(w[1], 1)] ^ crc_tab[9][gz_byte(w[1], 2)] ^ crc_tab[8][gz_byte(w[1], 3)] ^ \
              crc_tab[7][gz_byte(w[2], 0)] ^ crc_tab[6][gz_byte(w[2], 1)] ^ crc_tab[5][gz_byte(w[2], 2)] ^ crc_tab[4][gz_byte(w[2], 3)] ^ \
              crc_tab[3][gz_byte(w[3], 0)] ^ crc_tab[2][gz_byte(w[3], 1)] ^ crc_tab[1][gz_byte(w[3], 2)] ^ crc_tab[0][gz_byte(w[3], 3)]; \
    }
    for (; p + 16 <= hi && (p & 63u); p += 16) GZ_MEASURE16(*reinterpret_cast<const gz_vec16 *>(text + p));
    // 64 bytes at a time, the four loads issued together: every lane walks its own chunk, so a warp's load touches 32
    // lines -- the other three loads of the group find them in L1 while they are still there
    for (; p + 64 <= hi; p += 64) {
        const gz_vec16 *v = reinterpret_cast<const gz_vec16 *>(text + p);
        const gz_vec16 q0 = v[0], q1 = v[1], q2 = v[2], q3 = v[3];
        GZ_MEASURE16(q0); GZ_MEASURE16(q1); GZ_MEASURE16(q2); GZ_MEASURE16(q3);
    }
    for (; p + 16 <= hi; p += 16) GZ_MEASURE16(*reinterpret_cast<const gz_vec16 *>(text + p));
#undef GZ_MEASURE16
    for (; p < hi; p++) {
        const uint8_t c = text[p];
        nb += len[c];
        crc = crc_tab[0][(crc ^ c) & 255u] ^ (crc >> 8);
    }
    bits = nb;
    crc_out = crc ^ 0xFFFFFFFFu;
}

// The chunk's bytes as codes, from bit `bit0` of the (zeroed) output on.  The first word (it may hold bits of the
// chunk in front) and the last one go in by OR, the words between by plain stores: byte by byte until the first
// word is out and the text stands at a 16-byte boundary, then 16 bytes at a time with nothing but predicated
// stores in the loop.
ORC_HD uint32_t gz_spill(uint32_t v, uint32_t sh)      // the bits of v that leave a 32-bit word when v moves up by sh < 32
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(v, 0u, sh);
#else
    return sh ? v >> (32u - sh) : 0u;
#endif
}
ORC_HD void gz_chunk_encode(const uint8_t *__restrict__ text, uint64_t lo, uint64_t hi, const uint32_t *sym,
                            uint64_t bit0, uint32_t *__restrict__ out)
{
    uint32_t *const base = out + (bit0 >> 5);
    uint32_t wo = 0;                                    // word of `base` the accumulator's low half belongs to
    uint32_t fill = (uint32_t)(bit0 & 31u);
    uint32_t a_lo = 0, a_hi = 0;                        // the accumulator: fill bits, fill < 32 between the steps
    // two codes add 30 bits at most, so the 64 bits hold them
#define GZ_BYTE(c)                                                                      \
    { const uint32_t e = sym[c], v = e & 0xFFFFu;                                       \
      a_hi |= gz_spill(v, fill); a_lo |= v << fill; fill += e >> 16; }
#define GZ_PAIR(c0, c1)                                                                 \
    { const uint32_t e0 = sym[c0], e1 = sym[c1];                                        \
      const uint32_t v = (e0 & 0xFFFFu) | ((e1 & 0xFFFFu) << (e0 >> 16));               \
      a_hi |= gz_spill(v, fill); a_lo |= v << fill; fill += (e0 >> 16) + (e1 >> 16); }
#define GZ_TAKE_WORD()                                                                  \
    { const bool full = fill >= 32u;                                                    \
      if (full) base[wo] = a_lo;                                                        \
      a_lo = full ? a_hi : a_lo; a_hi = full ? 0u : a_hi;                               \
      wo += full ? 1u : 0u; fill -= full ? 32u : 0u; }
#define GZ_TAKE_WORD_EDGE()                                                             \
    if (fill >= 32u) {                                                                  \
        if (first) { gz_or(base + wo, a_lo); first = false; }                           \
        else base[wo] = a_lo;                                                           \
        a_lo = a_hi; a_hi = 0u; wo++; fill -= 32u;                                      \
    }
    uint64_t p = lo;
    bool first = true;
    for (; p < hi && (first || (p & 15u)); p++) { GZ_BYTE(text[p]); GZ_TAKE_WORD_EDGE(); }
#define GZ_ENCODE16(q)                                                                  \
    {                                                                                   \
        const uint32_t v4[4] = {(q).x, (q).y, (q).z, (q).w};                              \
        for (int k = 0; k < 4; k++) {                                                   \
            GZ_PAIR(gz_byte(v4[k], 0), gz_byte(v4[k], 1)); GZ_TAKE_WORD();              \
            GZ_PAIR(gz_byte(v4[k], 2), gz_byte(v4[k], 3)); GZ_TAKE_WORD();              \
        }                                                                               \
    }
    for (; p + 16 <= hi && (p & 63u); p += 16) GZ_ENCODE16(*reinterpret_cast<const gz_vec16 *>(text + p));
    for (; p + 64 <= hi; p += 64) {             // the four loads of a group issued together (see gz_chunk_measure)
        const gz_vec16 *v = reinterpret_cast<const gz_vec16 *>(text + p);
        const gz_vec16 q0 = v[0], q1 = v[1], q2 = v[2], q3 = v[3];
        GZ_ENCODE16(q0); GZ_ENCODE16(q1); GZ_ENCODE16(q2); GZ_ENCODE16(q3);
    }
    for (; p + 16 <= hi; p += 16) GZ_ENCODE16(*reinterpret_cast<const gz_vec16 *>(text + p));
#undef GZ_ENCODE16
    for (; p < hi; p++) { GZ_BYTE(text[p]); GZ_TAKE_WORD_EDGE(); }
#undef GZ_BYTE
#undef GZ_PAIR
#undef GZ_TAKE_WORD
#undef GZ_TAKE_WORD_EDGE
    if (fill) gz_or(base + wo, a_lo);
}

// Per member, after its chunks were measured: the frame around the codes.  data_bits = the chunks' bits.
ORC_HD uint64_t gz_member_bytes(const GzTable &T, uint64_t data_bits)
{
    return (uint64_t)GZ_HEADER_BYTES + ((uint64_t)T.hdr_nbits + data_bits + T.len[GZ_EOB] + 7u) / 8u + GZ_TRAILER_BYTES;
}
// (out_bytes: the whole output buffer, 4-byte aligned and zeroed)
ORC_HD void gz_member_frame(const GzTable &T, uint8_t *out_bytes, uint64_t member_off, uint64_t member_bytes,
                            uint64_t data_bits, uint32_t crc, uint64_t text_bytes)
{
    // every byte and bit goes in by OR on 32-bit words: the words at the seams (last header byte / first codes,
    // last codes / trailer, trailer / next member's header) are shared with other threads' atomic ORs, which a
    // plain byte store beside them would race with
    uint32_t *words = reinterpret_cast<uint32_t *>(out_bytes);
    auto put_byte = [&](uint64_t at, uint32_t v) { gz_or(words + (at >> 2), v << (8u * (uint32_t)(at & 3u))); };
    const uint8_t head[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 255, 8, 0, 'O', 'C', 4, 0};
    for (int i = 0; i < 16; i++) put_byte(member_off + (uint64_t)i, head[i]);
    for (int i = 0; i < 4; i++) put_byte(member_off + 16u + (uint64_t)i, (uint32_t)(member_bytes >> (8 * i)) & 255u);
    // block header and end-of-block code, bit by word (the words are shared with the first / last chunk)
    uint64_t bit = 8u * (member_off + GZ_HEADER_BYTES);
    for (uint32_t done = 0; done < T.hdr_nbits; ) {
        const uint32_t take = T.hdr_nbits - done < 32u ? T.hdr_nbits - done : 32u;
        uint64_t v = T.hdr[done >> 5];
        if (take < 32u) v &= (1ull << take) - 1ull;
        v <<= (bit & 31u);
        gz_or(words + (bit >> 5), (uint32_t)v);
        if (v >> 32) gz_or(words + (bit >> 5) + 1, (uint32_t)(v >> 32));
        bit += take;
        done += take;
    }
    bit += data_bits;
    {
        const uint64_t v = (uint64_t)T.code[GZ_EOB] << (bit & 31u);
        gz_or(words + (bit >> 5), (uint32_t)v);
        if (v >> 32) gz_or(words + (bit >> 5) + 1, (uint32_t)(v >> 32));
    }
    const uint64_t t = member_off + member_bytes - GZ_TRAILER_BYTES;
    for (int i = 0; i < 4; i++) {
        put_byte(t + (uint64_t)i, (crc >> (8 * i)) & 255u);
        put_byte(t + 4u + (uint64_t)i, (uint32_t)(text_bytes >> (8 * i)) & 255u);
    }
}

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------ kernels
// hist[b] += occurrences of byte b in the sampled vectors of text[0, *total) (gz_sample_at) and in the tail behind
// the last whole vector
__global__ void __launch_bounds__(256)
gz_hist_kernel(const uint8_t *__restrict__ text, const uint64_t *__restrict__ total, unsigned long long *__restrict__ hist)
{
    __shared__ uint32_t s_h[8][256];            // one histogram per warp: fewer same-address atomics
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&s_h[0][0])[i] = 0;
    __syncthreads();
    uint32_t *h = s_h[threadIdx.x >> 5];
    const uint64_t n16 = *total >> 4;
    const uint64_t n_samples = gz_n_samples(n16);
    const uint4 *t4 = reinterpret_cast<const uint4 *>(text);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_samples; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 v = t4[gz_sample_at(i)];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            atomicAdd(&h[w[k] & 255u], 1u); atomicAdd(&h[(w[k] >> 8) & 255u], 1u);
            atomicAdd(&h[(w[k] >> 16) & 255u], 1u); atomicAdd(&h[w[k] >> 24], 1u);
        }
    }
    if (blockIdx.x == 0)
        for (uint64_t p = (n16 << 4) + threadIdx.x; p < *total; p += blockDim.x) atomicAdd(&h[text[p]], 1u);
    __syncthreads();
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) sum += s_h[k][threadIdx.x];
    if (sum) atomicAdd(hist + threadIdx.x, (unsigned long long)sum);
}

// one block: the code of the batch; chunk_base[m] = chunks of the members in front of member m
__global__ void __launch_bounds__(512)
gz_table_kernel(const unsigned long long *__restrict__ hist, GzTable *__restrict__ T, int n_members,
                const uint64_t *__restrict__ bin_offsets, uint32_t *__restrict__ chunk_base, int flat)
{
    __shared__ GzWork K;
    __shared__ uint32_t s_chunks[MAX_BINS_GZ];
    const int t = threadIdx.x;
    // flat: every symbol the same weight -- 8- and 9-bit codes, a member can then not exceed 9/8 of its text (what
    // orc_wait() falls back to when a batch's sampled histogram was so far off its bytes that the arena overflowed)
    if (t < 257) K.w[t] = flat ? 1ull : gz_weight(hist, t);
    for (int m = t; m < n_members; m += blockDim.x) s_chunks[m] = gz_member_chunks(bin_offsets[m], bin_offsets[m + 1]);
    __syncthreads();
    if (t < 257) K.order[gz_rank(K.w, t)] = t;
    if (t == 257) {
        uint32_t acc = 0;
        for (int m = 0; m < n_members; m++) { chunk_base[m] = acc; acc += s_chunks[m]; }
        chunk_base[n_members] = acc;
    }
    __syncthreads();
    gz_build_from_sorted(*T, K, t, (int)blockDim.x, GzBlockSync());
}

__device__ __forceinline__ int gz_member_of(const uint32_t *s_base, int n_members, uint32_t c)
{
    int lo = 0, hi = n_members;         // last m with s_base[m] <= c
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_base[mid] <= c) lo = mid; else hi = mid;
    }
    return lo;
}

// exclusive prefix sum over the threads of a block (blockDim.x a multiple of 32, <= 1024); total: the block's sum
template <typename V>
__device__ __forceinline__ V gz_block_scan(V v, V *s_warp /* [33] */, V &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    V incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const V o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    __syncthreads();                    // s_warp may still be read from the scan before
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        V w = lane < n_warps ? s_warp[lane] : (V)0, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const V o = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += o;
        }
        if (lane < n_warps) s_warp[lane] = wi - w;
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    total = s_warp[32];
    return s_warp[warp] + incl - v;
}

// One thread per chunk, tiles of GZ_TILE chunks per block: the chunk's bits as an exclusive sum inside its tile
// (chunk_local) and the tile's sum (tile_bits); the chunk's CRC moved to where the chunk stands in its member
// (past the bytes behind it) and xor-ed into the member's CRC -- crc32_combine is linear, so the order in which
// the chunks arrive does not matter.  member_crc must be zero before.
__global__ void __launch_bounds__(GZ_TILE)
gz_measure_kernel(const uint8_t *__restrict__ text, const uint64_t *__restrict__ bin_offsets, int n_members,
                  const uint32_t *__restrict__ chunk_base, const GzTable *__restrict__ T,
                  uint32_t *__restrict__ chunk_local, uint32_t *__restrict__ tile_bits, uint32_t *__restrict__ member_crc)
{
    __shared__ uint32_t s_base[MAX_BINS_GZ + 1];
    __shared__ uint32_t s_crc[16][256];
    __shared__ uint32_t s_pow[32][32];
    __shared__ uint8_t s_len[260];
    __shared__ uint32_t s_warp[33];
    for (int i = threadIdx.x; i <= n_members; i += blockDim.x) s_base[i] = chunk_base[i];
    for (int i = threadIdx.x; i < 16 * 256; i += blockDim.x) (&s_crc[0][0])[i] = (&T->crc_tab[0][0])[i];
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) (&s_pow[0][0])[i] = (&T->crc_pow[0][0])[i];
    for (int i = threadIdx.x; i < 257; i += blockDim.x) s_len[i] = T->len[i];
    __syncthreads();
    const uint32_t n_chunks = s_base[n_members];
    const uint32_t n_tiles = (n_chunks + GZ_TILE - 1) / GZ_TILE;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t c = tile * GZ_TILE + threadIdx.x;
        uint32_t bits = 0;
        if (c < n_chunks) {
            const int m = gz_member_of(s_base, n_members, c);
            const uint64_t end = bin_offsets[m + 1];
            uint64_t lo, hi;
            gz_chunk_range(bin_offsets[m], end, c - s_base[m], lo, hi);
            uint32_t crc;
            gz_chunk_measure(text, lo, hi, s_len, s_crc, bits, crc);
            const uint32_t moved = gz_crc_shift(s_pow, crc, end - hi);
            if (moved) atomicXor(member_crc + m, moved);
        }
        uint32_t total;
        const uint32_t before = gz_block_scan(bits, s_warp, total);
        if (c < n_chunks) chunk_local[c] = before;
        if (threadIdx.x == 0) tile_bits[tile] = total;
    }
}

// One block: where the tiles' bits start (tile_off, exclusive; tile_off[n_tiles] = all bits), then per member
// where its bits start in that count (member_pos), how many they are, its size as a gzip member, and where the
// members go in the output (gz_offsets[n_members] = all of them).  A bin without text gets no member.
__global__ void __launch_bounds__(1024)
gz_layout_kernel(const uint64_t *__restrict__ bin_offsets, int n_members, const uint32_t *__restrict__ chunk_base,
                 const GzTable *__restrict__ T, const uint32_t *__restrict__ chunk_local, const uint32_t *__restrict__ tile_bits,
                 uint64_t *__restrict__ tile_off, uint64_t *__restrict__ member_pos, uint64_t *__restrict__ member_bits,
                 uint64_t *__restrict__ member_bytes, uint64_t *__restrict__ gz_offsets)
{
    __shared__ unsigned long long s_warp[33];
    __shared__ unsigned long long s_pos[MAX_BINS_GZ + 1];
    const uint32_t n_chunks = chunk_base[n_members];
    const uint32_t n_tiles = (n_chunks + GZ_TILE - 1) / GZ_TILE;
    const uint32_t per = (n_tiles + blockDim.x - 1) / blockDim.x;
    const uint32_t t0 = min(n_tiles, threadIdx.x * per), t1 = min(n_tiles, t0 + per);
    unsigned long long sum = 0;
    for (uint32_t t = t0; t < t1; t++) sum += tile_bits[t];
    unsigned long long all_bits;
    unsigned long long run = gz_block_scan(sum, s_warp, all_bits);
    for (uint32_t t = t0; t < t1; t++) { tile_off[t] = run; run += tile_bits[t]; }
    if (threadIdx.x == 0) tile_off[n_tiles] = all_bits;
    __syncthreads();                    // tile_off is read back below
    for (int m = threadIdx.x; m <= n_members; m += blockDim.x) {
        const uint32_t c = chunk_base[m];
        s_pos[m] = (m < n_members && c < n_chunks) ? tile_off[c / GZ_TILE] + chunk_local[c] : all_bits;
    }
    __syncthreads();
    unsigned long long bytes = 0;
    const int m = threadIdx.x;          // n_members <= MAX_BINS_GZ <= blockDim.x
    if (m < n_members) {
        const unsigned long long bits = s_pos[m + 1] - s_pos[m];
        bytes = bin_offsets[m + 1] > bin_offsets[m] ? gz_member_bytes(*T, bits) : 0ull;
        member_pos[m] = s_pos[m];
        member_bits[m] = bits;
        member_bytes[m] = bytes;
    }
    unsigned long long all_bytes;
    const unsigned long long off = gz_block_scan(bytes, s_warp, all_bytes);
    if (m < n_members) gz_offsets[m] = off;
    if (m == 0) gz_offsets[n_members] = all_bytes;
}

// zeroes out[0, gz_offsets[n_members]) rounded up to 16 bytes: everything below is written by OR
__global__ void __launch_bounds__(256)
gz_zero_kernel(const uint64_t *__restrict__ gz_offsets, int n_members, uint64_t cap_bytes, uint4 *__restrict__ out)
{
    const uint64_t total = gz_offsets[n_members] < cap_bytes ? gz_offsets[n_members] : cap_bytes;
    const uint64_t n16 = (total + 15u) >> 4;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = make_uint4(0u, 0u, 0u, 0u);
}

// one thread per chunk: its codes; the first threads also write the members' frames
__global__ void __launch_bounds__(128)
gz_encode_kernel(const uint8_t *__restrict__ text, const uint64_t *__restrict__ bin_offsets, int n_members,
                 const uint32_t *__restrict__ chunk_base, const GzTable *__restrict__ T,
                 const uint32_t *__restrict__ chunk_local, const uint64_t *__restrict__ tile_off,
                 const uint64_t *__restrict__ member_pos, const uint64_t *__restrict__ member_bits,
                 const uint32_t *__restrict__ member_crc, const uint64_t *__restrict__ member_bytes,
                 const uint64_t *__restrict__ gz_offsets, uint64_t cap_bytes, uint8_t *__restrict__ out)
{
    __shared__ uint32_t s_base[MAX_BINS_GZ + 1];
    __shared__ uint32_t s_sym[257];
    for (int i = threadIdx.x; i <= n_members; i += blockDim.x) s_base[i] = chunk_base[i];
    for (int i = threadIdx.x; i < 257; i += blockDim.x) s_sym[i] = T->sym[i];
    __syncthreads();
    if (gz_offsets[n_members] > cap_bytes) return;         // orc_wait() reports it
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < (uint32_t)n_members && member_bytes[gid])
        gz_member_frame(*T, out, gz_offsets[gid], member_bytes[gid], member_bits[gid], member_crc[gid],
                        bin_offsets[gid + 1] - bin_offsets[gid]);
    const uint32_t n_chunks = s_base[n_members];
    const uint32_t hdr_nbits = T->hdr_nbits;
    for (uint32_t c = gid; c < n_chunks; c += gridDim.x * blockDim.x) {
        const int m = gz_member_of(s_base, n_members, c);
        uint64_t lo, hi;
        gz_chunk_range(bin_offsets[m], bin_offsets[m + 1], c - s_base[m], lo, hi);
        const uint64_t bit0 = 8u * (gz_offsets[m] + GZ_HEADER_BYTES) + hdr_nbits + (tile_off[c / GZ_TILE] + chunk_local[c] - member_pos[m]);
        gz_chunk_encode(text, lo, hi, s_sym, bit0, reinterpret_cast<uint32_t *>(out));
    }
}
#endif  // __CUDACC__

}  // namespace orc
