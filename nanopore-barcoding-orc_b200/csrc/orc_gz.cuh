// orc_gz.cuh -- gzip members made on the device (orc_params.emit_gzip).
//
// What the reference leaves on disk are .fastq.gz files (02_cutadapt_loop.sh:64-72, :94-102: `-o ...fastq.gz`);
// the bin-major FASTQ text of a batch becomes one gzip member per bin here, so that only the compressed bytes
// cross PCIe and the host appends them to the bin files as they are (a concatenation of gzip members is a
// gzip file; orc_io.cpp writes the same member format, with the "OC" size field its reader uses to inflate
// members in parallel).
//
// A member holds ONE dynamic-Huffman DEFLATE block of literals only (no LZ77 matches: bases and qualities of
// nanopore reads hardly repeat, zlib's own matches gain a few per cent on them): the code is built per batch
// from the byte histogram of the batch's whole FASTQ text, so every member of the batch carries the same block
// header.  Encoding is then a table look-up per byte; where a byte's bits go is a prefix sum of code lengths,
// taken per 512-byte chunk, per member, per batch.  CRC-32 of a member = the chunks' CRCs folded with the
// "append n zero bytes" operator (zlib's crc32_combine), whose matrices for n = 2^k come from the host.
//
// Everything that computes is ORC_HD so that tests/gzsim.cpp can run the same code on the CPU and hand the
// members to zlib (tests/test_gz.py); the kernels at the end only distribute the work.
#pragma once
#include <stdint.h>

#include "orc_core.cuh"

namespace orc {

constexpr int GZ_CHUNK = 512;               // bytes of text per encoding thread (chunks are 512-aligned in the text)
constexpr int GZ_HEADER_BYTES = 20;         // 10 + XLEN + "OC" subfield with the member's size (orc_io.cpp GZ_SIZE_AT)
constexpr int GZ_TRAILER_BYTES = 8;         // CRC-32, ISIZE
constexpr int GZ_HDR_WORDS = 44;            // block header: 3 + 5 + 5 + 4 + 19 * 3 + 258 * 5 = 1364 bits at most
constexpr int GZ_EOB = 256;
constexpr int MAX_BINS_GZ = 512;            // == MAX_BINS (orc_kernels.cuh)

struct GzTable {
    uint16_t code[257];     // Huffman code of byte b (256: end of block), bit-reversed: DEFLATE packs codes MSB first
    uint8_t len[257];       // its length in bits; 0: b does not occur in this batch
    uint8_t pad_;
    uint32_t hdr_nbits;
    uint32_t hdr[GZ_HDR_WORDS];     // BFINAL BTYPE HLIT HDIST HCLEN, the code-length code, the 258 code lengths
    uint32_t crc_tab[256];          // CRC-32 (0xEDB88320) byte table
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
        T.crc_tab[n] = c;
    }
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

// The batch's code from its byte histogram (one thread: at most 257 symbols).  Code lengths are limited to 15
// bits by flattening the weights and building again, which keeps the code complete (inflate rejects codes
// that are not).  At least two symbols get a code, so that the code is never a single one-bit code.
struct GzWork {             // scratch of gz_build_table (shared memory on the device)
    unsigned long long w[257], nw[2 * 257];
    int parent[2 * 257], leaf_of[257], depth[257];
    uint8_t alive[2 * 257];
};
ORC_HD void gz_build_table(const unsigned long long *hist, GzTable &T, GzWork &K)
{
    unsigned long long *w = K.w;
    for (int b = 0; b < 256; b++) w[b] = hist[b];
    w[GZ_EOB] = 1;
    {
        int n = 0, first = -1;
        for (int b = 0; b < 256; b++) if (w[b]) { n++; if (first < 0) first = b; }
        if (n == 0) w['\n'] = 1;            // an empty batch still has a valid (unused) code: EOB and one literal
    }
    int *depth = K.depth;
    for (;;) {
        // Huffman by repeated selection of the two lightest nodes (n <= 257: a few ten thousand steps at worst)
        unsigned long long *nw = K.nw;
        int *parent = K.parent, *leaf_of = K.leaf_of;
        uint8_t *alive = K.alive;
        int n_nodes = 0;
        for (int s = 0; s < 257; s++) {
            leaf_of[s] = -1;
            if (w[s]) { leaf_of[s] = n_nodes; nw[n_nodes] = w[s]; parent[n_nodes] = -1; alive[n_nodes] = 1; n_nodes++; }
        }
        int n_alive = n_nodes;
        while (n_alive > 1) {
            int a = -1, b = -1;
            for (int i = 0; i < n_nodes; i++) {
                if (!alive[i]) continue;
                if (a < 0 || nw[i] < nw[a]) { b = a; a = i; }
                else if (b < 0 || nw[i] < nw[b]) b = i;
            }
            nw[n_nodes] = nw[a] + nw[b];
            parent[n_nodes] = -1;
            alive[n_nodes] = 1;
            parent[a] = parent[b] = n_nodes;
            alive[a] = alive[b] = 0;
            n_nodes++;
            n_alive--;
        }
        int deepest = 0;
        for (int s = 0; s < 257; s++) {
            depth[s] = 0;
            if (leaf_of[s] < 0) continue;
            int d = 0;
            for (int i = leaf_of[s]; parent[i] >= 0; i = parent[i]) d++;
            depth[s] = d;
            if (d > deepest) deepest = d;
        }
        if (deepest <= 15) break;
        for (int s = 0; s < 257; s++) if (w[s]) w[s] = (w[s] >> 2) + 1;     // flatter weights, shallower tree
    }
    // canonical codes (RFC 1951, 3.2.2)
    int bl_count[16], next_code[16];
    for (int i = 0; i < 16; i++) bl_count[i] = 0;
    for (int s = 0; s < 257; s++) if (depth[s]) bl_count[depth[s]]++;
    int code = 0;
    next_code[0] = 0;
    for (int bits = 1; bits < 16; bits++) { code = (code + bl_count[bits - 1]) << 1; next_code[bits] = code; }
    for (int s = 0; s < 257; s++) {
        T.len[s] = (uint8_t)depth[s];
        T.code[s] = depth[s] ? (uint16_t)gz_rev((uint32_t)next_code[depth[s]]++, depth[s]) : (uint16_t)0;
    }
    T.pad_ = 0;
    // the block header.  The code-length code is a fixed complete code: the lengths 0..12 take four bits
    // (codes 0..12), the lengths 13..15 and the three repeat symbols five bits (codes 26..31); the lengths are
    // then written one by one, without the repeat symbols (about 140 bytes per member).
    for (int i = 0; i < GZ_HDR_WORDS; i++) T.hdr[i] = 0;
    uint32_t pos = 0;
    gz_put(T.hdr, pos, 1u, 1);              // BFINAL
    gz_put(T.hdr, pos, 2u, 2);              // BTYPE = dynamic Huffman
    gz_put(T.hdr, pos, 0u, 5);              // HLIT: 257 literal/length codes
    gz_put(T.hdr, pos, 0u, 5);              // HDIST: 1 distance code (of length 0: there are no matches)
    gz_put(T.hdr, pos, 15u, 4);             // HCLEN: all 19 code-length codes
    const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    for (int i = 0; i < 19; i++) gz_put(T.hdr, pos, order[i] <= 12 ? 4u : 5u, 3);
    for (int s = 0; s < 258; s++) {
        const int v = s < 257 ? depth[s] : 0;       // s == 257: the distance code
        if (v <= 12) gz_put(T.hdr, pos, gz_rev((uint32_t)v, 4), 4);
        else gz_put(T.hdr, pos, gz_rev((uint32_t)(26 + (v - 13)), 5), 5);
    }
    T.hdr_nbits = pos;
}

// Chunk k of the member whose text is [start, end): the part of it inside the 512-aligned block k of the text
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

// Bits the chunk's bytes take, and their CRC-32.
ORC_HD void gz_chunk_measure(const uint8_t *__restrict__ text, uint64_t lo, uint64_t hi, const uint8_t *len,
                             const uint32_t *crc_tab, uint32_t &bits, uint32_t &crc_out)
{
    uint32_t nb = 0, crc = 0xFFFFFFFFu;
    for (uint64_t p = lo; p < hi; p++) {
        const uint8_t c = text[p];
        nb += len[c];
        crc = crc_tab[(crc ^ c) & 255u] ^ (crc >> 8);
    }
    bits = nb;
    crc_out = crc ^ 0xFFFFFFFFu;
}

// or a 32-bit word into the output (words at the borders of a chunk are shared with its neighbours)
ORC_HD void gz_or(uint32_t *p, uint32_t v)
{
#if defined(__CUDA_ARCH__)
    if (v) atomicOr(p, v);
#else
    *p |= v;
#endif
}

// The chunk's bytes as codes, from bit `bit0` of the (zeroed) output on.
ORC_HD void gz_chunk_encode(const uint8_t *__restrict__ text, uint64_t lo, uint64_t hi, const uint16_t *code,
                            const uint8_t *len, uint64_t bit0, uint32_t *__restrict__ out)
{
    uint64_t wi = bit0 >> 5;
    uint32_t fill = (uint32_t)(bit0 & 31u);
    uint64_t acc = 0;
    bool first = true;          // the first word may hold bits of the chunk in front: or, do not store
    for (uint64_t p = lo; p < hi; p++) {
        const uint8_t c = text[p];
        acc |= (uint64_t)code[c] << fill;
        fill += len[c];
        if (fill >= 32u) {
            if (first) { gz_or(out + wi, (uint32_t)acc); first = false; }
            else out[wi] = (uint32_t)acc;
            wi++;
            acc >>= 32;
            fill -= 32u;
        }
    }
    if (fill) gz_or(out + wi, (uint32_t)acc);
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
// hist[b] += occurrences of byte b in text[0, *total)
__global__ void __launch_bounds__(256)
gz_hist_kernel(const uint8_t *__restrict__ text, const uint64_t *__restrict__ total, unsigned long long *__restrict__ hist)
{
    __shared__ uint32_t s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t n16 = *total >> 4;
    const uint4 *t4 = reinterpret_cast<const uint4 *>(text);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 v = t4[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            atomicAdd(&s_h[w[k] & 255u], 1u); atomicAdd(&s_h[(w[k] >> 8) & 255u], 1u);
            atomicAdd(&s_h[(w[k] >> 16) & 255u], 1u); atomicAdd(&s_h[w[k] >> 24], 1u);
        }
    }
    if (blockIdx.x == 0)
        for (uint64_t p = (n16 << 4) + threadIdx.x; p < *total; p += blockDim.x) atomicAdd(&s_h[text[p]], 1u);
    __syncthreads();
    if (s_h[threadIdx.x]) atomicAdd(hist + threadIdx.x, (unsigned long long)s_h[threadIdx.x]);
}

// one thread: the code of the batch; chunk_base[m] = chunks of the members in front of member m
__global__ void gz_table_kernel(const unsigned long long *__restrict__ hist, GzTable *__restrict__ T, int n_members,
                                const uint64_t *__restrict__ bin_offsets, uint32_t *__restrict__ chunk_base)
{
    __shared__ GzWork K;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    gz_build_table(hist, *T, K);
    uint32_t acc = 0;
    for (int m = 0; m < n_members; m++) { chunk_base[m] = acc; acc += gz_member_chunks(bin_offsets[m], bin_offsets[m + 1]); }
    chunk_base[n_members] = acc;
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

// one thread per chunk: its bits and its CRC
__global__ void __launch_bounds__(128)
gz_measure_kernel(const uint8_t *__restrict__ text, const uint64_t *__restrict__ bin_offsets, int n_members,
                  const uint32_t *__restrict__ chunk_base, const GzTable *__restrict__ T,
                  uint32_t *__restrict__ chunk_bits, uint32_t *__restrict__ chunk_crc)
{
    __shared__ uint32_t s_base[MAX_BINS_GZ + 1];
    __shared__ uint32_t s_crc[256];
    __shared__ uint8_t s_len[257];
    for (int i = threadIdx.x; i <= n_members; i += blockDim.x) s_base[i] = chunk_base[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_crc[i] = T->crc_tab[i];
    for (int i = threadIdx.x; i < 257; i += blockDim.x) s_len[i] = T->len[i];
    __syncthreads();
    const uint32_t n_chunks = s_base[n_members];
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += gridDim.x * blockDim.x) {
        const int m = gz_member_of(s_base, n_members, c);
        uint64_t lo, hi;
        gz_chunk_range(bin_offsets[m], bin_offsets[m + 1], c - s_base[m], lo, hi);
        uint32_t bits, crc;
        gz_chunk_measure(text, lo, hi, s_len, s_crc, bits, crc);
        chunk_bits[c] = bits;
        chunk_crc[c] = crc;
    }
}

// one warp per member: where its chunks' bits start (exclusive sums), its CRC, its size
__global__ void __launch_bounds__(128)
gz_member_kernel(const uint64_t *__restrict__ bin_offsets, int n_members, const uint32_t *__restrict__ chunk_base,
                 const GzTable *__restrict__ T, const uint32_t *__restrict__ chunk_bits, const uint32_t *__restrict__ chunk_crc,
                 uint64_t *__restrict__ chunk_bitoff, uint64_t *__restrict__ member_bits, uint32_t *__restrict__ member_crc,
                 uint64_t *__restrict__ member_bytes)
{
    const int lane = threadIdx.x & 31;
    const int m = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (m >= n_members) return;
    const uint64_t start = bin_offsets[m], end = bin_offsets[m + 1];
    const uint32_t c0 = chunk_base[m], n = chunk_base[m + 1] - c0;
    if (n == 0) {
        if (lane == 0) { member_bits[m] = 0; member_crc[m] = 0; member_bytes[m] = 0; }
        return;
    }
    // lane l takes the chunks [l * per, (l + 1) * per): sums and CRC of its run, then the runs in order
    const uint32_t per = (n + 31u) / 32u;
    const uint32_t k0 = min(n, (uint32_t)lane * per), k1 = min(n, k0 + per);
    unsigned long long sum = 0, bytes = 0;
    uint32_t crc = 0;
    for (uint32_t k = k0; k < k1; k++) {
        uint64_t lo, hi;
        gz_chunk_range(start, end, k, lo, hi);
        sum += chunk_bits[c0 + k];
        crc = gz_crc_shift(T->crc_pow, crc, hi - lo) ^ chunk_crc[c0 + k];
        bytes += hi - lo;
    }
    unsigned long long before = 0, total = 0;
    uint32_t crc_all = 0;
    for (int l = 0; l < 32; l++) {
        const unsigned long long s_l = __shfl_sync(0xffffffffu, sum, l);
        const unsigned long long b_l = __shfl_sync(0xffffffffu, bytes, l);
        const uint32_t c_l = __shfl_sync(0xffffffffu, crc, l);
        if (l < lane) before += s_l;
        total += s_l;
        if (lane == 0 && b_l) crc_all = gz_crc_shift(T->crc_pow, crc_all, b_l) ^ c_l;
    }
    unsigned long long run = before;
    for (uint32_t k = k0; k < k1; k++) {
        chunk_bitoff[c0 + k] = run;
        run += chunk_bits[c0 + k];
    }
    if (lane == 0) { member_bits[m] = total; member_crc[m] = crc_all; member_bytes[m] = gz_member_bytes(*T, total); }
}

// one thread: where the members go (gz_offsets[n_members] = all of them)
__global__ void gz_offsets_kernel(int n_members, const uint64_t *__restrict__ member_bytes, uint64_t *__restrict__ gz_offsets)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long acc = 0;
    for (int m = 0; m < n_members; m++) { gz_offsets[m] = acc; acc += member_bytes[m]; }
    gz_offsets[n_members] = acc;
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
                 const uint64_t *__restrict__ chunk_bitoff, const uint64_t *__restrict__ member_bits,
                 const uint32_t *__restrict__ member_crc, const uint64_t *__restrict__ member_bytes,
                 const uint64_t *__restrict__ gz_offsets, uint64_t cap_bytes, uint8_t *__restrict__ out)
{
    __shared__ uint32_t s_base[MAX_BINS_GZ + 1];
    __shared__ uint16_t s_code[257];
    __shared__ uint8_t s_len[257];
    for (int i = threadIdx.x; i <= n_members; i += blockDim.x) s_base[i] = chunk_base[i];
    for (int i = threadIdx.x; i < 257; i += blockDim.x) { s_code[i] = T->code[i]; s_len[i] = T->len[i]; }
    __syncthreads();
    if (gz_offsets[n_members] > cap_bytes) return;         // orc_wait() reports it
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < (uint32_t)n_members && member_bytes[gid])
        gz_member_frame(*T, out, gz_offsets[gid], member_bytes[gid], member_bits[gid], member_crc[gid],
                        bin_offsets[gid + 1] - bin_offsets[gid]);
    const uint32_t n_chunks = s_base[n_members];
    for (uint32_t c = gid; c < n_chunks; c += gridDim.x * blockDim.x) {
        const int m = gz_member_of(s_base, n_members, c);
        uint64_t lo, hi;
        gz_chunk_range(bin_offsets[m], bin_offsets[m + 1], c - s_base[m], lo, hi);
        const uint64_t bit0 = 8u * (gz_offsets[m] + GZ_HEADER_BYTES) + T->hdr_nbits + chunk_bitoff[c];
        gz_chunk_encode(text, lo, hi, s_code, s_len, bit0, reinterpret_cast<uint32_t *>(out));
    }
}
#endif  // __CUDACC__

}  // namespace orc
