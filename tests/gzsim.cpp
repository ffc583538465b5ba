// tests/gzsim.cpp -- the device-side gzip encoder (csrc/orc_gz.cuh) run on the CPU: the same ORC_HD functions,
// with loops where the kernels have threads (one iteration per chunk / per lane / per member), so that the
// members can be handed to zlib without a GPU (tests/test_gz.py).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../nanopore-barcoding-orc_b200/csrc/orc_gz.cuh"

using namespace orc;

// text: bin-major FASTQ text, bin m = [bin_offsets[m], bin_offsets[m + 1]).  out: zeroed by the callee, cap
// bytes (a multiple of 4).  gz_offsets[n_bins + 1].  Returns the total size, or -1 if it exceeds cap.
extern "C" long long gzsim_compress(const uint8_t *text, const uint64_t *bin_offsets, int n_bins, uint8_t *out,
                                    uint64_t cap, uint64_t *gz_offsets, uint8_t *code_lengths /* [257] or NULL */)
{
    static GzTable T;
    static GzWork K;
    memset(&T, 0, sizeof T);
    gz_fill_crc_tables(T);
    std::vector<unsigned long long> hist(256, 0);
    {                                                                                 // gz_hist_kernel
        const uint64_t total = bin_offsets[n_bins], n16 = total >> 4;
        for (uint64_t i = 0; i < gz_n_samples(n16); i++)
            for (int k = 0; k < 16; k++) hist[text[16 * gz_sample_at(i) + k]]++;
        for (uint64_t p = n16 << 4; p < total; p++) hist[text[p]]++;
    }
    if (getenv("GZSIM_FLAT")) {                                                       // gz_table_kernel, flat
        for (int t = 0; t < 257; t++) K.w[t] = 1;
        for (int t = 0; t < 257; t++) K.order[gz_rank(K.w, t)] = t;
        gz_build_from_sorted(T, K, 0, 1, GzNoSync());
    } else
    gz_build_table(hist.data(), T, K);                                                // gz_table_kernel
    if (code_lengths) memcpy(code_lengths, T.len, 257);
    std::vector<uint32_t> chunk_base(n_bins + 1, 0);
    for (int m = 0; m < n_bins; m++) chunk_base[m + 1] = chunk_base[m] + gz_member_chunks(bin_offsets[m], bin_offsets[m + 1]);
    const uint32_t n_chunks = chunk_base[n_bins];
    const uint32_t n_tiles = (n_chunks + GZ_TILE - 1) / GZ_TILE;
    std::vector<uint32_t> chunk_local(n_chunks), tile_bits(n_tiles, 0), member_crc(n_bins, 0);
    auto member_of = [&](uint32_t c) { int m = 0; while (!(chunk_base[m] <= c && c < chunk_base[m + 1])) m++; return m; };
    for (uint32_t c = 0; c < n_chunks; c++) {                                         // gz_measure_kernel
        const int m = member_of(c);
        uint64_t lo, hi;
        gz_chunk_range(bin_offsets[m], bin_offsets[m + 1], c - chunk_base[m], lo, hi);
        uint32_t bits, crc;
        gz_chunk_measure(text, lo, hi, T.len, T.crc_tab, bits, crc);
        member_crc[m] ^= gz_crc_shift(T.crc_pow, crc, bin_offsets[m + 1] - hi);
        chunk_local[c] = tile_bits[c / GZ_TILE];
        tile_bits[c / GZ_TILE] += bits;
    }
    std::vector<uint64_t> tile_off(n_tiles + 1, 0), member_pos(n_bins + 1), member_bits(n_bins, 0), member_bytes(n_bins, 0);
    for (uint32_t t = 0; t < n_tiles; t++) tile_off[t + 1] = tile_off[t] + tile_bits[t];   // gz_layout_kernel
    for (int m = 0; m <= n_bins; m++) {
        const uint32_t c = chunk_base[m];
        member_pos[m] = (m < n_bins && c < n_chunks) ? tile_off[c / GZ_TILE] + chunk_local[c] : tile_off[n_tiles];
    }
    uint64_t acc = 0;
    for (int m = 0; m < n_bins; m++) {
        member_bits[m] = member_pos[m + 1] - member_pos[m];
        member_bytes[m] = bin_offsets[m + 1] > bin_offsets[m] ? gz_member_bytes(T, member_bits[m]) : 0;
        gz_offsets[m] = acc;
        acc += member_bytes[m];
    }
    gz_offsets[n_bins] = acc;
    if (acc > cap) return -1;
    memset(out, 0, cap);
    for (int m = 0; m < n_bins; m++)                                                  // gz_encode_kernel
        if (member_bytes[m])
            gz_member_frame(T, out, gz_offsets[m], member_bytes[m], member_bits[m], member_crc[m],
                            bin_offsets[m + 1] - bin_offsets[m]);
    for (uint32_t c = 0; c < n_chunks; c++) {
        const int m = member_of(c);
        uint64_t lo, hi;
        gz_chunk_range(bin_offsets[m], bin_offsets[m + 1], c - chunk_base[m], lo, hi);
        const uint64_t bit0 = 8u * (gz_offsets[m] + GZ_HEADER_BYTES) + T.hdr_nbits + (tile_off[c / GZ_TILE] + chunk_local[c] - member_pos[m]);
        gz_chunk_encode(text, lo, hi, T.sym, bit0, reinterpret_cast<uint32_t *>(out));
    }
    return (long long)acc;
}
