// tests/gzsim.cpp -- the device-side gzip encoder (csrc/orc_gz.cuh) run on the CPU: the same ORC_HD functions,
// with loops where the kernels have threads (one iteration per chunk / per lane / per member), so that the
// members can be handed to zlib without a GPU (tests/test_gz.py).
#include <stdint.h>
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
    for (uint64_t p = 0; p < bin_offsets[n_bins]; p++) hist[text[p]]++;              // gz_hist_kernel
    gz_build_table(hist.data(), T, K);                                                // gz_table_kernel
    if (code_lengths) memcpy(code_lengths, T.len, 257);
    std::vector<uint32_t> chunk_base(n_bins + 1, 0);
    for (int m = 0; m < n_bins; m++) chunk_base[m + 1] = chunk_base[m] + gz_member_chunks(bin_offsets[m], bin_offsets[m + 1]);
    const uint32_t n_chunks = chunk_base[n_bins];
    std::vector<uint32_t> chunk_bits(n_chunks), chunk_crc(n_chunks);
    std::vector<uint64_t> chunk_bitoff(n_chunks);
    auto member_of = [&](uint32_t c) { int m = 0; while (!(chunk_base[m] <= c && c < chunk_base[m + 1])) m++; return m; };
    for (uint32_t c = 0; c < n_chunks; c++) {                                         // gz_measure_kernel
        const int m = member_of(c);
        uint64_t lo, hi;
        gz_chunk_range(bin_offsets[m], bin_offsets[m + 1], c - chunk_base[m], lo, hi);
        gz_chunk_measure(text, lo, hi, T.len, T.crc_tab, chunk_bits[c], chunk_crc[c]);
    }
    std::vector<uint64_t> member_bits(n_bins, 0), member_bytes(n_bins, 0);
    std::vector<uint32_t> member_crc(n_bins, 0);
    for (int m = 0; m < n_bins; m++) {                                                // gz_member_kernel, lanes as a loop
        const uint64_t start = bin_offsets[m], end = bin_offsets[m + 1];
        const uint32_t c0 = chunk_base[m], n = chunk_base[m + 1] - c0;
        if (n == 0) continue;
        const uint32_t per = (n + 31u) / 32u;
        unsigned long long sum[32], bytes[32];
        uint32_t crc[32];
        for (uint32_t lane = 0; lane < 32; lane++) {
            const uint32_t k0 = n < lane * per ? n : lane * per, k1 = n < k0 + per ? n : k0 + per;
            sum[lane] = 0; bytes[lane] = 0; crc[lane] = 0;
            for (uint32_t k = k0; k < k1; k++) {
                uint64_t lo, hi;
                gz_chunk_range(start, end, k, lo, hi);
                sum[lane] += chunk_bits[c0 + k];
                crc[lane] = gz_crc_shift(T.crc_pow, crc[lane], hi - lo) ^ chunk_crc[c0 + k];
                bytes[lane] += hi - lo;
            }
        }
        unsigned long long total = 0;
        uint32_t crc_all = 0;
        for (uint32_t lane = 0; lane < 32; lane++) {
            const uint32_t k0 = n < lane * per ? n : lane * per, k1 = n < k0 + per ? n : k0 + per;
            unsigned long long run = total;
            for (uint32_t k = k0; k < k1; k++) { chunk_bitoff[c0 + k] = run; run += chunk_bits[c0 + k]; }
            total += sum[lane];
            if (bytes[lane]) crc_all = gz_crc_shift(T.crc_pow, crc_all, bytes[lane]) ^ crc[lane];
        }
        member_bits[m] = total; member_crc[m] = crc_all; member_bytes[m] = gz_member_bytes(T, total);
    }
    uint64_t acc = 0;
    for (int m = 0; m < n_bins; m++) { gz_offsets[m] = acc; acc += member_bytes[m]; }  // gz_offsets_kernel
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
        const uint64_t bit0 = 8u * (gz_offsets[m] + GZ_HEADER_BYTES) + T.hdr_nbits + chunk_bitoff[c];
        gz_chunk_encode(text, lo, hi, T.code, T.len, bit0, reinterpret_cast<uint32_t *>(out));
    }
    return (long long)acc;
}
