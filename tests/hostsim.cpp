// hostsim.cpp -- TEST HARNESS (never part of the product).
//
// Runs the per-lane device arithmetic of nanopore-barcoding-orc_b200/csrc/orc_core.cuh on
// the CPU, lane by lane, in the order the kernels would, so that the exactness of the
// scan -> candidate hull -> banded resolve -> select chain can be checked against the
// oracle on machines without a GPU.  Built by tests/conftest.py with g++.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "orc_core.cuh"
#include "orc_edit.cuh"
#include "orc_table.h"

using namespace orc;

static uint64_t g_pairs[2], g_kept[2], g_seeded[2], g_band, g_wide;
// tasks resolved by the band resolver / the wide one
extern "C" void hostsim_resolved(uint64_t *out) { out[0] = g_band; out[1] = g_wide; }
extern "C" void hostsim_stats(uint64_t *out) { out[0] = g_pairs[0]; out[1] = g_kept[0]; out[2] = g_pairs[1]; out[3] = g_kept[1]; }
// reads whose stage 1 went through the seed table, per round
extern "C" void hostsim_seeded(uint64_t *out) { out[0] = g_seeded[0]; out[1] = g_seeded[1]; }
extern "C" int hostsim_demux(int n_rounds,
                             int n_ad0, int type0, const char *const *seq0, double e0, int ov0, int rc0,
                             int n_ad1, int type1, const char *const *seq1, double e1, int ov1, int rc1,
                             const uint8_t *seq, const uint64_t *offsets, const uint32_t *lengths,
                             uint32_t n_reads, uint64_t n_bytes,
                             Match *m0, Match *m1, uint64_t *out_lo, uint32_t *out_len, uint32_t *out_rc,
                             uint64_t *n_tasks, char *err, int err_len, int filter_mode_in, uint64_t *n_columns,
                             int indels)
{
    const int filter_mode = filter_mode_in & 3;
    RoundTable *T = new RoundTable[2];
    AnchoredTable *AT = new AnchoredTable[2];
    LongTable *LT = new LongTable[2];
    bool anch[2] = {type0 >= 2, n_rounds > 1 && type1 >= 2};
    // a round with an adapter over 64 nt takes cutadapt's recurrence as it is (long_kernel on the device)
    const bool longr[2] = {!anch[0] && round_is_long(n_ad0, seq0), n_rounds > 1 && !anch[1] && round_is_long(n_ad1, seq1)};
    std::string e = anch[0] ? build_anchored_table(AT[0], T[0], n_ad0, type0 == 3, seq0, e0, 0, rc0)
                  : longr[0] ? build_long_table(LT[0], T[0], n_ad0, type0, seq0, e0, ov0, indels, rc0)
                            : build_round_table(T[0], n_ad0, type0, seq0, e0, ov0, indels, rc0, filter_mode);
    if (e.empty() && n_rounds > 1)
        e = anch[1] ? build_anchored_table(AT[1], T[1], n_ad1, type1 == 3, seq1, e1, 0, rc1)
          : longr[1] ? build_long_table(LT[1], T[1], n_ad1, type1, seq1, e1, ov1, indels, rc1)
                    : build_round_table(T[1], n_ad1, type1, seq1, e1, ov1, indels, rc1, filter_mode);
    for (int rd = 0; rd < 2; rd++) anch[rd] = anch[rd] || longr[rd];     // below: "not the bit-parallel pipeline"
    // one code array for every round (orc_api.cu ctx_init): masks everywhere if anywhere, before the seed tables
    bool uses_codes[2] = {type0 < 2, n_rounds > 1 && type1 < 2};
    const bool need_u_check = e.empty() && unify_wildcards(T, uses_codes, n_rounds);
    bool wild_codes = false;
    for (int rd = 0; e.empty() && rd < n_rounds; rd++) wild_codes = wild_codes || (uses_codes[rd] && T[rd].wild != 0);
    // filter_mode bit 2 (value 4) keeps the flank scan of stage 1 although seeds would be usable
    SeedTable *ST = new SeedTable[2];
    for (int rd = 0; rd < n_rounds; rd++) {
        ST[rd].on = 0;
        if (e.empty() && !anch[rd]) build_seed_table(T[rd], ST[rd], (filter_mode_in & 4) == 0);
    }
    uint8_t comp_lut[256];
    build_complement_lut(comp_lut);
    if (!e.empty()) {
        strncpy(err, e.c_str(), (size_t)err_len - 1);
        err[err_len - 1] = 0;
        delete[] T;
        delete[] AT;
        delete[] LT;
        delete[] ST;
        return -1;
    }
    if (need_u_check) {
        // u_scan_kernel + orc_wait: plain and IUPAC adapters side by side, and a read with U
        for (uint32_t r = 0; r < n_reads; r++)
            for (uint32_t i = 0; i < lengths[r]; i++)
                if ((seq[offsets[r] + i] & 0xDFu) == 'U') {
                    strncpy(err, "unsupported: a read holds U, and the adapters mix plain ACGT sequences with IUPAC ones", (size_t)err_len - 1);
                    err[err_len - 1] = 0;
                    delete[] T;
                    delete[] AT;
                    delete[] LT;
                    delete[] ST;
                    return -1;
                }
    }
    uint8_t lut[256];
    build_pack_lut(lut, wild_codes);
    // flat pack with 4 guard words in front and 16 + 4 behind, like the device buffers
    const uint64_t n_words = (n_bytes + 7) / 8;
    std::vector<uint32_t> codes(n_words + 24, 0);
    uint32_t *W = codes.data() + 4;
    for (uint64_t i = 0; i < n_bytes; i++) W[i >> 3] |= (uint32_t)lut[seq[i]] << ((i & 7) * 4);

    ColRing *ring = new ColRing;
    BandEntry *band = new BandEntry[BAND_ENTRIES];
    uint32_t band_codes[BAND_CODE_WORDS];
    BandRing bring; bring.p = band; bring.stride = 1; bring.cw = band_codes; bring.w0 = 0;
    const bool force_wide = (filter_mode_in & 8) != 0;
    std::vector<uint32_t> band_tab[2];
    std::vector<BandAdapter> band_ads[2];
    for (int rd = 0; rd < n_rounds; rd++) {
        band_ads[rd].resize(MAX_AD);
        if (!anch[rd]) for (int a = 0; a < T[rd].n_adapters; a++) band_adapter_fill(T[rd], a, band_ads[rd][a]);
        band_tab[rd].assign((size_t)(MAX_LANES / 32) * BAND_BANK_BYTES / 4, 0u);
        if (!anch[rd])
            for (int bank = 0; bank < MAX_LANES / 32; bank++)
                for (int c = 0; c < 16; c++)
                    for (int l = 0; l < 32; l++)
                        band_table_entry(T[rd], bank, c, l, &band_tab[rd][((size_t)bank * 16 * 32 + (size_t)c * 32 + l) * 4]);
    }
    n_tasks[0] = n_tasks[1] = 0;
    n_columns[0] = n_columns[1] = 0;
    std::vector<LongCell> long_col(MAX_M_LONG + 1);
    for (uint32_t r = 0; r < n_reads; r++) {
        View v; v.lo = offsets[r]; v.len = lengths[r]; v.rc = 0;
        Match *out[2] = {&m0[r], &m1[r]};
        memset(&m1[r], 0, sizeof(Match));
        m1[r].adapter = -1;
        for (int rd = 0; rd < n_rounds; rd++) {
            const RoundTable &R = T[rd];
            std::vector<PairResult> results;
            uint64_t keys[2] = {0, 0};
            if (anch[rd]) {
                for (int o = 0; o < 2; o++) {
                    if (o == 1 && !R.revcomp) continue;
                    PairResult pr;
                    const int a = longr[rd] ? long_match(W, v, o, LT[rd], long_col.data(), pr)
                                            : anchored_match(seq, comp_lut, v, o, AT[rd], pr);
                    if (a >= 0) keys[o] = pack_key(pr.score, pr.errors, a, (uint32_t)results.size());
                    results.push_back(pr);
                }
                View next;
                select_read(R.type, R.revcomp, v, keys, results.data(), *out[rd], next);
                v = next;
                if (out[rd]->adapter < 0) break;
                continue;
            }
            WinList wl[2];
            // the read's segments as seed_kernel probes them: sw[segment][direction]
            SeedWins sw[SEED_SEGS_MAX][2];
            const int n_seg = seed_segments(v.len);
            if (R.use_filter && ST[rd].on) {
                for (int g = 0; g < n_seg; g++) {
                    uint32_t ra, rb;
                    seed_range(v.len, g, n_seg, ST[rd].m_max, ST[rd].kt, ra, rb);
                    seed_scan(W, v.lo, v.len, ST[rd].key, ST[rd].val, ST[rd].mult, ST[rd].list, ST[rd].need, ST[rd].kt,
                              ST[rd].m_max, sw[g], ra, rb);
                }
                g_seeded[rd]++;
            }
            if (R.use_filter) {
                for (int dir = 0; dir < 2; dir++) {
                    trigger_lane(W, v.lo, v.len, dir, (const char *)&R.peq32[0][0], dir, R.lcp, R.k_max, R.type,
                                 (uint32_t)(R.m_max - R.lcp + R.k_max), (uint32_t)(R.lcp + R.k_max + 1), wl[dir],
                                 R.lcs > 0 ? (const char *)&R.peq32s[0][0] : nullptr, R.lcs,
                                 R.kmax_any, R.min_ov_min, R.m_max, R.m_min, R.sfx_primary, R.first_lim, R.chunk_lut,
                                 ST[rd].on ? &sw[0][dir] : nullptr, n_seg, 2);
                    n_columns[rd] += win_columns(wl[dir]);
                }
            } else n_columns[rd] += 2ull * v.len;
            for (int lane = 0; lane < R.n_lanes; lane++) {
                const int a = lane % R.n_adapters, dir = lane / R.n_adapters;
                const int o = dir ^ (int)(v.rc & 1u);
                if (o == 1 && !R.revcomp) continue;
                g_pairs[rd]++;
                if (R.indels && !block_test(W, v.lo, v.len, dir, R.use_filter ? &wl[dir] : nullptr,
                                            (const char *)&R.peq32b[0][0], lane, R.block_len[a], R.k[a], R.type,
                                            R.kmax[a][0], R.min_ov[a], R.first_lim,
                                            R.use_filter ? R.m[a] - R.block_len[a] - R.k_max : 0)) continue;
                g_kept[rd]++;
                LaneScan L;
                scan_lane(W, v.lo, v.len, dir, R.use_filter ? &wl[dir] : nullptr, peq_bank(R, lane), lane,
                          R.pv0[lane], R.d0[lane], R.m[a], R.k[a], R.kmax[a][0], R.min_ov[a], R.type, L,
                          R.indels, R.code4[a], R.rcode4[a], R.chunk_lut);
                if (L.h.jf <= L.h.jl || L.h.i1 <= L.h.i2) {
                    PairResult pr; memset(&pr, 0, sizeof(pr));
                    if (!L.need) {
                        best_to_result(L.best, R.m[a], (int)v.len, pr);
                    } else {
                        Task t; t.read = r; t.lane = (uint32_t)lane;
                        t.jf = L.h.jf; t.jl = L.h.jl; t.i1 = L.h.i1; t.i2 = L.h.i2; t.slot = (uint32_t)results.size();
                        t.pad_ = task_anchors(L);
                        // what scan_kernel decides per task: the band resolver when the task fits it
                        if (!force_wide && task_band_ok(R.type, R.m[a], R.k[a], (int)v.len, t)) {
                            band_resolve_pair(W, v, R, band_ads[rd].data(), t, pr, bring, (const char *)band_tab[rd].data());
                            g_band++;
                        } else {
                            resolve_pair(W, v, R, t, pr, *ring);
                            g_wide++;
                        }
                        n_tasks[rd]++;
                    }
                    if (pr.has) {
                        const uint64_t key = pack_key(pr.score, pr.errors, a, (uint32_t)results.size());
                        if (key > keys[o]) keys[o] = key;
                    }
                    results.push_back(pr);
                }
            }
            View next;
            select_read(R.type, R.revcomp, v, keys, results.data(), *out[rd], next);
            v = next;
            if (out[rd]->adapter < 0) break;     // 02:75-80: "unknown" never enters round 2
        }
        out_lo[r] = v.lo; out_len[r] = v.len; out_rc[r] = v.rc;
    }
    delete ring;
    delete[] band;
    delete[] T;
    delete[] AT;
    delete[] LT;
    delete[] ST;
    return 0;
}

// the block formulation of orc_edit.cuh (what edit_kernel computes, without the lane skew)
extern "C" uint32_t hostsim_edit_distance(const uint8_t *q, uint32_t m, const uint8_t *t, uint32_t n, int mode)
{
    return orc::edit_distance_blocks(q, m, t, n, mode);
}
