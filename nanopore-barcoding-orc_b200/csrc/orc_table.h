// orc_table.h -- host-side construction of the per-round constant tables (RoundTable).
// Restates cutadapt 4.9 adapters.py SingleAdapter.__init__ (sequence normalisation,
// max_error_rate, min_overlap) and the fp64 acceptance threshold of _align.pyx
// (`cost <= length * max_error_rate`) as an integer table per aligned length.
#pragma once
#include <math.h>
#include <string.h>

#include <string>

#include "orc_core.cuh"

namespace orc {

inline int base_code(char c)
{
    switch (c) {
    case 'A': return 1;
    case 'C': return 2;
    case 'G': return 4;
    case 'T': return 8;
    default: return -1;
    }
}

// Adapter character -> 4-bit mask, cutadapt's _iupac_table() (X matches nothing, N everything).
inline int iupac_code(char c)
{
    switch (c) {
    case 'X': return 0;  case 'A': return 1;  case 'C': return 2;  case 'M': return 3;
    case 'G': return 4;  case 'R': return 5;  case 'S': return 6;  case 'V': return 7;
    case 'T': return 8;  case 'W': return 9;  case 'Y': return 10; case 'H': return 11;
    case 'K': return 12; case 'D': return 13; case 'B': return 14; case 'N': return 15;
    default: return -1;
    }
}

// Returns "" on success, else the reason the round is not supported.
// filter_mode: 0 = never use the shared-prefix trigger filter, 1 = when it pays (prefix >= 12 and
// > 2k), 2 = whenever it is valid (prefix > k; for tests).
inline std::string build_round_table(RoundTable &T, int n_adapters, int type, const char *const *sequences,
                                     double max_errors, int min_overlap, int indels, int revcomp,
                                     int filter_mode = 1)
{
    memset(&T, 0, sizeof(T));
    if (n_adapters < 1 || n_adapters > MAX_AD)
        return "unsupported: between 1 and " + std::to_string(MAX_AD) + " adapters per round";
    if (type != TYPE_FRONT && type != TYPE_BACK)
        return "unsupported: only regular 5' (-g) and 3' (-a) adapters take the edit-distance path";
    if (min_overlap < 1) return "min_overlap must be >= 1";
    T.n_adapters = n_adapters;
    T.type = type;
    T.revcomp = revcomp ? 1 : 0;
    T.indels = indels ? 1 : 0;
    T.min_overlap = min_overlap;
    T.n_lanes = 2 * n_adapters;
    int n_wild = 0;
    for (int a = 0; a < n_adapters; a++) {
        const char *s = sequences[a];
        const int m = (int)strlen(s);
        if (m < 1 || m > MAX_M) return "unsupported: adapter length must be 1..64";
        T.m[a] = m;
        // n_counts[i] = number of N in adapter[0 : i] (_align.pyx _set_reference); an adapter with any
        // character outside ACGT is compared through the IUPAC masks (adapters.py adapter_wildcards)
        int n_counts[MAX_M + 1];
        bool wild = false;
        int nN = 0;
        for (int i = 0; i < m; i++) {
            char c = s[i];
            if (c >= 'a' && c <= 'z') c = (char)(c - 32);
            if (c == 'U') c = 'T';
            if (c == 'I') c = 'N';
            const int code = iupac_code(c);
            if (code < 0) return "unsupported: adapter character outside the IUPAC alphabet";
            T.code[a][i] = (uint8_t)code;
            n_counts[i] = nN;
            if (base_code(c) < 0) wild = true;
            if (c == 'N') nN++;
        }
        n_counts[m] = nN;
        if (wild) n_wild++;
        for (int q = 0; q < m; q++) {
            const int f = 16 + q;                   // code4: 16 pad nibbles, then the adapter
            T.code4[a][f >> 3] |= (uint32_t)T.code[a][q] << ((f & 7) * 4);
            T.rcode4[a][q >> 3] |= comp4(T.code[a][m - 1 - q]) << ((q & 7) * 4);
        }
        double rate = max_errors;
        if (rate >= 1.0) {                          // absolute error count (adapters.py)
            if (m - nN < 1) return "unsupported: an absolute error count for an adapter made of N only";
            rate /= (m - nN);
        }
        if (!(rate >= 0.0) || rate >= 1.0) return "unsupported: error rate must be in [0, 1) for every adapter";
        T.k[a] = (int)(rate * m);                   // _align.pyx: k = <int>(max_error_rate * m)
        T.min_ov[a] = min_overlap < m ? min_overlap : m;
        for (int L = 0; L <= m; L++) {
            // largest integer cost with (double)cost <= effective_length * rate.  _align.pyx takes the
            // N of adapter[0 : L] off an overlap of L < m characters in the last-row test (R5) and the N
            // of the aligned adapter part in the last-column test (R6).  A 3' adapter always aligns
            // from its first character, so the two agree; a 5' adapter reaches R6 with L < m only with
            // its last L characters (a read that lies inside the adapter).
            const int eff5 = L - n_counts[L];
            const int eff6 = (type == TYPE_BACK) ? eff5 : L - (n_counts[m] - n_counts[m - L]);
            int c5 = (int)floor(eff5 * rate), c6 = (int)floor(eff6 * rate);
            c5 = c5 < 0 ? 0 : (c5 > 255 ? 255 : c5);
            c6 = c6 < 0 ? 0 : (c6 > 255 ? 255 : c6);
            T.kmax[a][1][L] = (uint8_t)c5;
            T.kmax[a][2][L] = (uint8_t)c6;
            T.kmax[a][0][L] = (uint8_t)(c5 > c6 ? c5 : c6);
        }
    }
    // cutadapt decides per adapter: plain ones are compared as ASCII, those with wildcards through the IUPAC masks.
    // A plain adapter is its own mask set with no N, so the masks serve a mixed round as well -- up to a read with
    // U, which equals T only through the masks (T.mixed: orc_api.cu looks for U in the batch and refuses it)
    T.wild = n_wild ? 1 : 0;
    T.mixed = (n_wild != 0 && n_wild != n_adapters) ? 1 : 0;
    for (int lane = 0; lane < T.n_lanes; lane++) {
        const int a = lane % n_adapters, dir = lane / n_adapters;
        const int m = T.m[a];
        const uint64_t pad = (m == 64) ? 0ull : ((1ull << (64 - m)) - 1ull);
        for (uint32_t c = 0; c < 16; c++) {
            const uint32_t cc = dir ? comp4(c) : c;
            uint64_t bits = pad;
            for (int i = 0; i < m; i++)
                if (T.code[a][i] & cc) bits |= 1ull << (64 - m + i);
            T.peq[lane >> 5][c][lane & 31] = bits;
        }
        // R2: FRONT column 0 costs are all 0; BACK column 0 cost is i
        T.pv0[lane] = (type == TYPE_FRONT) ? 0ull : ~pad;
        T.d0[lane] = (type == TYPE_FRONT) ? 0 : m;
        // stage 2a block: the last Lb rows of a 5' adapter, the first Lb rows of a 3' adapter
        const int Lb = m < 32 ? m : 32;
        const int first = (type == TYPE_FRONT) ? m - Lb : 0;
        const uint32_t pad32 = (Lb == 32) ? 0u : ((1u << (32 - Lb)) - 1u);
        for (uint32_t c = 0; c < 16; c++) {
            const uint32_t cc = dir ? comp4(c) : c;
            uint32_t bits = pad32;
            for (int i = 0; i < Lb; i++)
                if (T.code[a][first + i] & cc) bits |= 1u << (32 - Lb + i);
            T.peq32b[c][lane] = bits;
        }
        // with k >= Lb a path may cross the block for free; with k_max >= 32 the first-column limits have no room
        T.block_len[a] = (indels && T.k[a] < Lb && T.k[a] < 16) ? Lb : 0;
    }
    // longest common prefix of the adapters (as code masks), capped at one 32-bit word
    int lcp = T.m[0];
    for (int a = 1; a < n_adapters; a++) {
        int l = 0;
        while (l < lcp && l < T.m[a] && T.code[a][l] == T.code[0][l]) l++;
        lcp = l;
    }
    if (lcp > 32) lcp = 32;
    int k_max = 0, m_max = 0;
    for (int a = 0; a < n_adapters; a++) {
        if (T.k[a] > k_max) k_max = T.k[a];
        if (T.m[a] > m_max) m_max = T.m[a];
    }
    T.lcp = lcp; T.k_max = k_max; T.m_max = m_max;
    T.m_min = T.m[0];
    for (int a = 1; a < n_adapters; a++) if (T.m[a] < T.m_min) T.m_min = T.m[a];
    const bool valid = lcp > k_max && lcp >= 1;
    T.use_filter = (filter_mode == 2) ? valid : (filter_mode == 1 ? (valid && lcp >= 12 && lcp > 2 * k_max) : 0);
    int lcs_ = T.m[0];
    for (int a = 1; a < n_adapters; a++) {
        int l = 0;
        while (l < lcs_ && l < T.m[a] && T.code[a][T.m[a] - 1 - l] == T.code[0][T.m[0] - 1 - l]) l++;
        lcs_ = l;
    }
    if (lcs_ > 32) lcs_ = 32;
    // a 3' round scans its shared suffix instead when that is the longer (more selective) flank
    T.sfx_primary = (T.use_filter && type == TYPE_BACK && lcs_ > lcp && lcs_ > k_max) ? 1 : 0;
    // shared suffix and the loosest acceptance limits (decide the mandatory windows in stage 1)
    int lcs = T.m[0];
    for (int a = 1; a < n_adapters; a++) {
        int l = 0;
        while (l < lcs && l < T.m[a] && T.code[a][T.m[a] - 1 - l] == T.code[0][T.m[0] - 1 - l]) l++;
        lcs = l;
    }
    if (lcs > 32) lcs = 32;
    T.lcs = lcs;
    T.min_ov_min = T.min_ov[0];
    for (int a = 1; a < n_adapters; a++) if (T.min_ov[a] < T.min_ov_min) T.min_ov_min = T.min_ov[a];
    for (int L = 0; L <= MAX_M; L++) {
        int v = 0;
        for (int a = 0; a < n_adapters; a++) {
            const int l = L < T.m[a] ? L : T.m[a];
            if (T.kmax[a][0][l] > v) v = T.kmax[a][0][l];
        }
        T.kmax_any[L] = (uint8_t)v;
    }
    for (int P = 0; P < 16; P++)
        for (int M = 0; M < 16; M++) {
            int sum = 0, low = 9;
            for (int u = 3; u >= 0; u--) {               // bit 3 is the oldest column
                sum += ((P >> u) & 1) - ((M >> u) & 1);
                if (sum < low) low = sum;
            }
            T.chunk_lut[P | (M << 4)] = (uint8_t)((low + 4) | ((sum + 4) << 4));
        }
    for (int j = 0; j < MAX_M + 32; j++) {
        int lim = -1;
        for (int c = 0; c <= T.k_max && c < 32; c++) {
            const int lmax = (j + c < T.m_max) ? j + c : T.m_max;
            if (lmax >= T.min_ov_min && c <= (int)T.kmax_any[lmax]) lim = c;
        }
        T.first_lim[j] = (int8_t)lim;
    }
    if (T.use_filter && lcs > 0) {
        const uint32_t pad32 = (lcs == 32) ? 0u : ((1u << (32 - lcs)) - 1u);
        for (int lane = 0; lane < 64; lane++) {
            const int dir = lane & 1;
            for (uint32_t c = 0; c < 16; c++) {
                const uint32_t cc = dir ? comp4(c) : c;
                uint32_t bits = pad32;
                for (int i = 0; i < lcs; i++)
                    if (T.code[0][T.m[0] - lcs + i] & cc) bits |= 1u << (32 - lcs + i);
                T.peq32s[c][lane] = bits;
            }
        }
    }
    if (T.use_filter) {
        const uint32_t pad32 = (lcp == 32) ? 0u : ((1u << (32 - lcp)) - 1u);
        for (int lane = 0; lane < 64; lane++) {
            const int dir = lane & 1;
            for (uint32_t c = 0; c < 16; c++) {
                const uint32_t cc = dir ? comp4(c) : c;
                uint32_t bits = pad32;
                for (int i = 0; i < lcp; i++)
                    if (T.code[0][i] & cc) bits |= 1u << (32 - lcp + i);
                T.peq32[c][lane] = bits;
            }
        }
    }
    return "";
}

// Stage 1s seed table (orc_core.cuh seed_scan): the leading floor(m / 8) eight-row pieces of every
// adapter, in both directions, behind a perfect hash.  Off (S.on == 0) when the round has no
// stage-1 filter, when some adapter has fewer pieces than errors + 1, or when no collision-free
// multiplier turns up.
inline void build_seed_table(const RoundTable &T, SeedTable &S, bool enable)
{
    memset(&S, 0, sizeof(S));
    for (int i = 0; i < SEED_SLOTS; i++) S.key[i] = SEED_EMPTY;
    if (!enable || !T.use_filter || T.wild) return;      // a piece with a wildcard equals no read 8-mer
    int need = 2;
    for (int a = 0; a < T.n_adapters; a++) {
        const int avail = T.m[a] / 8 - T.k[a];
        if (avail < 1) return;
        if (avail < need) need = avail;
    }
    struct Ent { uint32_t key, info; };
    Ent ents[2 * MAX_AD * (MAX_M / 8)];
    int n_ent = 0;
    for (int a = 0; a < T.n_adapters; a++)
        for (int t = 0; t < T.m[a] / 8; t++)
            for (int d = 0; d < 2; d++) {
                uint32_t key = 0;
                for (int u = 0; u < 8; u++) {
                    // direction 0: storage codes == adapter codes; direction 1: the storage holds the
                    // reverse complement of what the lane reads
                    const uint32_t c = d ? comp4(T.code[a][8 * t + 7 - u]) : (uint32_t)T.code[a][8 * t + u];
                    key |= c << (4 * u);
                }
                const uint32_t info = ((uint32_t)t << 16) | ((uint32_t)d << 20);
                // which adapters own a piece is never looked at (a superfluous window only costs time):
                // one entry per distinct (key, piece, direction)
                int f = -1;
                for (int i = 0; i < n_ent; i++) if (ents[i].key == key && ents[i].info == info) f = i;
                if (f < 0) { ents[n_ent].key = key; ents[n_ent].info = info; n_ent++; }
            }
    // group the entries by key
    uint32_t keys[2 * MAX_AD * (MAX_M / 8)];
    int n_keys = 0;
    for (int i = 0; i < n_ent; i++) {
        bool seen = false;
        for (int j = 0; j < n_keys; j++) if (keys[j] == ents[i].key) seen = true;
        if (!seen) keys[n_keys++] = ents[i].key;
    }
    if (n_ent > SEED_LIST_MAX) return;
    uint32_t mult = 0;
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    static int stamp[SEED_SLOTS];
    for (int attempt = 1; attempt <= 200000 && !mult; attempt++) {
        rng = rng * 6364136223846793005ull + 1442695040888963407ull;
        const uint32_t cand = (uint32_t)(rng >> 32) | 1u;
        bool ok = true;
        for (int j = 0; j < n_keys && ok; j++) {
            const uint32_t h = (keys[j] * cand) >> SEED_SHIFT;
            if (stamp[h] == attempt) ok = false;
            stamp[h] = attempt;
        }
        if (ok) mult = cand;
    }
    for (int i = 0; i < SEED_SLOTS; i++) stamp[i] = 0;
    if (!mult) return;
    int n_list = 0;
    for (int j = 0; j < n_keys; j++) {
        const uint32_t h = (keys[j] * mult) >> SEED_SHIFT;
        const int first = n_list;
        for (int i = 0; i < n_ent; i++) if (ents[i].key == keys[j]) S.list[n_list++] = ents[i].info;
        S.key[h] = keys[j];
        S.val[h] = (uint32_t)first | ((uint32_t)(n_list - first) << 16);
    }
    S.on = 1; S.need = need; S.mult = mult; S.n_list = n_list; S.kt = T.k_max; S.m_max = T.m_max;
}

// A 5' / 3' round with an adapter of more than MAX_M characters (LongTable, orc_core.cuh): the same quantities as
// build_round_table takes from cutadapt -- k = int(rate * m), min_overlap clamped to m, the fp64 acceptance limits
// per overlap length with the N counts of R5 / R6 -- for up to MAX_AD_LONG adapters of up to MAX_M_LONG characters.
// Also fills the few RoundTable fields the selection kernel reads.
inline bool round_is_long(int n_adapters, const char *const *sequences)
{
    for (int a = 0; a < n_adapters; a++)
        if (sequences[a] && strlen(sequences[a]) > (size_t)MAX_M) return true;
    return false;
}
inline std::string build_long_table(LongTable &L, RoundTable &T, int n_adapters, int type, const char *const *sequences,
                                    double max_errors, int min_overlap, int indels, int revcomp)
{
    memset(&L, 0, sizeof(L));
    memset(&T, 0, sizeof(T));
    if (n_adapters < 1 || n_adapters > MAX_AD_LONG)
        return "unsupported: between 1 and " + std::to_string(MAX_AD_LONG) + " adapters in a round with adapters over 64 nt";
    if (type != TYPE_FRONT && type != TYPE_BACK)
        return "unsupported: only regular 5' (-g) and 3' (-a) adapters take the edit-distance path";
    if (min_overlap < 1) return "min_overlap must be >= 1";
    L.n_adapters = n_adapters; L.type = type; L.revcomp = revcomp ? 1 : 0; L.indels = indels ? 1 : 0;
    int n_wild = 0;
    for (int a = 0; a < n_adapters; a++) {
        const char *s = sequences[a];
        const int m = (int)strlen(s);
        if (m < 1 || m > MAX_M_LONG) return "unsupported: adapter length must be 1.." + std::to_string(MAX_M_LONG);
        L.m[a] = m;
        std::vector<int> n_counts((size_t)m + 1, 0);
        bool wild = false;
        int nN = 0;
        for (int i = 0; i < m; i++) {
            char c = s[i];
            if (c >= 'a' && c <= 'z') c = (char)(c - 32);
            if (c == 'U') c = 'T';
            if (c == 'I') c = 'N';
            const int code = iupac_code(c);
            if (code < 0) return "unsupported: adapter character outside the IUPAC alphabet";
            L.code[a][i] = (uint8_t)code;
            n_counts[(size_t)i] = nN;
            if (base_code(c) < 0) wild = true;
            if (c == 'N') nN++;
        }
        n_counts[(size_t)m] = nN;
        if (wild) n_wild++;
        double rate = max_errors;
        if (rate >= 1.0) {                          // absolute error count (adapters.py)
            if (m - nN < 1) return "unsupported: an absolute error count for an adapter made of N only";
            rate /= (m - nN);
        }
        if (!(rate >= 0.0) || rate >= 1.0) return "unsupported: error rate must be in [0, 1) for every adapter";
        L.k[a] = (int)(rate * m);
        if (L.k[a] > 63) return "unsupported: more than 63 errors allowed in one adapter";       // pack_key()
        L.min_ov[a] = min_overlap < m ? min_overlap : m;
        for (int len = 0; len <= m; len++) {
            const int eff5 = len - n_counts[(size_t)len];
            const int eff6 = (type == TYPE_BACK) ? eff5 : len - (n_counts[(size_t)m] - n_counts[(size_t)(m - len)]);
            int c5 = (int)floor(eff5 * rate), c6 = (int)floor(eff6 * rate);
            L.kmax5[a][len] = (uint8_t)(c5 < 0 ? 0 : (c5 > 255 ? 255 : c5));
            L.kmax6[a][len] = (uint8_t)(c6 < 0 ? 0 : (c6 > 255 ? 255 : c6));
        }
    }
    T.n_adapters = n_adapters;
    T.mixed = (n_wild != 0 && n_wild != n_adapters) ? 1 : 0;
    T.type = type;
    T.revcomp = revcomp ? 1 : 0;
    T.indels = indels ? 1 : 0;
    T.min_overlap = min_overlap;
    T.wild = n_wild ? 1 : 0;
    T.n_lanes = 0;
    return "";
}

// One code array serves every round of a ctx: if any round compares through the IUPAC masks, all of them do (reads
// packed with U = T).  Returns whether a batch must be searched for U: some adapter of some unanchored round is plain
// ACGT and would, in cutadapt, not take a read's U for its T.  (Anchored rounds read the ASCII bases themselves.)
inline bool unify_wildcards(RoundTable *T, const bool *bit_parallel_or_long, int n_rounds)
{
    bool any_wild = false, any_plain = false;
    for (int r = 0; r < n_rounds; r++) {
        if (!bit_parallel_or_long[r]) continue;
        any_wild = any_wild || T[r].wild != 0;
        any_plain = any_plain || T[r].wild == 0 || T[r].mixed != 0;
    }
    if (!any_wild) return false;
    for (int r = 0; r < n_rounds; r++)
        if (bit_parallel_or_long[r]) T[r].wild = 1;
    return any_plain;
}

// Anchored, no-indel round (ORC_PREFIX / ORC_SUFFIX).  Also fills the few RoundTable fields the
// selection kernel reads (type as trimming side, revcomp, n_adapters).
inline std::string build_anchored_table(AnchoredTable &A, RoundTable &T, int n_adapters, int suffix,
                                        const char *const *sequences, double max_errors, int indels, int revcomp)
{
    memset(&A, 0, sizeof(A));
    memset(&T, 0, sizeof(T));
    if (n_adapters < 1 || n_adapters > MAX_ANCH)
        return "unsupported: between 1 and " + std::to_string(MAX_ANCH) + " anchored adapters per round";
    if (indels) return "unsupported: anchored adapters need --no-indels (the indel variant is not built)";
    A.n_adapters = n_adapters; A.suffix = suffix ? 1 : 0; A.revcomp = revcomp ? 1 : 0;
    bool one_length = true, small_k = true;
    for (int a = 0; a < n_adapters; a++) {
        const char *s = sequences[a];
        const int m = (int)strlen(s);
        if (m < 1 || m > MAX_M) return "unsupported: adapter length must be 1..64";
        A.m[a] = m;
        for (int i = 0; i < m; i++) {
            char c = s[i];
            if (c >= 'a' && c <= 'z') c = (char)(c - 32);
            if (c == 'U') c = 'T';
            if (base_code(c) < 0) return "unsupported: adapter characters other than ACGT (IUPAC wildcards)";
            A.seq[a][i] = (uint8_t)c;
            A.nib[a][i >> 4] |= (uint64_t)(c == 'A' ? 1u : c == 'C' ? 2u : c == 'G' ? 4u : 8u) << (4 * (i & 15));
        }
        double rate = max_errors;
        if (rate >= 1.0) rate /= m;
        if (!(rate >= 0.0) || rate >= 1.0) return "unsupported: error rate must be in [0, 1) for every adapter";
        A.k[a] = (int)(rate * m);
        if (m != A.m[0]) one_length = false;
        if (A.k[a] > 2) small_k = false;
    }
    if (n_adapters >= 2 && small_k && !one_length)
        return "unsupported: anchored adapters of several lengths (cutadapt's multi-length index)";
    A.indexed = (n_adapters >= 2 && small_k && one_length) ? 1 : 0;
    A.one_length = one_length ? 1 : 0;
    T.n_adapters = n_adapters;
    T.type = suffix ? TYPE_BACK : TYPE_FRONT;       // which side the selection trims
    T.revcomp = revcomp ? 1 : 0;
    T.n_lanes = 0;
    return "";
}

// ASCII -> 4-bit code used by the pack kernel.  cutadapt compares ASCII after upper()
// when the adapter is plain ACGT, so only ACGT/acgt get a code; anything
// else (N, U, IUPAC, '-') is 0 and matches nothing.
// Adapters with IUPAC wildcards are compared through _acgt_table() instead, which also knows U = T.
inline void build_pack_lut(uint8_t lut[256], bool wild = false)
{
    memset(lut, 0, 256);
    lut[(int)'A'] = lut[(int)'a'] = 1;
    lut[(int)'C'] = lut[(int)'c'] = 2;
    lut[(int)'G'] = lut[(int)'g'] = 4;
    lut[(int)'T'] = lut[(int)'t'] = 8;
    if (wild) lut[(int)'U'] = lut[(int)'u'] = 8;
}

// dnaio SequenceRecord.reverse_complement(): IUPAC complement, case preserved.
inline void build_complement_lut(uint8_t lut[256])
{
    for (int i = 0; i < 256; i++) lut[i] = (uint8_t)i;
    const char *from = "ACGTUMRWSYKVHDBN";
    const char *to = "TGCAAKYWSRMBDHVN";
    for (int i = 0; from[i]; i++) {
        lut[(int)from[i]] = (uint8_t)to[i];
        lut[(int)(from[i] | 0x20)] = (uint8_t)(to[i] | 0x20);
    }
}

}  // namespace orc
