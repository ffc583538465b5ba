// orc_pgz.h -- chunk-parallel inflate of a FOREIGN gzip stream (included by orc_io.cpp).
//
// The first input of the path (/root/reference/scripts/02_cutadapt_loop.sh:64-72 reads
// pychopped_<dataset>.fastq.gz) is a .gz somebody else wrote: one long DEFLATE stream, no size fields, so
// the member-parallel reader of orc_io.cpp has nothing to hop over and a single zlib inflate (about 0.2 GB/s
// of text) bounds files -> files at some 170 k reads/s while the GPU does 200 M.  This source cuts the
// COMPRESSED bytes into chunks and inflates them on a thread pool, the way pugz / rapidgzip do:
//
//   * a worker looks for the first DEFLATE block that starts in its chunk: a bit position is a candidate when
//     it reads as a non-final dynamic-Huffman block header with a complete code-length code, and it stands
//     when everything from there to the end of the chunk decodes without an error;
//   * what a block copies from the 32 KiB in front of the chunk is not known yet, so the worker decodes into
//     16-bit symbols: 0..255 a byte, 256 + j "byte j of the window in front of this chunk"; copies of copies
//     carry the markers along;
//   * the consumer takes the chunks in file order.  A chunk is accepted only if it starts at the very bit the
//     chunk before it ended on -- that chain, back to the gzip header, is what makes a guessed start a real
//     one; where it does not hold (fixed or stored blocks at the border, a block larger than a chunk, a wrong
//     guess, the seam between two members) the consumer decodes from the known position itself until the
//     chain closes again;
//   * an accepted piece goes back to the pool: its markers are replaced through a 33 KiB look-up table built
//     from the window the text in front left behind (the consumer resolves only a piece's last 32 KiB itself --
//     they are the next window), the bytes are written straight into the caller's buffer when the piece fits
//     the read() under way, and their CRC-32 is summed; the sums are combined in file order and compared, with
//     ISIZE, to every member's trailer.  Members that begin inside a chunk (cat a.gz b.gz) are decoded on by
//     the same worker from their own header, without markers.
//
// The block decoder below is this library's own (zlib has no way to start without a window); its acceptance
// rules are zlib's (inflate.c / inftrees.c: over-subscribed and incomplete codes, missing end-of-block code,
// distances too far back, stored-length complement), so what it accepts zlib accepts, byte for byte
// (tests/test_pgz.py: streams written bit by bit and zlib's output under every strategy; tests/test_io.py: through
// orc_reader -- gzip / zlib levels, several members, stored blocks, cut and damaged files).
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace orcpgz {

constexpr int WIN = 32768;
constexpr int LIT_PB = 11, DIST_PB = 9, PRE_PB = 7;
constexpr uint64_t NONE = ~0ull;

// table entry: bits 0-7 code length (primary entry that points to a subtable: number of subtable bits),
// 8-11 extra bits, 12-15 kind, 16-31 value (literal, base length / distance, subtable offset)
enum : uint32_t { K_LIT = 0, K_BASE = 1, K_EOB = 2, K_SUB = 3, K_BAD = 4 };
static inline uint32_t mk(uint32_t len, uint32_t extra, uint32_t kind, uint32_t val)
{
    return len | (extra << 8) | (kind << 12) | (val << 16);
}
static inline uint32_t kind_of(uint32_t e) { return (e >> 12) & 15u; }

static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

enum TableType { T_PRE, T_LIT, T_DIST };

struct Huff {
    std::vector<uint32_t> t;
    int pb = 0;

    static inline uint32_t rev(uint32_t code, int len)
    {
        uint32_t r = 0;
        for (int i = 0; i < len; i++) { r = (r << 1) | (code & 1u); code >>= 1; }
        return r;
    }
    static uint32_t entry_for(TableType type, int sym, int len)
    {
        if (type == T_PRE) return mk(len, 0, K_LIT, sym);
        if (type == T_LIT) {
            if (sym < 256) return mk(len, 0, K_LIT, sym);
            if (sym == 256) return mk(len, 0, K_EOB, 0);
            if (sym < 286) return mk(len, LEN_EXTRA[sym - 257], K_BASE, LEN_BASE[sym - 257]);
            return mk(len, 0, K_BAD, 0);
        }
        if (sym < 30) return mk(len, DIST_EXTRA[sym], K_BASE, DIST_BASE[sym]);
        return mk(len, 0, K_BAD, 0);
    }

    // inftrees.c's rules: an over-subscribed set is an error; an incomplete one too, unless it is a
    // literal/length or distance set whose longest code has one bit; no code at all gives a table of bad entries
    bool build(const uint8_t *lens, int n, TableType type, int primary_bits)
    {
        int count[16] = {0};
        for (int i = 0; i < n; i++) count[lens[i]]++;
        count[0] = 0;
        int maxlen = 0;
        for (int l = 15; l >= 1; l--) if (count[l]) { maxlen = l; break; }
        int left = 1;
        for (int l = 1; l <= 15; l++) {
            left <<= 1;
            left -= count[l];
            if (left < 0) return false;
        }
        if (maxlen && left > 0 && (type == T_PRE || maxlen != 1)) return false;
        pb = primary_bits;
        const uint32_t bad = mk(1, 0, K_BAD, 0);
        const size_t primary = (size_t)1 << pb;
        t.assign(primary, bad);
        if (!maxlen) return true;
        uint32_t next[16];
        uint32_t code = 0;
        for (int l = 1; l <= 15; l++) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
        if (maxlen > pb) {
            // longest code behind every primary index that long codes share
            std::vector<uint8_t> deepest(primary, 0);
            uint32_t nx[16];
            memcpy(nx, next, sizeof nx);
            for (int s = 0; s < n; s++) {
                const int l = lens[s];
                if (l <= pb) { if (l) nx[l]++; continue; }
                const uint32_t r = rev(nx[l]++, l) & (uint32_t)(primary - 1);
                if (deepest[r] < l) deepest[r] = (uint8_t)l;
            }
            for (size_t i = 0; i < primary; i++)
                if (deepest[i]) {
                    const int sb = deepest[i] - pb;
                    t[i] = mk((uint32_t)sb, 0, K_SUB, (uint32_t)t.size());
                    if (t.size() + ((size_t)1 << sb) > 65535 + primary) return false;      // (cannot happen: <= 2^15 codes)
                    t.resize(t.size() + ((size_t)1 << sb), bad);
                }
        }
        for (int s = 0; s < n; s++) {
            const int l = lens[s];
            if (!l) continue;
            const uint32_t r = rev(next[l]++, l);
            const uint32_t e = entry_for(type, s, l);
            if (l <= pb) {
                for (size_t i = r; i < primary; i += (size_t)1 << l) t[i] = e;
            } else {
                const uint32_t pe = t[r & (primary - 1)];
                const size_t base = pe >> 16, size = (size_t)1 << (pe & 0xff);
                for (size_t i = r >> pb; i < size; i += (size_t)1 << (l - pb)) t[base + i] = e;
            }
        }
        return true;
    }
};

struct Fixed {
    Huff lit, dist;
    Fixed()
    {
        uint8_t l[288];
        for (int i = 0; i < 144; i++) l[i] = 8;
        for (int i = 144; i < 256; i++) l[i] = 9;
        for (int i = 256; i < 280; i++) l[i] = 7;
        for (int i = 280; i < 288; i++) l[i] = 8;
        lit.build(l, 288, T_LIT, LIT_PB);
        uint8_t d[32];
        for (int i = 0; i < 32; i++) d[i] = 5;
        dist.build(d, 32, T_DIST, DIST_PB);
    }
};
static inline const Fixed &fixed_tables()
{
    static const Fixed f;
    return f;
}

struct RunResult {
    uint64_t end_bit = 0;
    bool final = false;
    const char *err = nullptr;
};

// Decodes DEFLATE blocks into 16-bit symbols.  out[0 .. WIN) stands for the window in front of the start
// (markers 256 + j, or real bytes when the caller knows them); avail_before says how many of its entries,
// counted from the end, a distance may reach.
struct Decoder {
    const uint8_t *in = nullptr, *end = nullptr;
    const uint8_t *p = nullptr;
    uint64_t buf = 0;
    int cnt = 0;
    Huff pre, lit, dist;
    std::vector<uint16_t> out;
    size_t op = WIN;
    size_t max_out = (size_t)32 << 20;      // symbols after which run() stops at the next block border

    inline void refill()
    {
        if (p + 8 <= end) {
            uint64_t w;
            memcpy(&w, p, 8);
            buf |= w << cnt;
            p += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56 && p < end) { buf |= (uint64_t)*p++ << cnt; cnt += 8; }
        }
    }
    inline uint64_t bitpos() const { return 8ull * (uint64_t)(p - in) - (uint64_t)cnt; }
    inline void seek(uint64_t bit)
    {
        p = in + (bit >> 3);
        buf = 0;
        cnt = 0;
        refill();
        const int d = (int)(bit & 7);
        buf >>= d;
        cnt -= d;
    }
    inline uint32_t take(int n)
    {
        const uint32_t v = (uint32_t)(buf & ((1ull << n) - 1));
        buf >>= n;
        cnt -= n;
        return v;
    }
    void reserve(size_t more)
    {
        if (op + more > out.size()) out.resize(std::max(out.size() * 2, op + more + (1u << 20)));
    }

    const char *dynamic_header()
    {
        refill();
        const int nlen = (int)take(5) + 257, ndist = (int)take(5) + 1, ncode = (int)take(4) + 4;
        if (nlen > 286 || ndist > 30) return "too many length or distance symbols";
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint8_t cl[19] = {0};
        for (int i = 0; i < ncode; i++) {
            if (i == 0 || i == 10) refill();
            cl[order[i]] = (uint8_t)take(3);
        }
        if (cnt < 0) return "unexpected end of the stream";
        if (!pre.build(cl, 19, T_PRE, PRE_PB)) return "invalid code lengths set";
        uint8_t lens[286 + 30];
        int i = 0;
        while (i < nlen + ndist) {
            refill();
            const uint32_t e = pre.t[buf & ((1u << PRE_PB) - 1)];
            if (kind_of(e) == K_BAD) return "invalid code lengths set";
            buf >>= (e & 0xff);
            cnt -= (int)(e & 0xff);
            const int sym = (int)(e >> 16);
            if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
            int rep;
            uint8_t val = 0;
            if (sym == 16) {
                if (i == 0) return "invalid bit length repeat";
                val = lens[i - 1];
                rep = 3 + (int)take(2);
            } else if (sym == 17) rep = 3 + (int)take(3);
            else rep = 11 + (int)take(7);
            if (i + rep > nlen + ndist) return "invalid bit length repeat";
            while (rep--) lens[i++] = val;
        }
        if (cnt < 0) return "unexpected end of the stream";
        if (lens[256] == 0) return "invalid code -- missing end-of-block";
        if (!lit.build(lens, nlen, T_LIT, LIT_PB)) return "invalid literal/lengths set";
        if (!dist.build(lens + nlen, ndist, T_DIST, DIST_PB)) return "invalid distances set";
        return nullptr;
    }

    const char *symbols(const Huff &L, const Huff &D, size_t avail_before)
    {
        const uint32_t *lt = L.t.data(), *dt = D.t.data();
        uint16_t *o = out.data();
        size_t cap = out.size();
#define ORC_PGZ_LITLEN(e)                                                                              \
    e = lt[buf & ((1u << LIT_PB) - 1)];                                                                \
    if (kind_of(e) == K_SUB) e = lt[(e >> 16) + ((buf >> LIT_PB) & ((1u << (e & 0xff)) - 1))];
#define ORC_PGZ_EAT(e)                                                                                 \
    buf >>= (e & 0xff);                                                                                \
    cnt -= (int)(e & 0xff);
        for (;;) {
            if (op + 320 > cap) {
                reserve(320);
                o = out.data();
                cap = out.size();
            }
            refill();                       // 56 bits or more: three literals, or one length / distance pair
            uint32_t e;
            ORC_PGZ_LITLEN(e)
            if (kind_of(e) == K_LIT) {
                ORC_PGZ_EAT(e)
                o[op++] = (uint16_t)(e >> 16);
                ORC_PGZ_LITLEN(e)
                if (kind_of(e) == K_LIT) {
                    ORC_PGZ_EAT(e)
                    o[op++] = (uint16_t)(e >> 16);
                    ORC_PGZ_LITLEN(e)
                    if (kind_of(e) == K_LIT) {
                        ORC_PGZ_EAT(e)
                        o[op++] = (uint16_t)(e >> 16);
                        if (cnt < 0) return "unexpected end of the stream";
                        continue;
                    }
                }
                if (cnt < 0) return "unexpected end of the stream";
                refill();                   // (the bits of e are still in front)
            }
            ORC_PGZ_EAT(e)
            const uint32_t k = kind_of(e);
            if (k == K_EOB) return cnt < 0 ? "unexpected end of the stream" : nullptr;
            if (k != K_BASE) return "invalid literal/length code";
            const int xl = (int)((e >> 8) & 15u);
            const size_t len = (e >> 16) + take(xl);
            uint32_t d = dt[buf & ((1u << DIST_PB) - 1)];
            if (kind_of(d) == K_SUB) d = dt[(d >> 16) + ((buf >> DIST_PB) & ((1u << (d & 0xff)) - 1))];
            ORC_PGZ_EAT(d)
            if (kind_of(d) != K_BASE) return "invalid distance code";
            const int xd = (int)((d >> 8) & 15u);
            const size_t dd = (d >> 16) + take(xd);
            if (cnt < 0) return "unexpected end of the stream";
            if (dd > op - WIN + avail_before) return "invalid distance too far back";
            uint16_t *dst = o + op;
            const uint16_t *src = dst - dd;
            if (dd >= 8) {
                for (size_t i = 0; i < len; i += 8) memcpy(dst + i, src + i, 16);
            } else {
                for (size_t i = 0; i < len; i++) dst[i] = src[i];
            }
            op += len;
        }
#undef ORC_PGZ_LITLEN
#undef ORC_PGZ_EAT
    }

    // From start_bit on, block after block, until a block ends at or behind stop_at, exactly on `exact`, or
    // the final block is done.  The output so far is kept (op).
    RunResult run(uint64_t start_bit, uint64_t stop_at, uint64_t exact, size_t avail_before)
    {
        RunResult r;
        seek(start_bit);
        for (;;) {
            const uint64_t pos = bitpos();
            // (max_out: text that inflates a thousandfold must not become gigabytes of symbols in one piece;
            // whoever continues the chain takes up from end_bit)
            if (pos >= stop_at || pos == exact || op - WIN >= max_out) { r.end_bit = pos; return r; }
            refill();
            if (cnt < 3) { r.err = "unexpected end of the stream"; return r; }
            const uint32_t bfinal = take(1), type = take(2);
            const char *err = nullptr;
            if (type == 0) {
                take(cnt & 7);
                const uint8_t *q = in + (bitpos() >> 3);
                if (q + 4 > end) { r.err = "unexpected end of the stream"; return r; }
                const uint32_t len = q[0] | ((uint32_t)q[1] << 8), nlen = q[2] | ((uint32_t)q[3] << 8);
                if (len != (~nlen & 0xffffu)) { r.err = "invalid stored block lengths"; return r; }
                if (q + 4 + len > end) { r.err = "unexpected end of the stream"; return r; }
                reserve(len + 320);
                uint16_t *o = out.data() + op;
                for (uint32_t i = 0; i < len; i++) o[i] = q[4 + i];
                op += len;
                p = q + 4 + len;
                buf = 0;
                cnt = 0;
            } else if (type == 1) {
                err = symbols(fixed_tables().lit, fixed_tables().dist, avail_before);
            } else if (type == 2) {
                err = dynamic_header();
                if (!err) err = symbols(lit, dist, avail_before);
            } else err = "invalid block type";
            if (err) { r.err = err; return r; }
            if (bfinal) { r.end_bit = bitpos(); r.final = true; return r; }
        }
    }
};

// gzip member header at byte `at`: the bit the first block starts on, or NONE
static inline uint64_t parse_gzip_header(const uint8_t *in, size_t n, size_t at)
{
    if (at + 18 > n || in[at] != 0x1f || in[at + 1] != 0x8b || in[at + 2] != 8 || (in[at + 3] & 0xe0)) return NONE;
    const int flg = in[at + 3];
    size_t q = at + 10;
    if (flg & 4) {
        if (q + 2 > n) return NONE;
        q += 2 + ((size_t)in[q] | ((size_t)in[q + 1] << 8));
    }
    for (int bit : {8, 16})
        if (flg & bit) {
            while (q < n && in[q]) q++;
            q++;
        }
    if (flg & 2) q += 2;
    return q < n ? 8ull * q : NONE;
}

// vectors that go round between the workers and the consumer (a fresh 8 MiB vector costs its page faults)
template <class T>
struct BufPool {
    std::mutex mu;
    std::vector<std::vector<T>> free_list;
    std::vector<T> get()
    {
        std::lock_guard<std::mutex> lk(mu);
        if (free_list.empty()) return std::vector<T>();
        std::vector<T> v = std::move(free_list.back());
        free_list.pop_back();
        return v;
    }
    void put(std::vector<T> &&v)
    {
        std::lock_guard<std::mutex> lk(mu);
        if (free_list.size() < 64) free_list.push_back(std::move(v));
    }
};

struct Piece {                      // decoded text: WIN entries of prefix, then n symbols
    std::vector<uint16_t> sym;
    size_t n = 0;
    bool markers = false;
};

struct Seg {                        // blocks of ONE member, decoded by a worker
    uint64_t start = NONE, end_bit = 0;
    bool final = false;
    Piece piece;
};

struct Chunk {
    std::vector<Seg> segs;          // empty: no block start found; more than one: members end inside the chunk
};

// a piece the chain has accepted: the pool turns its symbols into bytes and sums them, the consumer hands
// the bytes out in order
struct Job {
    Piece piece;
    std::vector<uint8_t> lut;       // the window in front of the piece (only with markers)
    std::vector<uint8_t> bytes;     // the text, unless it goes straight to its place:
    uint8_t *dst = nullptr;         // where in the caller's buffer the piece belongs (it fits the current read())
    size_t n_bytes = 0;
    uint32_t crc = 0;
    bool done = false;
    bool ends_member = false;
    uint32_t want_crc = 0, want_n = 0;
};

struct Source {
    int fd = -1;
    const uint8_t *in = nullptr;
    size_t n = 0;
    uint64_t chunk_bits = 8ull << 20;
    size_t n_chunks = 0;
    std::vector<std::thread> pool;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::map<size_t, std::unique_ptr<Chunk>> results;
    std::deque<std::shared_ptr<Job>> resolve_q;
    size_t next_k = 0, consume_k = 0, max_inflight = 4;
    bool stop = false;
    uint64_t first_block = NONE;
    std::string err;
    BufPool<uint16_t> sym_pool;
    BufPool<uint8_t> byte_pool;

    // the chain (consumer thread)
    uint64_t P = NONE;                  // bit position of the next block header
    bool finished = false, failed = false;
    uint8_t window[WIN];
    uint64_t member_out = 0;
    Decoder seq;
    std::deque<Seg> carry;              // further segments of the chunk accepted last
    // delivery (consumer thread)
    std::deque<std::shared_ptr<Job>> pending;
    uint64_t pending_bytes = 0;         // text of the pending pieces that is not handed out yet
    uint8_t *rd_dst = nullptr;          // the read() call under way
    uint64_t rd_want = 0, rd_got = 0;
    size_t cur_pos = 0;
    uint32_t crc = 0;
    uint64_t delivered = 0;             // bytes of the current member handed out
    std::atomic<uint64_t> stat_parallel{0}, stat_serial{0};     // text bytes that came from the workers / from the consumer's own decoding (read by orc_reader_inflate_mode from another thread)
    double stat_wait_s = 0, stat_copy_s = 0, stat_serial_s = 0;     // consumer: waiting, copying out, own decoding
    static inline double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

    static bool probe(const char *path, size_t min_bytes)
    {
        const int f = open(path, O_RDONLY);
        if (f < 0) return false;
        struct stat st;
        uint8_t h[4];
        const bool ok = fstat(f, &st) == 0 && S_ISREG(st.st_mode) && (size_t)st.st_size >= std::max<size_t>(min_bytes, 64) &&
                        pread(f, h, 4, 0) == 4 && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8;
        close(f);
        return ok;
    }

    inline uint64_t threshold(size_t k) const { return std::min<uint64_t>((uint64_t)k * chunk_bits, 8ull * n); }

    static inline uint64_t peek(const uint8_t *in, uint64_t bit)
    {
        uint64_t w;
        memcpy(&w, in + (bit >> 3), 8);
        return w >> (bit & 7);
    }

    void prepare(Decoder &d)
    {
        d.in = in;
        d.end = in + n;
        if (d.out.empty()) d.out = sym_pool.get();
        if (d.out.size() < WIN + (4u << 20)) d.out.resize(WIN + (4u << 20));
        d.op = WIN;
    }
    void hand_over(Decoder &d, Piece &pc, bool markers)
    {
        pc.n = d.op - WIN;
        pc.markers = markers;
        if (d.op < d.out.size() / 4) {          // a short piece (a small member): copy it, keep the large buffer
            pc.sym.assign(d.out.begin(), d.out.begin() + (ptrdiff_t)d.op);
        } else {
            pc.sym.swap(d.out);                 // (its size stays the capacity mark; n says how much of it is text)
            d.out.clear();
        }
    }

    // after a member's final block at `end_bit`: the first block of the next member, if one follows in front of `hi`
    uint64_t next_member(uint64_t end_bit, uint64_t hi) const
    {
        const size_t q = (size_t)((end_bit + 7) >> 3);
        if (q + 8 > n) return NONE;
        const uint64_t next = parse_gzip_header(in, n, q + 8);
        return (next != NONE && next < hi) ? next : NONE;
    }

    void work(size_t k, Chunk &c, Decoder &d)
    {
        const uint64_t lo = threshold(k), hi = threshold(k + 1);
        prepare(d);
        uint64_t from = NONE;
        bool markers = false;
        if (k == 0) {
            if (first_block == NONE || first_block >= hi) return;
            from = first_block;
        } else {
            for (int j = 0; j < WIN; j++) d.out[j] = (uint16_t)(256 + j);
            const uint64_t last = n >= 32 ? std::min<uint64_t>(hi, 8ull * (n - 32)) : 0;
            for (uint64_t b = lo; b < last && from == NONE; b++) {
                const uint64_t v = peek(in, b);
                if ((v & 7u) != 4u) continue;                            // not final, dynamic Huffman codes
                if (((v >> 3) & 31u) > 29u || ((v >> 8) & 31u) > 29u) continue;
                const int ncode = (int)((v >> 13) & 15u) + 4;
                const uint64_t w = peek(in, b + 17);
                int kraft = 0;
                for (int i = 0; i < ncode; i++) {
                    const int l = (int)((w >> (3 * i)) & 7u);
                    if (l) kraft += 128 >> l;
                }
                if (kraft != 128) continue;                              // the code-length code must be complete
                d.op = WIN;
                const RunResult r = d.run(b, hi, NONE, WIN);
                if (r.err) continue;
                from = b;
                markers = true;
                c.segs.emplace_back();
                Seg &sg = c.segs.back();
                sg.start = b;
                sg.end_bit = r.end_bit;
                sg.final = r.final;
                hand_over(d, sg.piece, true);
            }
            if (from == NONE) return;
            from = c.segs.back().final ? next_member(c.segs.back().end_bit, hi) : NONE;
        }
        // members that begin inside the chunk: their window is empty, so no markers are needed
        (void)markers;
        while (from != NONE) {
            prepare(d);
            const RunResult r = d.run(from, hi, NONE, 0);
            if (r.err) return;          // the consumer meets the same error and reports it
            c.segs.emplace_back();
            Seg &sg = c.segs.back();
            sg.start = from;
            sg.end_bit = r.end_bit;
            sg.final = r.final;
            hand_over(d, sg.piece, false);
            from = r.final ? next_member(r.end_bit, hi) : NONE;
        }
    }

    void resolve(Job &j)
    {
        if (!j.dst) {
            j.bytes = byte_pool.get();
            j.bytes.resize(j.piece.n);
        }
        const uint16_t *s = j.piece.sym.data() + WIN;
        uint8_t *o = j.dst ? j.dst : j.bytes.data();
        const size_t m = j.piece.n;
        uint32_t c = 0;
        for (size_t at = 0; at < m; at += 1u << 16) {          // sum while the bytes are in the cache
            const size_t e = std::min<size_t>(m, at + (1u << 16));
            if (j.piece.markers) {
                const uint8_t *lut = j.lut.data();
                for (size_t i = at; i < e; i++) o[i] = lut[s[i]];
            } else {
                for (size_t i = at; i < e; i++) o[i] = (uint8_t)s[i];
            }
            c = (uint32_t)crc32(c, o + at, (uInt)(e - at));
        }
        j.crc = c;
        sym_pool.put(std::move(j.piece.sym));
        j.piece.sym = std::vector<uint16_t>();
        if (!j.lut.empty()) byte_pool.put(std::move(j.lut));
        j.lut = std::vector<uint8_t>();
    }

    void worker()
    {
        Decoder d;
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_work.wait(lk, [&] { return stop || !resolve_q.empty() || (next_k < n_chunks && next_k < consume_k + max_inflight); });
            if (stop) return;
            if (!resolve_q.empty()) {
                std::shared_ptr<Job> j = resolve_q.front();
                resolve_q.pop_front();
                lk.unlock();
                resolve(*j);
                lk.lock();
                j->done = true;
                cv_done.notify_all();
                continue;
            }
            const size_t k = next_k++;
            lk.unlock();
            std::unique_ptr<Chunk> c(new Chunk());
            work(k, *c, d);
            lk.lock();
            results[k] = std::move(c);
            cv_done.notify_all();
        }
    }

    bool start(const char *path, int threads, size_t chunk_bytes)
    {
        fd = open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size < 64) return false;
        n = (size_t)st.st_size;
        void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return false;
        in = (const uint8_t *)m;
        madvise(m, n, MADV_SEQUENTIAL);
        first_block = parse_gzip_header(in, n, 0);
        if (first_block == NONE) return false;
        chunk_bits = 8ull * std::max<size_t>(chunk_bytes, 4096);
        n_chunks = (size_t)((8ull * n + chunk_bits - 1) / chunk_bits);
        P = first_block;
        memset(window, 0, sizeof window);
        if (threads < 1) threads = 1;
        max_inflight = (size_t)threads * 2;
        for (int t = 0; t < threads; t++) pool.emplace_back([this] { worker(); });
        return true;
    }

    // The text of `pc` follows the text so far: the window moves on (only the last 32 KiB are resolved here),
    // the rest of the piece goes to the pool.  If the piece ends its member, the trailer is noted for the
    // moment the piece is handed out and the chain goes on behind the next member's header.
    bool adopt(Piece &&pc, bool final)
    {
        std::shared_ptr<Job> j(new Job());
        j->piece = std::move(pc);
        const Piece &p = j->piece;
        if (p.markers) {
            j->lut = byte_pool.get();
            j->lut.resize(256 + WIN);
            for (int i = 0; i < 256; i++) j->lut[i] = (uint8_t)i;
            memcpy(j->lut.data() + 256, window, WIN);
        }
        const size_t m = std::min<size_t>(p.n, WIN);
        if (m < WIN) memmove(window, window + m, WIN - m);
        const uint16_t *s = p.sym.data() + WIN + p.n - m;
        uint8_t *w = window + WIN - m;
        if (p.markers) for (size_t i = 0; i < m; i++) w[i] = j->lut[s[i]];
        else for (size_t i = 0; i < m; i++) w[i] = (uint8_t)s[i];
        member_out += p.n;
        if (final) {
            const size_t q = (size_t)((P + 7) >> 3);
            if (q + 8 > n) { err = "reading the input: truncated gzip stream"; return false; }
            j->ends_member = true;
            j->want_crc = in[q] | ((uint32_t)in[q + 1] << 8) | ((uint32_t)in[q + 2] << 16) | ((uint32_t)in[q + 3] << 24);
            j->want_n = in[q + 4] | ((uint32_t)in[q + 5] << 8) | ((uint32_t)in[q + 6] << 16) | ((uint32_t)in[q + 7] << 24);
            if (j->want_n != (uint32_t)member_out) { err = "reading the input: corrupt gzip stream (length differs)"; return false; }
            member_out = 0;
            const uint64_t next = parse_gzip_header(in, n, q + 8);     // anything but another member is ignored, as gzread does
            if (next == NONE) finished = true;
            else P = next;
        }
        // the pool writes the piece straight into the caller's buffer if all of it belongs to the current read()
        // (behind what is pending in front of it); otherwise into a buffer of its own, copied when its turn comes
        j->n_bytes = p.n;
        if (rd_dst && rd_got + pending_bytes + p.n <= rd_want) j->dst = rd_dst + rd_got + pending_bytes;
        pending_bytes += p.n;
        pending.push_back(j);
        {
            std::lock_guard<std::mutex> lk(mu);
            resolve_q.push_back(j);
            cv_work.notify_all();
        }
        return true;
    }

    // the consumer's own decoding from P (window known) until a block ends at or behind stop_at or on `exact`
    bool serial(uint64_t stop_at, uint64_t exact, RunResult &r)
    {
        const size_t avail = (size_t)std::min<uint64_t>(member_out, WIN);
        prepare(seq);
        for (int j = 0; j < WIN; j++) seq.out[j] = window[j];
        r = seq.run(P, stop_at, exact, avail);
        if (r.err) { err = std::string("reading the input: corrupt gzip stream (") + r.err + ")"; return false; }
        Piece pc;
        hand_over(seq, pc, false);
        stat_serial += pc.n;
        P = r.end_bit;
        return adopt(std::move(pc), r.final);
    }

    // one more piece of text joins the chain
    bool advance()
    {
        if (8ull * n <= P) { err = "reading the input: truncated gzip stream"; return false; }
        while (!carry.empty()) {
            if (carry.front().start != P) { carry.clear(); break; }
            Seg sg = std::move(carry.front());
            carry.pop_front();
            stat_parallel += sg.piece.n;
            P = sg.end_bit;
            return adopt(std::move(sg.piece), sg.final);
        }
        std::unique_ptr<Chunk> c;
        size_t k;
        uint64_t exact = NONE;
        {
            std::unique_lock<std::mutex> lk(mu);
            const double w0 = now();
            while (consume_k < n_chunks && threshold(consume_k + 1) <= P) {     // regions P is already behind
                cv_done.wait(lk, [&] { return results.count(consume_k) != 0; });
                results.erase(consume_k++);
                cv_work.notify_all();
            }
            k = consume_k;
            if (k < n_chunks) {
                cv_done.wait(lk, [&] { return results.count(k) != 0; });
                Chunk *got = results[k].get();
                const uint64_t start = got->segs.empty() ? NONE : got->segs[0].start;
                if (start == P) {
                    c = std::move(results[k]);
                    results.erase(k);
                    consume_k++;
                    cv_work.notify_all();
                } else if (start != NONE && start > P) {
                    exact = start;              // the consumer decodes up to it
                } else {
                    results.erase(k);           // no start found, or one in front of P: not a block of this stream
                    consume_k++;
                    cv_work.notify_all();
                }
            }
            stat_wait_s += now() - w0;
        }
        if (c) {
            for (size_t i = 1; i < c->segs.size(); i++) carry.push_back(std::move(c->segs[i]));
            Seg &sg = c->segs[0];
            stat_parallel += sg.piece.n;
            P = sg.end_bit;
            return adopt(std::move(sg.piece), sg.final);
        }
        RunResult r;
        const uint64_t stop_at = k < n_chunks ? threshold(k + 1) : 8ull * n;
        const double s0 = now();
        if (!serial(stop_at, exact, r)) return false;
        stat_serial_s += now() - s0;
        if (exact != NONE && (P > exact || finished)) {
            // the chain went past the worker's start: it was not a block of this stream (P short of it, e.g.
            // behind a member's last block, leaves the chunk in place for the next call)
            std::lock_guard<std::mutex> lk(mu);
            if (consume_k == k && results.count(k)) { results.erase(k); consume_k++; cv_work.notify_all(); }
        }
        return true;
    }

    // no worker may still be writing into the caller's buffer when read() gives up
    void drain_direct()
    {
        std::unique_lock<std::mutex> lk(mu);
        for (auto &j : pending)
            if (j->dst) cv_done.wait(lk, [&] { return j->done; });
    }

    // like gzread: up to `want` bytes, 0 at the end of the input, -1 on error (err set; the text in front of the
    // damage has been handed out by then)
    int64_t read(uint8_t *dst, uint64_t want)
    {
        rd_dst = dst;
        rd_want = want;
        rd_got = 0;
        int64_t rc = 0;
        while (rd_got < want) {
            const bool can_advance = !finished && !failed && pending.size() < max_inflight;    // the chain runs ahead of the delivery
            if (pending.empty()) {
                if (can_advance) {
                    if (!advance()) failed = true;
                    continue;
                }
                if (failed && !rd_got) rc = -1;
                break;
            }
            Job &j = *pending.front();
            {
                std::unique_lock<std::mutex> lk(mu);
                if (!j.done) {
                    if (can_advance) {
                        lk.unlock();
                        if (!advance()) failed = true;
                        continue;
                    }
                    const double w0 = now();
                    cv_done.wait(lk, [&] { return j.done; });
                    stat_wait_s += now() - w0;
                }
            }
            size_t m;
            if (j.dst) {
                m = j.n_bytes;              // already in its place
            } else {
                const double c0 = now();
                m = (size_t)std::min<uint64_t>(want - rd_got, j.n_bytes - cur_pos);
                memcpy(dst + rd_got, j.bytes.data() + cur_pos, m);
                stat_copy_s += now() - c0;
            }
            rd_got += m;
            cur_pos += m;
            pending_bytes -= m;
            if (cur_pos == j.n_bytes) {
                crc = (uint32_t)crc32_combine(crc, j.crc, (z_off_t)j.n_bytes);
                delivered += j.n_bytes;
                bool bad = false;
                if (j.ends_member) {
                    bad = crc != j.want_crc || (uint32_t)delivered != j.want_n;
                    crc = 0;
                    delivered = 0;
                }
                if (!j.dst) byte_pool.put(std::move(j.bytes));
                pending.pop_front();
                cur_pos = 0;
                if (bad) {
                    err = "reading the input: corrupt gzip stream (CRC-32 differs)";
                    failed = true;
                    drain_direct();
                    pending.clear();
                    pending_bytes = 0;
                    rc = -1;
                    break;
                }
            }
        }
        if (rc < 0) drain_direct();
        rd_dst = nullptr;
        return rc < 0 ? rc : (int64_t)rd_got;
    }

    ~Source()
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
            cv_work.notify_all();
        }
        for (std::thread &t : pool) t.join();
        if (in) munmap((void *)in, n);
        if (fd >= 0) close(fd);
    }
};

}  // namespace orcpgz
