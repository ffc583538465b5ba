// orc_io.cpp -- host-side FASTQ(.gz) streaming of liborcdemux.so: what dnaio's C parser and
// xopen's gzip pipes do under cutadapt (the reader -> workers -> ordered
// writer of cutadapt's ParallelPipelineRunner, /root/reference/scripts/02_cutadapt_loop.sh:64-72
// `-j 24`).
//
//   orc_reader   one thread fills a ring of page-locked text buffers with the inflated input and indexes
//                the records (orc_fastq_index); the text itself is the batch that goes to the GPU
//                (raw-text layout of orc_batch), so the host never copies a read.  The inflating is
//                done by a thread pool where the file allows it: member by member for files of
//                orc_writer (size fields), chunk by chunk for any other .gz (orc_pgz.h).
//   orc_writer   one output file per bin; the bin-major FASTQ text of a batch is cut into
//                chunks, a thread pool deflates them as independent gzip members and every file
//                receives its members in submission order (a concatenation of gzip members is a
//                valid .gz stream; the decompressed bytes are what the reference's tools read).
//
// No matching arithmetic lives here: bytes in, bytes out.
#include "../../include/orcdemux.h"

#include <cuda_runtime_api.h>
#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "orc_pgz.h"

namespace {

void set_err(char *err, size_t err_len, const std::string &s)
{
    if (err && err_len) snprintf(err, err_len, "%s", s.c_str());
}

void *host_alloc(size_t bytes, bool want_pinned, bool *pinned)
{
    void *p = nullptr;
    *pinned = false;
    if (want_pinned && cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess) {
        *pinned = true;
        return p;
    }
    cudaGetLastError();                 // page-locking is an optimisation of the copies only
    if (posix_memalign(&p, 4096, bytes ? bytes : 1) != 0) return nullptr;
    return p;
}

void host_free(void *p, bool pinned)
{
    if (!p) return;
    if (pinned) cudaFreeHost(p);
    else free(p);
}

struct ReaderBuf {
    uint8_t *text = nullptr;
    uint64_t *off = nullptr, *qoff = nullptr, *noff = nullptr;
    uint32_t *len = nullptr, *nlen = nullptr;
    bool pinned[6] = {false, false, false, false, false, false};
    uint64_t n_bytes = 0, bases = 0;
    uint32_t n_reads = 0;
};

// Every gzip member this library writes carries an FEXTRA subfield "OC" with its own total size (header +
// deflate stream + trailer) -- BGZF's idea with a 32-bit size -- so that a reader can find the members without
// inflating them and inflate them in parallel (orc_reader does; any other gzip reader skips the field).
constexpr size_t GZ_EXTRA_LEN = 8, GZ_SIZE_AT = 16;     // 10 header bytes, XLEN, then SI1 SI2 LEN(2) size(4)


// ------------------------------------------------------------------------------------ member-parallel inflate
// A .gz whose members carry this library's "OC" size field (everything orc_writer writes, e.g. the round-1
// bins that round 2 reads back, 02_cutadapt_loop.sh:94-102) or BGZF's "BC" field (bgzip) is inflated member by member on a thread
// pool: the headers are hopped over with the size field, the uncompressed size is the trailer's ISIZE, and
// the text comes out in file order.  A member without the field ends the parallel part; the rest of the
// file is then read through zlib's gzread from that offset on.
struct Member {
    uint64_t off = 0, clen = 0;
    uint32_t isize = 0;
    uint8_t *out = nullptr;
    bool done = false;
    std::string err;
};

struct MemberSource {
    int fd = -1;
    uint64_t file_size = 0, pos = 0;
    bool parallel_done = false;         // no further member with a size field
    gzFile tail = nullptr;              // serial reader of the rest
    std::vector<std::thread> pool;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::deque<Member *> order, jobs;
    size_t max_inflight = 8;
    bool stop = false;
    Member *cur = nullptr;
    uint64_t cur_pos = 0;
    std::string err;

    // the size of the member at `at` from its FEXTRA field, 0 if it says none: this library's "OC" subfield
    // (32-bit size) or BGZF's "BC" (bgzip, htslib: 16-bit size - 1; members of at most 64 KiB)
    static uint64_t member_size(int fd, uint64_t at, uint64_t file_size)
    {
        uint8_t h[12 + 64];
        if (at + 20 > file_size) return 0;
        const size_t have = (size_t)std::min<uint64_t>(sizeof h, file_size - at);
        if (pread(fd, h, have, (off_t)at) != (ssize_t)have) return 0;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return 0;
        const size_t xlen = (size_t)h[10] | ((size_t)h[11] << 8);
        if (12 + xlen > have) return 0;
        for (size_t q = 12; q + 4 <= 12 + xlen;) {
            const size_t len = (size_t)h[q + 2] | ((size_t)h[q + 3] << 8);
            if (q + 4 + len > 12 + xlen) return 0;
            uint64_t n = 0;
            if (h[q] == 'O' && h[q + 1] == 'C' && len == 4)
                n = (uint64_t)h[q + 4] | ((uint64_t)h[q + 5] << 8) | ((uint64_t)h[q + 6] << 16) | ((uint64_t)h[q + 7] << 24);
            else if (h[q] == 'B' && h[q + 1] == 'C' && len == 2)
                n = ((uint64_t)h[q + 4] | ((uint64_t)h[q + 5] << 8)) + 1;
            if (n) return (n >= 12 + xlen + 10 && at + n <= file_size) ? n : 0;
            q += 4 + len;
        }
        return 0;
    }

    static bool probe(const char *path)
    {
        int fd = open(path, O_RDONLY);
        if (fd < 0) return false;
        const off_t end = lseek(fd, 0, SEEK_END);
        const bool ok = end > 0 && member_size(fd, 0, (uint64_t)end) != 0;
        close(fd);
        return ok;
    }

    void worker()
    {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_job.wait(lk, [&] { return stop || !jobs.empty(); });
            if (stop) return;
            Member *m = jobs.front();
            jobs.pop_front();
            lk.unlock();
            std::vector<uint8_t> in(m->clen);
            std::string why;
            if (pread(fd, in.data(), m->clen, (off_t)m->off) != (ssize_t)m->clen) why = "short read of a gzip member";
            if (why.empty()) {
                m->out = (uint8_t *)malloc(m->isize ? m->isize : 1);
                z_stream zs;
                memset(&zs, 0, sizeof zs);
                if (!m->out || inflateInit2(&zs, 15 + 16) != Z_OK) why = "out of memory";
                else {
                    zs.next_in = in.data();
                    zs.avail_in = (uInt)m->clen;
                    zs.next_out = m->out;
                    zs.avail_out = m->isize;
                    const int rc = inflate(&zs, Z_FINISH);
                    if (rc != Z_STREAM_END || zs.total_out != m->isize) why = "corrupt gzip member";
                    inflateEnd(&zs);
                }
            }
            lk.lock();
            m->err = why;
            m->done = true;
            cv_done.notify_all();
        }
    }

    bool start(const char *path, int threads)
    {
        fd = open(path, O_RDONLY);
        if (fd < 0) return false;
        file_size = (uint64_t)lseek(fd, 0, SEEK_END);
        if (threads < 1) threads = 1;
        max_inflight = (size_t)threads * 2;
        if (member_size(fd, 0, file_size) <= (1u << 16)) max_inflight = (size_t)threads * 16;      // BGZF: small members
        for (int t = 0; t < threads; t++) pool.emplace_back([this] { worker(); });
        return true;
    }

    void dispatch()         // mu held
    {
        while (!parallel_done && order.size() < max_inflight) {
            if (pos >= file_size) { parallel_done = true; break; }
            const uint64_t n = member_size(fd, pos, file_size);
            uint8_t tr[4];
            if (!n || pread(fd, tr, 4, (off_t)(pos + n - 4)) != 4) { parallel_done = true; break; }
            Member *m = new Member();
            m->off = pos;
            m->clen = n;
            m->isize = (uint32_t)tr[0] | ((uint32_t)tr[1] << 8) | ((uint32_t)tr[2] << 16) | ((uint32_t)tr[3] << 24);
            pos += n;
            order.push_back(m);
            jobs.push_back(m);
            cv_job.notify_one();
        }
    }

    // like gzread: up to `want` bytes, 0 at the end of the input, -1 on error (err set)
    int64_t read(uint8_t *dst, uint64_t want)
    {
        uint64_t got = 0;
        while (got < want) {
            if (!cur) {
                std::unique_lock<std::mutex> lk(mu);
                dispatch();
                if (order.empty()) {
                    lk.unlock();
                    if (pos >= file_size) break;                        // clean end
                    if (!tail) {                                        // a foreign member: serial from here on
                        if (lseek(fd, (off_t)pos, SEEK_SET) < 0 || !(tail = gzdopen(dup(fd), "rb"))) {
                            err = "cannot continue reading the input";
                            return -1;
                        }
                        gzbuffer(tail, 1u << 20);
                    }
                    uint64_t w = want - got;
                    if (w > (1u << 30)) w = 1u << 30;
                    const int g = gzread(tail, dst + got, (unsigned)w);
                    if (g < 0) { err = "reading the input: gzread failed"; return -1; }
                    if (g == 0) {
                        int zerr = Z_OK;
                        gzerror(tail, &zerr);
                        if (zerr != Z_OK && zerr != Z_STREAM_END) { err = "reading the input: truncated gzip stream"; return -1; }
                        pos = file_size;
                        break;
                    }
                    got += (uint64_t)g;
                    continue;
                }
                Member *m = order.front();
                cv_done.wait(lk, [&] { return m->done; });
                order.pop_front();
                dispatch();
                if (!m->err.empty()) { err = "reading the input: " + m->err; free(m->out); delete m; return -1; }
                cur = m;
                cur_pos = 0;
            }
            const uint64_t n = std::min<uint64_t>(want - got, (uint64_t)cur->isize - cur_pos);
            memcpy(dst + got, cur->out + cur_pos, n);
            got += n;
            cur_pos += n;
            if (cur_pos == cur->isize) { free(cur->out); delete cur; cur = nullptr; }
        }
        return (int64_t)got;
    }

    ~MemberSource()
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
            cv_job.notify_all();
        }
        for (std::thread &t : pool) t.join();
        for (Member *m : order) { free(m->out); delete m; }
        if (cur) { free(cur->out); delete cur; }
        if (tail) gzclose(tail);
        if (fd >= 0) close(fd);
    }
};

}  // namespace

struct orc_reader {
    gzFile gz = nullptr;
    MemberSource *members = nullptr;    // set instead of gz for files orc_writer wrote
    orcpgz::Source *pgz = nullptr;      // set instead of gz for any other .gz file: chunk-parallel inflate
    int plain_fd = -1;                  // set instead of gz for a file that is not gzip: read() straight into the buffers
    uint32_t max_reads = 0;
    uint64_t max_bytes = 0;
    std::vector<ReaderBuf> bufs;
    std::deque<int> free_q, ready_q;
    std::mutex mu;
    std::condition_variable cv_free, cv_ready;
    std::thread th;
    bool stop = false, done = false;
    int error = 0;
    std::string err;
    std::vector<uint8_t> carry;
    double t_read = 0, t_index = 0, t_wait_free = 0;    // reader thread: inflating / copying, indexing, waiting for a free buffer (ORC_IO_DEBUG=1)

    void fail(const std::string &what)
    {
        std::lock_guard<std::mutex> lk(mu);
        error = ORC_EINVAL;
        err = what;
        done = true;
        cv_ready.notify_all();
    }

    void finish()
    {
        std::lock_guard<std::mutex> lk(mu);
        done = true;
        cv_ready.notify_all();
    }

    void run()
    {
        bool eof = false;
        for (;;) {
            int b;
            {
                const double w0 = orcpgz::Source::now();
                std::unique_lock<std::mutex> lk(mu);
                cv_free.wait(lk, [&] { return stop || !free_q.empty(); });
                t_wait_free += orcpgz::Source::now() - w0;
                if (stop) return;
                b = free_q.front();
                free_q.pop_front();
            }
            ReaderBuf &rb = bufs[b];
            uint64_t fill = carry.size();
            if (fill) memcpy(rb.text, carry.data(), fill);
            carry.clear();
            const double r0 = orcpgz::Source::now();
            while (fill < max_bytes && !eof) {
                uint64_t want = max_bytes - fill;
                if (want > (1u << 30)) want = 1u << 30;
                int64_t got;
                if (members) {
                    got = members->read(rb.text + fill, want);
                    if (got < 0) return fail(members->err);
                    if (got == 0) eof = true;
                } else if (pgz) {
                    got = pgz->read(rb.text + fill, want);
                    if (got < 0) return fail(pgz->err);
                    if (got == 0) eof = true;
                } else if (plain_fd >= 0) {
                    do got = ::read(plain_fd, rb.text + fill, want); while (got < 0 && errno == EINTR);
                    if (got < 0) return fail(std::string("reading the input: ") + strerror(errno));
                    if (got == 0) eof = true;
                } else {
                    got = gzread(gz, rb.text + fill, (unsigned)want);
                    if (got < 0) {
                        int zerr = 0;
                        const char *msg = gzerror(gz, &zerr);
                        return fail(std::string("reading the input: ") + (msg ? msg : "gzread failed"));
                    }
                    if (got == 0) {
                        int zerr = Z_OK;
                        const char *msg = gzerror(gz, &zerr);       // a truncated .gz ends "cleanly" with Z_BUF_ERROR
                        if (zerr != Z_OK && zerr != Z_STREAM_END)
                            return fail(std::string("reading the input: ") + (msg && *msg ? msg : "truncated gzip stream"));
                        eof = true;
                    }
                }
                fill += (uint64_t)got;
                {
                    std::lock_guard<std::mutex> lk(mu);
                    if (stop) return;
                }
            }
            t_read += orcpgz::Source::now() - r0;
            if (fill == 0) return finish();
            uint64_t consumed = 0;
            char ebuf[256] = {0};
            const double i0 = orcpgz::Source::now();
            int64_t n = orc_fastq_index(rb.text, fill, max_reads, eof ? 1 : 0, rb.off, rb.len, rb.qoff, rb.noff,
                                        rb.nlen, &consumed, ebuf, sizeof ebuf);
            t_index += orcpgz::Source::now() - i0;
            if (n < 0) return fail(ebuf);
            if (n == 0) {
                if (eof) return finish();       // nothing but blank lines
                return fail("FASTQ record larger than the batch buffer");
            }
            carry.assign(rb.text + consumed, rb.text + fill);
            rb.n_bytes = consumed;
            rb.n_reads = (uint32_t)n;
            uint64_t bases = 0;
            for (int64_t i = 0; i < n; i++) bases += rb.len[i];
            rb.bases = bases;
            bool last = eof && carry.empty();
            {
                std::lock_guard<std::mutex> lk(mu);
                ready_q.push_back(b);
                if (last) done = true;
                cv_ready.notify_all();
            }
            if (last) return;
        }
    }
};

extern "C" orc_reader *orc_reader_open(const char *path, uint32_t max_reads, uint64_t max_bytes, int n_buffers,
                                       int pinned, char *err, size_t err_len)
{
    unsigned hw = std::thread::hardware_concurrency();
    return orc_reader_open_threads(path, max_reads, max_bytes, n_buffers, pinned, (int)(hw > 16 ? 8 : (hw > 1 ? hw / 2 : 1)),
                                   err, err_len);
}

extern "C" orc_reader *orc_reader_open_threads(const char *path, uint32_t max_reads, uint64_t max_bytes, int n_buffers,
                                               int pinned, int inflate_threads, char *err, size_t err_len)
{
    if (!path || !max_reads || max_bytes < 16 || n_buffers < 1 || n_buffers > 64) {
        set_err(err, err_len, "orc_reader_open: bad argument");
        return nullptr;
    }
    orc_reader *r = new orc_reader();
    r->max_reads = max_reads;
    r->max_bytes = max_bytes;
    if (strcmp(path, "-") != 0 && inflate_threads > 0 && MemberSource::probe(path)) {
        r->members = new MemberSource();
        if (!r->members->start(path, inflate_threads)) {
            delete r->members;
            r->members = nullptr;
        }
    }
    // a .gz of any other origin (the pychopped_<dataset>.fastq.gz of 02_cutadapt_loop.sh:64-72): chunks of the
    // compressed stream inflated on the pool (orc_pgz.h).  ORC_NO_PGZ=1 keeps the single zlib stream,
    // ORC_PGZ_MIN / ORC_PGZ_CHUNK (bytes) move the smallest file and the chunk size (tests).
    if (!r->members && strcmp(path, "-") != 0 && inflate_threads >= 2 && !getenv("ORC_NO_PGZ")) {
        const char *e_min = getenv("ORC_PGZ_MIN"), *e_chunk = getenv("ORC_PGZ_CHUNK");
        const size_t min_bytes = e_min ? (size_t)strtoull(e_min, nullptr, 10) : (size_t)4 << 20;
        const size_t chunk = e_chunk ? (size_t)strtoull(e_chunk, nullptr, 10) : (size_t)1 << 20;
        if (orcpgz::Source::probe(path, min_bytes)) {
            r->pgz = new orcpgz::Source();
            if (!r->pgz->start(path, inflate_threads, chunk)) {
                delete r->pgz;
                r->pgz = nullptr;
            }
        }
    }
    // a regular file that is not gzip (pychopper's *_pass.fastq, 01_pychopper.sh:57): no zlib in between
    if (!r->members && !r->pgz && strcmp(path, "-") != 0) {
        const int fd = open(path, O_RDONLY);
        uint8_t magic[2] = {0, 0};
        if (fd >= 0) {
            const ssize_t k = pread(fd, magic, 2, 0);
            if (k >= 0 && !(k == 2 && magic[0] == 0x1f && magic[1] == 0x8b)) r->plain_fd = fd;
            else close(fd);
        }
    }
    if (!r->members && !r->pgz && r->plain_fd < 0) {
        if (strcmp(path, "-") == 0) r->gz = gzdopen(dup(0), "rb");
        else r->gz = gzopen(path, "rb");        // transparent for files that are not gzip
        if (!r->gz) {
            set_err(err, err_len, std::string("cannot open ") + path + ": " + strerror(errno));
            delete r;
            return nullptr;
        }
        gzbuffer(r->gz, 1u << 20);
    }
    r->bufs.resize(n_buffers);
    bool ok = true;
    for (int b = 0; b < n_buffers && ok; b++) {
        ReaderBuf &rb = r->bufs[b];
        rb.text = (uint8_t *)host_alloc(max_bytes, pinned != 0, &rb.pinned[0]);
        rb.off = (uint64_t *)host_alloc(8ull * max_reads, pinned != 0, &rb.pinned[1]);
        rb.qoff = (uint64_t *)host_alloc(8ull * max_reads, pinned != 0, &rb.pinned[2]);
        rb.noff = (uint64_t *)host_alloc(8ull * max_reads, pinned != 0, &rb.pinned[3]);
        rb.len = (uint32_t *)host_alloc(4ull * max_reads, pinned != 0, &rb.pinned[4]);
        rb.nlen = (uint32_t *)host_alloc(4ull * max_reads, pinned != 0, &rb.pinned[5]);
        ok = rb.text && rb.off && rb.qoff && rb.noff && rb.len && rb.nlen;
        r->free_q.push_back(b);
    }
    if (!ok) {
        set_err(err, err_len, "orc_reader_open: out of host memory");
        orc_reader_close(r);
        return nullptr;
    }
    r->th = std::thread([r] { r->run(); });
    return r;
}

extern "C" int orc_reader_next(orc_reader *r, orc_text_batch *out)
{
    if (!r || !out) return ORC_EINVAL;
    std::unique_lock<std::mutex> lk(r->mu);
    r->cv_ready.wait(lk, [&] { return !r->ready_q.empty() || r->done; });
    if (r->ready_q.empty()) return r->error ? r->error : 0;
    int b = r->ready_q.front();
    r->ready_q.pop_front();
    const ReaderBuf &rb = r->bufs[b];
    out->text = rb.text;
    out->n_bytes = rb.n_bytes;
    out->n_reads = rb.n_reads;
    out->buffer = b;
    out->offsets = rb.off;
    out->lengths = rb.len;
    out->qual_offsets = rb.qoff;
    out->name_offsets = rb.noff;
    out->name_lengths = rb.nlen;
    out->total_bases = rb.bases;
    return 1;
}

extern "C" int orc_reader_release(orc_reader *r, int buffer)
{
    if (!r || buffer < 0 || buffer >= (int)r->bufs.size()) return ORC_EINVAL;
    std::lock_guard<std::mutex> lk(r->mu);
    for (int b : r->free_q) if (b == buffer) return ORC_ESTATE;
    for (int b : r->ready_q) if (b == buffer) return ORC_ESTATE;
    r->free_q.push_back(buffer);
    r->cv_free.notify_all();
    return ORC_OK;
}

// How the input is inflated: 0 one zlib stream (or plain text), 1 member-parallel (files of orc_writer),
// 2 chunk-parallel (any other .gz).  stats (may be NULL): text bytes the pool decoded / the reader thread decoded
// itself so far (mode 2 only).
extern "C" int orc_reader_inflate_mode(orc_reader *r, uint64_t stats[2])
{
    if (!r) return ORC_EINVAL;
    if (stats) {
        stats[0] = r->pgz ? r->pgz->stat_parallel.load() : 0;
        stats[1] = r->pgz ? r->pgz->stat_serial.load() : 0;
    }
    return r->pgz ? 2 : r->members ? 1 : 0;
}

// A whole .gz file through the chunk-parallel inflate (orc_pgz.h) into out[0 .. cap): what orc_reader does with a
// foreign .gz, without the FASTQ indexing -- for tools and for tests of streams that are not FASTQ.
extern "C" int64_t orc_gunzip_file(const char *path, int threads, uint64_t chunk_bytes, uint8_t *out, uint64_t cap,
                                   char *err, size_t err_len)
{
    if (!path || (!out && cap)) { set_err(err, err_len, "orc_gunzip_file: bad argument"); return ORC_EINVAL; }
    orcpgz::Source src;
    if (!src.start(path, threads, chunk_bytes ? (size_t)chunk_bytes : (size_t)1 << 20)) {
        set_err(err, err_len, std::string("orc_gunzip_file: ") + path + " is not a gzip file that can be mapped");
        return ORC_EINVAL;
    }
    uint64_t n = 0;
    for (;;) {
        uint8_t over;
        const int64_t got = n < cap ? src.read(out + n, cap - n) : src.read(&over, 1);
        if (got < 0) { set_err(err, err_len, src.err); return ORC_EINVAL; }
        if (got == 0) return (int64_t)n;
        if (n >= cap) { set_err(err, err_len, "orc_gunzip_file: the text does not fit"); return ORC_ECAPACITY; }
        n += (uint64_t)got;
    }
}

extern "C" const char *orc_reader_error(orc_reader *r)
{
    return r ? r->err.c_str() : "null reader";
}

extern "C" void orc_reader_close(orc_reader *r)
{
    if (!r) return;
    {
        std::lock_guard<std::mutex> lk(r->mu);
        r->stop = true;
        r->cv_free.notify_all();
    }
    if (r->th.joinable()) r->th.join();
    if (getenv("ORC_IO_DEBUG"))
        fprintf(stderr, "orc_reader: inflate mode %d, reader thread: read %.3f s (chunk-parallel: waiting %.3f, copying %.3f, own decoding %.3f), "
                        "index %.3f s, waiting for a free buffer %.3f s\n", r->pgz ? 2 : r->members ? 1 : 0, r->t_read,
                r->pgz ? r->pgz->stat_wait_s : 0.0, r->pgz ? r->pgz->stat_copy_s : 0.0, r->pgz ? r->pgz->stat_serial_s : 0.0,
                r->t_index, r->t_wait_free);
    if (r->gz) gzclose(r->gz);
    if (r->plain_fd >= 0) close(r->plain_fd);
    delete r->members;
    delete r->pgz;
    for (ReaderBuf &rb : r->bufs) {
        host_free(rb.text, rb.pinned[0]);
        host_free(rb.off, rb.pinned[1]);
        host_free(rb.qoff, rb.pinned[2]);
        host_free(rb.noff, rb.pinned[3]);
        host_free(rb.len, rb.pinned[4]);
        host_free(rb.nlen, rb.pinned[5]);
    }
    delete r;
}

// ------------------------------------------------------------------------------------ writer

namespace {

struct Chunk {
    int bin = 0;
    const uint8_t *src = nullptr;
    size_t len = 0;
    int64_t ticket = 0;
    uint8_t *out = nullptr;         // deflated member (or nullptr for plain files: src is copied)
    size_t out_len = 0;
    bool done = false;
    bool member = false;            // src is a finished gzip member (orc_writer_write_members)
};

struct BinFile {
    int fd = -1;
    bool gz = false;
    bool writing = false;           // one thread at a time appends to the file
    uint64_t bytes_in = 0;
    std::deque<Chunk *> q;          // submission order == file order
    std::string path;
    std::vector<uint64_t> log;      // (ticket, bytes written) per chunk, in file order
};

void put_member_size(uint8_t *member, size_t total)
{
    for (int i = 0; i < 4; i++) member[GZ_SIZE_AT + i] = (uint8_t)(total >> (8 * i));
}

}  // namespace

struct orc_writer {
    std::vector<BinFile> bins;
    int level = 5;
    size_t chunk_bytes = 4u << 20;
    std::vector<std::thread> pool;
    std::deque<Chunk *> tasks;
    std::mutex mu;
    std::condition_variable cv_task, cv_done;
    bool stop = false;
    int error = 0;
    std::string err;
    int64_t next_ticket = 0;
    std::deque<int64_t> pending;    // chunks of ticket (first_ticket + i) not yet deflated
    int64_t first_ticket = 0;
    size_t unwritten = 0;           // chunks not yet in their file

    void set_error(const std::string &what)
    {
        if (!error) {
            error = ORC_EINVAL;
            err = what;
        }
    }

    bool index_files = false;       // orc_writer_set_index: leave PATH.idx beside every bin file

    bool deflate_chunk(z_stream &zs, bool &zs_ready, Chunk *c, std::string &why)
    {
        static uint8_t extra[GZ_EXTRA_LEN] = {'O', 'C', 4, 0, 0, 0, 0, 0};
        gz_header head;
        memset(&head, 0, sizeof head);
        head.os = 255;
        head.extra = extra;
        head.extra_len = (uInt)GZ_EXTRA_LEN;
        if (!zs_ready) {
            memset(&zs, 0, sizeof zs);
            if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) {
                why = "deflateInit2 failed";
                return false;
            }
            zs_ready = true;
        } else {
            deflateReset(&zs);
        }
        if (deflateSetHeader(&zs, &head) != Z_OK) {
            why = "deflateSetHeader failed";
            return false;
        }
        size_t bound = deflateBound(&zs, (uLong)c->len) + 64 + GZ_EXTRA_LEN + 2;
        c->out = (uint8_t *)malloc(bound);
        if (!c->out) {
            why = "out of memory";
            return false;
        }
        zs.next_in = const_cast<Bytef *>(c->src);
        zs.avail_in = (uInt)c->len;
        zs.next_out = c->out;
        zs.avail_out = (uInt)bound;
        if (deflate(&zs, Z_FINISH) != Z_STREAM_END) {
            why = "deflate failed";
            return false;
        }
        c->out_len = bound - zs.avail_out;
        put_member_size(c->out, c->out_len);
        return true;
    }

    // a member of orc_writer_write_members() headed for a plain (not .gz) file
    static bool inflate_member(Chunk *c, std::string &why)
    {
        if (c->len < 18) { why = "gzip member too short"; return false; }
        uint32_t isize = 0;
        for (int i = 0; i < 4; i++) isize |= (uint32_t)c->src[c->len - 4 + i] << (8 * i);
        c->out = (uint8_t *)malloc(isize ? isize : 1);
        if (!c->out) { why = "out of memory"; return false; }
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, 15 + 16) != Z_OK) { why = "inflateInit2 failed"; return false; }
        zs.next_in = const_cast<Bytef *>(c->src);
        zs.avail_in = (uInt)c->len;
        zs.next_out = c->out;
        zs.avail_out = (uInt)isize;
        const int rc = inflate(&zs, Z_FINISH);
        const bool ok = rc == Z_STREAM_END && zs.avail_out == 0;
        inflateEnd(&zs);
        if (!ok) { why = "gzip member does not inflate to its ISIZE"; return false; }
        c->out_len = isize;
        return true;
    }

    static bool write_all(int fd, const uint8_t *p, size_t n)
    {
        while (n) {
            ssize_t w = ::write(fd, p, n);
            if (w < 0) {
                if (errno == EINTR) continue;
                return false;
            }
            p += w;
            n -= (size_t)w;
        }
        return true;
    }

    void worker()
    {
        z_stream zs;
        bool zs_ready = false;
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_task.wait(lk, [&] { return stop || !tasks.empty(); });
            if (tasks.empty()) break;       // stop is only set once everything is written
            Chunk *c = tasks.front();
            tasks.pop_front();
            BinFile &bf = bins[c->bin];
            lk.unlock();
            std::string why;
            bool ok = true;
            if (c->member && !bf.gz) {
                ok = inflate_member(c, why);
            } else if (bf.gz && !c->member) {
                ok = deflate_chunk(zs, zs_ready, c, why);
            } else {
                c->out = (uint8_t *)malloc(c->len ? c->len : 1);
                ok = c->out != nullptr;
                if (ok) memcpy(c->out, c->src, c->len);
                else why = "out of memory";
                c->out_len = c->len;
            }
            lk.lock();
            if (!ok) set_error(why);
            c->done = true;
            pending[(size_t)(c->ticket - first_ticket)]--;      // the source bytes are no longer needed
            cv_done.notify_all();
            if (!bf.writing) {
                bf.writing = true;
                while (!bf.q.empty() && bf.q.front()->done) {
                    Chunk *w = bf.q.front();
                    bf.q.pop_front();
                    lk.unlock();
                    bool wok = !w->out || write_all(bf.fd, w->out, w->out_len);
                    free(w->out);
                    lk.lock();
                    bf.log.push_back((uint64_t)w->ticket);
                    bf.log.push_back((uint64_t)w->out_len);
                    if (!wok) set_error(std::string("write failed: ") + strerror(errno));
                    delete w;
                    unwritten--;
                }
                bf.writing = false;
                cv_done.notify_all();
            }
        }
        lk.unlock();
        if (zs_ready) deflateEnd(&zs);
    }
};

// An empty gzip member in this library's format (with the size field).  Returns its length, 0 on failure.
extern "C" size_t orc_empty_gzip_member(uint8_t *out, size_t cap, int level)
{
    static uint8_t extra[GZ_EXTRA_LEN] = {'O', 'C', 4, 0, 0, 0, 0, 0};
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    gz_header head;
    memset(&head, 0, sizeof head);
    head.os = 255;
    head.extra = extra;
    head.extra_len = (uInt)GZ_EXTRA_LEN;
    if (cap < 40 || deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return 0;
    size_t n = 0;
    if (deflateSetHeader(&zs, &head) == Z_OK) {
        zs.next_out = out;
        zs.avail_out = (uInt)cap;
        if (deflate(&zs, Z_FINISH) == Z_STREAM_END) {
            n = cap - zs.avail_out;
            put_member_size(out, n);
        }
    }
    deflateEnd(&zs);
    return n;
}

extern "C" int orc_writer_set_index(orc_writer *w, int on)
{
    if (!w) return ORC_EINVAL;
    std::lock_guard<std::mutex> lk(w->mu);
    w->index_files = on != 0;
    return ORC_OK;
}

extern "C" orc_writer *orc_writer_open(const char *const *paths, int n_bins, int level, int threads, char *err,
                                       size_t err_len)
{
    if (!paths || n_bins < 1 || level < 0 || level > 9) {
        set_err(err, err_len, "orc_writer_open: bad argument");
        return nullptr;
    }
    orc_writer *w = new orc_writer();
    w->level = level;
    w->bins.resize(n_bins);
    for (int b = 0; b < n_bins; b++) {
        if (!paths[b]) continue;
        // created up front even if the bin stays empty: the reference's round-2 loop lists them
        // (02_cutadapt_loop.sh:75-85)
        int fd = open(paths[b], O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (fd < 0) {
            set_err(err, err_len, std::string("cannot create ") + paths[b] + ": " + strerror(errno));
            for (BinFile &bf : w->bins) if (bf.fd >= 0) close(bf.fd);
            delete w;
            return nullptr;
        }
        size_t n = strlen(paths[b]);
        w->bins[b].fd = fd;
        w->bins[b].path = paths[b];
        w->bins[b].gz = n >= 3 && strcmp(paths[b] + n - 3, ".gz") == 0;
    }
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    for (int t = 0; t < threads; t++) w->pool.emplace_back([w] { w->worker(); });
    return w;
}

extern "C" int64_t orc_writer_write(orc_writer *w, const uint8_t *fastq, const uint64_t *bin_offsets)
{
    if (!w || !bin_offsets) return ORC_EINVAL;
    std::lock_guard<std::mutex> lk(w->mu);
    if (w->error) return w->error;
    int64_t ticket = w->next_ticket++;
    w->pending.push_back(0);
    for (size_t b = 0; b < w->bins.size(); b++) {
        BinFile &bf = w->bins[b];
        if (bf.fd < 0) continue;
        uint64_t lo = bin_offsets[b], hi = bin_offsets[b + 1];
        if (hi <= lo) continue;
        if (!fastq) return ORC_EINVAL;
        bf.bytes_in += hi - lo;
        for (uint64_t p = lo; p < hi; p += w->chunk_bytes) {
            Chunk *c = new Chunk();
            c->bin = (int)b;
            c->src = fastq + p;
            c->len = (size_t)((hi - p < w->chunk_bytes) ? hi - p : w->chunk_bytes);
            c->ticket = ticket;
            bf.q.push_back(c);
            w->tasks.push_back(c);
            w->pending.back()++;
            w->unwritten++;
        }
    }
    w->cv_task.notify_all();
    return ticket;
}

// The bins of one batch as finished gzip members (orc_params.emit_gzip: orc_result.fastq / bin_offsets): a .gz
// bin file receives its member as it is, a plain one the inflated text.  A member's text is under 4 GiB (its
// ISIZE counts the bytes of the bin).
extern "C" int64_t orc_writer_write_members(orc_writer *w, const uint8_t *members, const uint64_t *member_offsets)
{
    if (!w || !member_offsets) return ORC_EINVAL;
    std::lock_guard<std::mutex> lk(w->mu);
    if (w->error) return w->error;
    for (size_t b = 0; b < w->bins.size(); b++) {
        const uint64_t lo = member_offsets[b], hi = member_offsets[b + 1];
        if (hi < lo || (hi > lo && (!members || hi - lo < 28 || members[lo] != 0x1f || members[lo + 1] != 0x8b))) return ORC_EINVAL;
    }
    int64_t ticket = w->next_ticket++;
    w->pending.push_back(0);
    for (size_t b = 0; b < w->bins.size(); b++) {
        BinFile &bf = w->bins[b];
        if (bf.fd < 0) continue;
        const uint64_t lo = member_offsets[b], hi = member_offsets[b + 1];
        if (hi <= lo) continue;
        uint32_t isize = 0;
        for (int i = 0; i < 4; i++) isize |= (uint32_t)members[hi - 4 + i] << (8 * i);
        bf.bytes_in += isize;
        Chunk *c = new Chunk();
        c->bin = (int)b;
        c->src = members + lo;
        c->len = (size_t)(hi - lo);
        c->ticket = ticket;
        c->member = true;
        bf.q.push_back(c);
        w->tasks.push_back(c);
        w->pending.back()++;
        w->unwritten++;
    }
    w->cv_task.notify_all();
    return ticket;
}

extern "C" int orc_writer_wait(orc_writer *w, int64_t ticket)
{
    if (!w) return ORC_EINVAL;
    std::unique_lock<std::mutex> lk(w->mu);
    if (ticket < 0 || ticket >= w->next_ticket) return ORC_EINVAL;
    if (ticket >= w->first_ticket)
        w->cv_done.wait(lk, [&] { return w->pending[(size_t)(ticket - w->first_ticket)] == 0; });
    while (!w->pending.empty() && w->pending.front() == 0 && w->first_ticket < w->next_ticket) {
        w->pending.pop_front();
        w->first_ticket++;
    }
    return w->error;
}

extern "C" const char *orc_writer_error(orc_writer *w)
{
    return w ? w->err.c_str() : "null writer";
}

extern "C" int orc_writer_close(orc_writer *w, uint64_t *bytes_per_bin)
{
    if (!w) return ORC_EINVAL;
    {
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv_done.wait(lk, [&] { return w->unwritten == 0; });
        w->stop = true;
        w->cv_task.notify_all();
    }
    for (std::thread &t : w->pool) t.join();
    int rc = w->error;
    for (size_t b = 0; b < w->bins.size(); b++) {
        BinFile &bf = w->bins[b];
        if (bytes_per_bin) bytes_per_bin[b] = bf.bytes_in;
        if (bf.fd < 0) continue;
        if (bf.gz && bf.bytes_in == 0) {
            // an empty but valid .gz, as xopen leaves for a bin without reads
            uint8_t out[64];
            const size_t n = orc_empty_gzip_member(out, sizeof out, w->level);
            if (!n || !orc_writer::write_all(bf.fd, out, n)) rc = ORC_EINVAL;
        }
        if (close(bf.fd) != 0 && !rc) rc = ORC_EINVAL;
        if (w->index_files) {
            // (ticket, bytes) of every chunk in file order: what merging the part files of several ranks
            // into one file in batch order needs (orcdemux/cli.py)
            FILE *fh = fopen((bf.path + ".idx").c_str(), "wb");
            if (!fh || (bf.log.size() && fwrite(bf.log.data(), sizeof(uint64_t), bf.log.size(), fh) != bf.log.size())) rc = rc ? rc : ORC_EINVAL;
            if (fh && fclose(fh) != 0 && !rc) rc = ORC_EINVAL;
        }
    }
    delete w;
    return rc;
}
