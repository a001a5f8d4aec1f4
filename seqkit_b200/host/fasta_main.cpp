// fasta_main.cpp -- the drop-in `fasta` binary for the per-read FASTQ batch path, on top of the C ABI
// of libseqkit_b200.so (include/seqkit_b200.h).
//
// Mirrors, for the four subcommands on the path, what the reference's dispatcher and modules do at
// the process boundary (paths relative to /root/reference/src/):
//   fasta_main.rs:42-82           word-prefix dispatch, top-level usage on stderr for anything else
//   common.rs:11-22               error! -> "ERROR: <msg>\n" on stderr, exit status 255; docopt failure
//   common.rs:88-104              FileReader::new: "-" = stdin, *.gz through a `gunzip -c` child
//   common.rs:49-81               GzipWriter: File::create(path) as stdout of a `gzip -c` / `pigz -c` child
//   fasta_trim_by_quality.rs, fasta_mask_by_quality.rs, fasta_add_barcode.rs, fasta_demultiplex.rs
//
// This file is the *batcher*: it reads bytes, cuts batches at record boundaries (a newline count; no
// per-read arithmetic happens here), packs them into pinned multi-MB buffers, calls the CUDA path
// through the ABI and streams the results to stdout / the per-sample gzip sinks in batch order.  Batches are
// dealt round-robin to every visible GPU (one context per device, two slots each: consecutive batches =
// contiguous record ranges on different GPUs; SK_DEVICE pins one device, SK_GPUS limits the count); results
// are consumed strictly in batch order, so output bytes do not depend on the number of GPUs.  Per-sample
// output comes back as one contiguous slice per sample, mate and batch (sk_demux_compact); the per-sample
// counters are summed on the devices and merged with one NCCL all-reduce at the end of the run.  gzip is
// block-parallel deflate on host threads (independent gzip members, zlib), or one `gzip -c` / `pigz -c`
// child per file as in the reference with SK_GZIP=child.  Every operator result comes from the GPU; there
// is no CPU fallback.
#include <errno.h>
#include <fcntl.h>
#include <signal.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include <time.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "seqkit_b200.h"

// ------------------------------------------------------------------------------------------------
// process-level helpers
// ------------------------------------------------------------------------------------------------
static const char *TOP_USAGE =  // fasta_main.rs:20-40, restricted to nothing: printed verbatim like the reference
    "\nUsage:\n"
    "  fasta check <fasta/fastq>\n"
    "  fasta to raw <fasta/fastq>\n"
    "  fasta add base qualities <fasta> <baseq>\n"
    "  fasta remove base qualities <fastq>\n"
    "  fasta simplify read ids <fastq_file>\n"
    "  fasta interleave <fastq_1> <fastq_2>\n"
    "  fasta deinterleave <interleaved_fastq> <out_prefix>\n"
    "  fasta split into anchors <fastq> <anchor_len>\n"
    "  fasta trim <fastq_file>\n"
    "  fasta trim by quality <fastq_file> <min_baseq>\n"
    "  fasta mask by quality <fastq_file> <min_baseq>\n"
    "  fasta gc content <genome.fa> <regions.bed>\n"
    "  fasta add barcode <fastq_file> <barcode_file> <barcode_format>\n"
    "  fasta extract dual umi <interleaved_fastq>\n"
    "  fasta convert basespace <fastq_file>\n"
    "  fasta demultiplex <sample_sheet> <fastq_1> <fastq_2>\n"
    "  fasta demultiplex spe <sample_sheet> <fastq_1> <fastq_2>\n"
    "  fasta statistics <fastq_file>\n";

static const char *USAGE_TRIM = "\nUsage:\n  fasta trim by quality <fastq_file> <min_baseq>\n";
static const char *USAGE_MASK = "\nUsage:\n  fasta mask by quality <fastq_file> <min_baseq>\n";
static const char *USAGE_ADDBC = "\nUsage:\n  fasta add barcode <fastq_file> <barcode_file>\n";
static const char *USAGE_DEMUX =
    "\nUsage:\n"
    "  fasta demultiplex [options] <sample_sheet> <fastq_1> [<fastq_2>]\n"
    "\nOptions:\n"
    "  --parallel      Use pigz (parallel gzip) for compression\n"
    "  --index1=FASTQ  Path to FASTQ file containing the first index (optional)\n"
    "  --index2=FASTQ  Path to FASTQ file containing the second index (optional)\n"
    "  --dry-run=N     Analyze N reads and generate table of indexes found in the run\n"
    "\nSplits a pooled FASTQ file into multiple individual FASTQ files, based on a\n"
    "sample sheet. Each read in the pooled FASTQ file must carry a BC:xxxxxxxx\n"
    "field in its header.\n";

// SK_TIMING=1: wall-clock seconds per phase of the batcher on stderr at exit (diagnostic; off by default)
static double now_s() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static const double g_t0 = now_s();
static double g_phase[6] = {0, 0, 0, 0, 0, 0};  // init, read+submit, wait, drain, close, -
struct Phase {
    int k;
    double t;
    explicit Phase(int k_) : k(k_), t(now_s()) {}
    ~Phase() { g_phase[k] += now_s() - t; }
};

struct Sink;
static std::vector<Sink *> g_sinks;  // flushed and closed (children reaped) before any exit
static void close_all_sinks();
static std::vector<pid_t> g_gunzip;  // `gunzip -c` children of *.gz inputs

[[noreturn]] static void finish(int code) {
    fflush(stdout);
    // On an early exit a gunzip child may sit blocked on its full output pipe: it goes first, so that nothing
    // that waits below can wait on it.
    for (pid_t pid : g_gunzip) kill(pid, SIGKILL);
    for (pid_t pid : g_gunzip) {
        int st;
        waitpid(pid, &st, 0);
    }
    g_gunzip.clear();
    {
        Phase ph(4);
        close_all_sinks();
    }
    if (getenv("SK_TIMING"))
        fprintf(stderr, "seqkit_b200 timing: total %.3f s = init %.3f + read/submit %.3f + wait %.3f + drain %.3f + close %.3f\n",
                now_s() - g_t0, g_phase[0], g_phase[1], g_phase[2], g_phase[3], g_phase[4]);
    fflush(stderr);
    _exit(code);
}
// error! (common.rs:11-16)
[[noreturn]] static void fatal(const char *fmt, ...) {
    fflush(stdout);
    va_list ap;
    va_start(ap, fmt);
    fputs("ERROR: ", stderr);
    vfprintf(stderr, fmt, ap);
    fputc('\n', stderr);
    va_end(ap);
    finish(255);
}
// Inputs this implementation refuses instead of guessing: not a reference message.
[[noreturn]] static void refuse(const char *fmt, ...) {
    fflush(stdout);
    va_list ap;
    va_start(ap, fmt);
    fputs("seqkit_b200: unsupported input: ", stderr);
    vfprintf(stderr, fmt, ap);
    fputc('\n', stderr);
    va_end(ap);
    finish(2);
}
[[noreturn]] static void panic101(const char *what) {
    fflush(stdout);
    fprintf(stderr, "thread 'main' panicked: %s\n", what);
    finish(101);
}
[[noreturn]] static void invalid_args(const char *usage) {  // parse_args (common.rs:18-22)
    fprintf(stderr, "ERROR: Invalid arguments.\n%s\n", usage);
    finish(255);
}
static void write_all(int fd, const uint8_t *p, size_t n) {
    while (n) {
        ssize_t k = write(fd, p, n);
        if (k < 0) {
            if (errno == EINTR) continue;
            return;  // #![allow(unused_must_use)]: write errors are ignored (fasta_main.rs:2)
        }
        p += k;
        n -= (size_t)k;
    }
}
static bool is_ws(uint8_t c) { return c == 32 || (c >= 9 && c <= 13); }  // ASCII White_Space
static size_t trim_end_len(const uint8_t *s, size_t n) {
    while (n && is_ws(s[n - 1])) n--;
    return n;
}

// ------------------------------------------------------------------------------------------------
// FileReader::new (common.rs:88-104) as a byte source
// ------------------------------------------------------------------------------------------------
struct Input {
    int fd = -1;
    pid_t child = 0;
    bool eof = false;
    // bytes of a plain regular file (0 for stdin, pipes and *.gz: their producers are far slower than one GPU)
    uint64_t plain_bytes() const {
        struct stat sb;
        if (fd < 0 || child || fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) return 0;
        return (uint64_t)sb.st_size;
    }
    void open_path(const std::string &path) {
        if (path == "-") {
            fd = 0;
            return;
        }
        int f = open(path.c_str(), O_RDONLY | O_CLOEXEC);  // (every descriptor is close-on-exec: children inherit 0/1/2 only)
        if (f < 0) fatal("Cannot open file %s for reading.", path.c_str());
        if (path.size() >= 3 && path.compare(path.size() - 3, 3, ".gz") == 0) {
            int pp[2];
            if (pipe2(pp, O_CLOEXEC) != 0) fatal("Cannot start gunzip process.");
            pid_t pid = fork();
            if (pid < 0) fatal("Cannot start gunzip process.");
            if (pid == 0) {
                dup2(f, 0);
                dup2(pp[1], 1);
                close(pp[0]);
                close(pp[1]);
                close(f);
                execlp("gunzip", "gunzip", "-c", (char *)nullptr);
                _exit(127);
            }
            close(pp[1]);
            close(f);
            fd = pp[0];
            child = pid;
            g_gunzip.push_back(pid);
            fcntl(fd, F_SETPIPE_SZ, 1 << 20);
        } else {
            fd = f;
        }
    }
    size_t read_some(uint8_t *dst, size_t cap) {
        size_t got = 0;
        while (got < cap && !eof) {
            ssize_t k = read(fd, dst + got, cap - got);
            if (k < 0) {
                if (errno == EINTR) continue;
                fatal("I/O error while reading from file.");  // common.rs:110
            }
            if (k == 0) eof = true;
            got += (size_t)k;
        }
        return got;
    }
};

// One input stream of the batcher: bytes accumulate in one buffer of a ring of pinned buffers; the buffers
// behind it in the ring belong to batches that are still in flight (sources of H2D copies, and of the header
// text quoted in messages), so a buffer is written again only after every batch cut from it has been
// consumed.  Complete records are found by counting newlines.
struct Stream {
    Input in;
    bool active = false;
    std::vector<uint8_t *> ring;
    size_t cap = 0;
    int cur = 0;
    size_t fill = 0, scanned = 0;
    uint32_t lines_mod = 0;
    uint32_t lpr = 4;
    std::vector<uint32_t> rec_ends;  // end offset (exclusive) of every complete record in ring[cur]
    uint64_t records_done = 0;

    uint8_t *buf() const { return ring[(size_t)cur]; }
    void top_up() {
        if (!active) return;
        if (!in.eof && fill < cap) fill += in.read_some(buf() + fill, cap - fill);
        const uint8_t *b = buf();
        while (scanned < fill) {
            const uint8_t *nl = (const uint8_t *)memchr(b + scanned, '\n', fill - scanned);
            if (!nl) {
                scanned = fill;
                break;
            }
            scanned = (size_t)(nl - b) + 1;
            if (++lines_mod == lpr) {
                lines_mod = 0;
                rec_ends.push_back((uint32_t)scanned);
            }
        }
    }
    size_t last_end() const { return rec_ends.empty() ? 0 : rec_ends.back(); }
    // records available now; at EOF the unterminated remainder is one more (truncated) record
    size_t avail() const { return rec_ends.size() + ((in.eof && fill > last_end()) ? 1 : 0); }
    bool drained() const { return in.eof && fill == 0; }
    size_t bytes_for(size_t n) const { return n == 0 ? 0 : (n <= rec_ends.size() ? rec_ends[n - 1] : fill); }
    // moves on to the next buffer of the ring without dropping anything (the current one stays as it is)
    void park() { cur = (cur + 1) % (int)ring.size(); }
    // drops the first `bytes` (= n records) of the buffer; the tail moves to the next buffer of the ring
    void consume(size_t n, size_t bytes) {
        const int nxt = (cur + 1) % (int)ring.size();
        const size_t tail = fill - bytes;
        if (tail) memcpy(ring[(size_t)nxt], buf() + bytes, tail);
        const size_t nr = std::min(n, rec_ends.size());
        rec_ends.erase(rec_ends.begin(), rec_ends.begin() + nr);
        for (auto &e : rec_ends) e -= (uint32_t)bytes;
        fill = tail;
        scanned -= bytes;
        cur = nxt;
        records_done += n;
    }
};

// ------------------------------------------------------------------------------------------------
// GzipWriter (common.rs:49-81)
// ------------------------------------------------------------------------------------------------
// A per-sample output file.  The reference wires File::create(path) to the stdout of a `gzip -c` (or, with
// --parallel, `pigz -c`) child and writes into the child's stdin; ChildSink does exactly that (SK_GZIP=child).
// The default, DeflateSink, compresses in this process: the bytes of a file are cut into blocks, every block
// becomes one gzip member (RFC 1952 allows a file to be a sequence of members; `gunzip`, `zcat` and every
// gzip reader concatenate them) on a pool of host threads, and the members are written in order.  The
// decompressed bytes are the same; the compressed bytes are not (nor are they with pigz).
struct Sink {
    virtual ~Sink() {}
    virtual void append(const uint8_t *p, size_t n) = 0;
    virtual void flush_last() = 0;  // no more data: hand the rest to the compressor / close the pipe
    virtual void close_wait() = 0;  // after flush_last (and, for DeflateSink, after the pool has drained)
};

struct ChildSink : Sink {
    int fd = -1;
    pid_t child = 0;
    void open_path(const std::string &path, bool pigz) {
        int f = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0666);
        if (f < 0) fatal("Cannot open file %s for writing.", path.c_str());
        int pp[2];
        if (pipe2(pp, O_CLOEXEC) != 0) fatal("Cannot start %s process.", pigz ? "pigz" : "gzip");
        pid_t pid = fork();
        if (pid < 0) fatal("Cannot start %s process.", pigz ? "pigz" : "gzip");
        if (pid == 0) {
            dup2(pp[0], 0);
            dup2(f, 1);
            execlp(pigz ? "pigz" : "gzip", pigz ? "pigz" : "gzip", "-c", (char *)nullptr);
            _exit(127);
        }
        close(pp[0]);
        close(f);
        fd = pp[1];
        child = pid;
        fcntl(fd, F_SETPIPE_SZ, 1 << 20);
        g_sinks.push_back(this);
    }
    void append(const uint8_t *p, size_t n) override { write_all(fd, p, n); }
    void flush_last() override {
        if (fd >= 0) close(fd);
        fd = -1;
    }
    void close_wait() override {
        flush_last();
        if (child > 0) {
            int st;
            waitpid(child, &st, 0);
        }
        child = 0;
    }
};

struct DeflateSink;
struct DeflateJob {
    DeflateSink *sink;
    uint64_t seq;
    std::vector<uint8_t> data;
};
// The compressor threads.  submit() blocks while more than `limit` uncompressed bytes are queued.
struct DeflatePool {
    std::mutex m;
    std::condition_variable cv_job, cv_space, cv_idle;
    std::deque<DeflateJob> q;
    size_t queued = 0, busy = 0;
    const size_t limit = 256u << 20;
    bool stop = false;
    int level = 4;
    std::vector<std::thread> th;
    void start() {
        if (!th.empty()) return;
        if (const char *e = getenv("SK_GZIP_LEVEL")) level = std::max(1, std::min(9, atoi(e)));
        unsigned n = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
        if (const char *e = getenv("SK_GZIP_THREADS")) n = (unsigned)std::max(1, atoi(e));
        for (unsigned i = 0; i < n; i++) th.emplace_back([this] { work(); });
    }
    void submit(DeflateJob &&j) {
        start();
        std::unique_lock<std::mutex> lk(m);
        cv_space.wait(lk, [&] { return queued < limit; });
        queued += j.data.size();
        q.push_back(std::move(j));
        cv_job.notify_one();
    }
    void wait_idle() {
        std::unique_lock<std::mutex> lk(m);
        cv_idle.wait(lk, [&] { return q.empty() && busy == 0; });
    }
    void shutdown() {
        {
            std::lock_guard<std::mutex> lk(m);
            stop = true;
        }
        cv_job.notify_all();
        for (auto &t : th) t.join();
        th.clear();
    }
    void work();
};
static DeflatePool g_pool;

struct DeflateSink : Sink {
    static constexpr size_t BLOCK = 512u << 10;
    int fd = -1;
    std::vector<uint8_t> pending;
    uint64_t next_seq = 0;
    std::mutex wm;  // members leave in sequence order
    uint64_t next_write = 0;
    std::map<uint64_t, std::vector<uint8_t>> done;
    void open_path(const std::string &path) {
        fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0666);
        if (fd < 0) fatal("Cannot open file %s for writing.", path.c_str());
        g_sinks.push_back(this);
    }
    void cut() {
        DeflateJob j{this, next_seq++, std::move(pending)};
        pending = std::vector<uint8_t>();
        g_pool.submit(std::move(j));
    }
    void append(const uint8_t *p, size_t n) override {
        while (n) {
            const size_t k = std::min(n, BLOCK - pending.size());
            pending.insert(pending.end(), p, p + k);
            p += k;
            n -= k;
            if (pending.size() >= BLOCK) cut();
        }
    }
    void flush_last() override {
        if (!pending.empty() || next_seq == 0) cut();  // an empty file is one empty member, as `gzip -c` writes it
    }
    void deliver(uint64_t seq, std::vector<uint8_t> &&z) {
        std::lock_guard<std::mutex> lk(wm);
        done.emplace(seq, std::move(z));
        while (!done.empty() && done.begin()->first == next_write) {
            write_all(fd, done.begin()->second.data(), done.begin()->second.size());
            done.erase(done.begin());
            next_write++;
        }
    }
    void close_wait() override {
        if (fd >= 0) close(fd);
        fd = -1;
    }
};
void DeflatePool::work() {
    for (;;) {
        DeflateJob j;
        {
            std::unique_lock<std::mutex> lk(m);
            cv_job.wait(lk, [&] { return stop || !q.empty(); });
            if (q.empty()) return;
            j = std::move(q.front());
            q.pop_front();
            busy++;
        }
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        std::vector<uint8_t> z;
        if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16 /* gzip wrapper */, 8, Z_DEFAULT_STRATEGY) == Z_OK) {
            z.resize(deflateBound(&zs, (uLong)j.data.size()) + 64);
            zs.next_in = j.data.data();
            zs.avail_in = (uInt)j.data.size();
            zs.next_out = z.data();
            zs.avail_out = (uInt)z.size();
            deflate(&zs, Z_FINISH);
            z.resize(z.size() - zs.avail_out);
            deflateEnd(&zs);
        }
        j.sink->deliver(j.seq, std::move(z));
        {
            std::lock_guard<std::mutex> lk(m);
            queued -= j.data.size();
            busy--;
        }
        cv_space.notify_all();
        cv_idle.notify_all();
    }
}
static void close_all_sinks() {
    for (auto *s : g_sinks) s->flush_last();
    if (!g_pool.th.empty()) {
        g_pool.wait_idle();
        g_pool.shutdown();
    }
    for (auto *s : g_sinks) s->close_wait();
    g_sinks.clear();
}

// ------------------------------------------------------------------------------------------------
// GPU contexts shared by the subcommands: one per device, batches dealt round-robin
// ------------------------------------------------------------------------------------------------
// A *lane* is one (device, slot) pair.  Batch i runs on lane i % lanes: device i % G, so consecutive batches
// (contiguous record ranges) sit on different GPUs, and slot (i / G) % NSLOT of that device.
struct Gpu {
    std::vector<sk_ctx *> ctxs;
    sk_ctx *ctx = nullptr;  // ctxs[0]: sheet-independent queries
    int G = 0;
    static const int NSLOT = 2;
    uint64_t batch_bytes = 0, max_records = 0, out_cap = 0;
    std::vector<uint8_t *> out_h[2];   // per lane and mate: pinned, grown on demand (ensure_out)
    std::vector<uint64_t> out_hcap[2];

    int lanes() const { return G * NSLOT; }
    sk_ctx *c(int lane) const { return ctxs[(size_t)(lane % G)]; }
    uint32_t slot(int lane) const { return (uint32_t)(lane / G); }
    // input_bytes: bytes of the plain input files (Input::plain_bytes).  Without SK_GPUS / SK_DEVICE the run takes one more
    // GPU per 4 GiB of such input, all visible GPUs at most: every context costs start-up time (page-locked buffers, the
    // NCCL communicator at the end), one GPU outruns any host-side producer of a small or compressed input, and the
    // output bytes do not depend on the number of GPUs anyway.
    void create(uint32_t max_samples, bool aux, int n_out, bool line_ops = false, uint64_t input_bytes = 0) {
        // 16 MiB batches: page-locking host memory is the slowest part of start-up (a few ms per MiB), and the
        // kernels lose nothing at this size
        uint64_t mb = 16;
        if (const char *e = getenv("SK_BATCH_MB")) mb = std::max<uint64_t>(1, strtoull(e, nullptr, 10));
        batch_bytes = mb << 20;
        max_records = std::max<uint64_t>(batch_bytes / 64, 1024);
        const int ndev = sk_device_count();
        std::vector<int> devs;
        if (const char *e = getenv("SK_DEVICE")) {
            devs.push_back(atoi(e));
        } else {
            int want = (int)std::min<uint64_t>((uint64_t)std::max(ndev, 1), 1 + (input_bytes >> 32));
            if (const char *g = getenv("SK_GPUS")) want = std::max(1, std::min(ndev, atoi(g)));
            for (int d = 0; d < want; d++) devs.push_back(d);
        }
        if (devs.empty()) devs.push_back(0);  // no device: sk_ctx_create reports it
        sk_limits lim;
        memset(&lim, 0, sizeof lim);
        lim.max_stream_bytes = batch_bytes;
        lim.max_records = max_records;
        lim.n_slots = NSLOT;
        lim.max_samples = max_samples;
        lim.aux_streams = aux ? 1 : 0;
        lim.reserved = line_ops ? 0x200u : 0u;
        // One thread per device: context creation and page-locking are the slow part of start-up (seconds with
        // eight GPUs when done one after the other), and nothing in them depends on another device.
        G = (int)devs.size();
        ctxs.assign((size_t)G, nullptr);
        std::vector<std::string> errs((size_t)G);
        out_cap = 0;
        for (int m = 0; m < n_out; m++) {
            out_h[m].assign((size_t)lanes(), nullptr);
            out_hcap[m].assign((size_t)lanes(), 0);
        }
        // the device-side capacity allows for 72 more bytes per record (add barcode); what the operators here
        // usually write is about the batch itself, so the host mirrors start there and grow when they must
        auto first_cap = [&](uint64_t cap) {
            return aux ? cap : std::min<uint64_t>(cap, batch_bytes + batch_bytes / 8 + (uint64_t)max_samples * 128 + (1u << 20));
        };
        per_device([&](int k) {
            sk_ctx *x = nullptr;
            if (sk_ctx_create(devs[(size_t)k], &lim, &x) != SK_OK) {
                errs[(size_t)k] = sk_last_error(nullptr);
                return;
            }
            ctxs[(size_t)k] = x;
            const uint64_t cap = first_cap(sk_out_capacity(x));
            for (int m = 0; m < n_out; m++)
                for (int l = k; l < lanes(); l += G) {
                    out_h[m][(size_t)l] = (uint8_t *)sk_pinned_alloc(x, cap);
                    out_hcap[m][(size_t)l] = cap;
                    if (!out_h[m][(size_t)l]) errs[(size_t)k] = "cannot allocate pinned memory";
                }
        });
        for (int k = 0; k < G; k++)
            if (!errs[(size_t)k].empty() || !ctxs[(size_t)k]) {
                fprintf(stderr, "seqkit_b200: cannot initialise the GPU path: %s\n", errs[(size_t)k].c_str());
                finish(3);
            }
        ctx = ctxs[0];
        out_cap = sk_out_capacity(ctx);
    }
    template <class F>
    void per_device(F f) {
        std::vector<std::thread> th;
        for (int k = 1; k < G; k++) th.emplace_back([&f, k] { f(k); });
        f(0);
        for (auto &t : th) t.join();
    }
    uint8_t *ensure_out(int m, int lane, uint64_t n) {
        if (n > out_cap) refuse("output larger than the slot capacity");
        if (n > out_hcap[m][(size_t)lane]) {
            sk_pinned_free(c(lane), out_h[m][(size_t)lane]);
            const uint64_t want = std::min<uint64_t>(out_cap, n + n / 4);
            out_h[m][(size_t)lane] = (uint8_t *)pinned(c(lane), want);
            out_hcap[m][(size_t)lane] = want;
        }
        return out_h[m][(size_t)lane];
    }
    void *pinned(sk_ctx *x, uint64_t n) {
        void *p = sk_pinned_alloc(x, n);
        if (!p) {
            fprintf(stderr, "seqkit_b200: cannot allocate %llu bytes of pinned memory\n", (unsigned long long)n);
            finish(3);
        }
        return p;
    }
    // The ring holds one buffer per batch in flight plus the one being filled: batch i is cut from ring[i % nb],
    // and ring[i % nb] is next written (the tail of batch i + nb - 1) when batch i + nb - 1 is submitted, which
    // happens only after batch i + nb - 1 - lanes() = i + G - 1 >= i has been completed and consumed.  Its length
    // is a multiple of G, so that ring[j] only ever feeds device j % G and can live on that device's NUMA node.
    void init_stream(Stream &st, const std::string &path) {
        st.active = true;
        st.cap = batch_bytes;
        const int nb = G * (NSLOT + 1);
        st.ring.assign((size_t)nb, nullptr);
        per_device([&](int k) {
            for (int j = k; j < nb; j += G) st.ring[(size_t)j] = (uint8_t *)sk_pinned_alloc(ctxs[(size_t)k], batch_bytes);
        });
        for (uint8_t *b : st.ring)
            if (!b) {
                fprintf(stderr, "seqkit_b200: cannot allocate %llu bytes of pinned memory\n", (unsigned long long)batch_bytes);
                finish(3);
            }
        if (st.in.fd < 0) st.in.open_path(path);  // (an input is opened exactly once: it may be a FIFO)
    }
    void ck(sk_ctx *x, int rc, const char *what) {
        if (rc != SK_OK) {
            fflush(stdout);
            fprintf(stderr, "seqkit_b200: %s failed (%d): %s\n", what, rc, sk_last_error(x));
            finish(3);
        }
    }
    void ck(int rc, const char *what) { ck(ctx, rc, what); }
};


static void refuse_status(const sk_result &r, uint64_t base) {
    const unsigned long long rec = (unsigned long long)(base + r.err_record);
    switch (r.status) {
        case SK_DATA_NON_ASCII: refuse("non-ASCII bytes in the input");
        case SK_DATA_RECORD_TOO_LONG: refuse("record %llu is longer than the chunk overhang", rec);
        case SK_DATA_CHUNK_TOO_DENSE: refuse("too many records in one chunk near record %llu", rec);
        case SK_DATA_MIXED_FORMAT: refuse("mixed FASTA/FASTQ records (record %llu)", rec);
        case SK_DATA_OUT_OVERFLOW: refuse("output capacity exceeded near record %llu", rec);
        case SK_DATA_TRUNCATED_FUSED: refuse("fused trim+demultiplex on a truncated header (record %llu)", rec);
        case SK_DATA_TOO_MANY_RECORDS: refuse("more than %llu records in one batch", rec);  // (the batcher counts records: not reached)
        default: break;
    }
}

// Rust's `u8::from_str` (fasta_trim_by_quality.rs:13): optional '+', decimal digits, <= 255.
static bool parse_u8(const char *s, unsigned *v) {
    if (*s == '+') s++;
    if (!*s) return false;
    unsigned long x = 0;
    for (; *s; s++) {
        if (*s < '0' || *s > '9') return false;
        x = x * 10 + (unsigned long)(*s - '0');
        if (x > 255) return false;
    }
    *v = (unsigned)x;
    return true;
}

// ------------------------------------------------------------------------------------------------
// trim by quality / mask by quality / add barcode: one ordered output stream to stdout
// ------------------------------------------------------------------------------------------------
enum StreamOp { OP_TRIM, OP_MASK, OP_ADDBC };

struct Batch {
    bool live = false;
    size_t n = 0;           // records
    size_t bytes[SK_N_INPUTS] = {0, 0, 0, 0};
    const uint8_t *src[SK_N_INPUTS] = {nullptr, nullptr, nullptr, nullptr};  // pinned copies (valid until the slot is reused)
    uint64_t first_record = 0;
};

static int run_stream_op(StreamOp op, const Input &fastq_in, const Input &aux_in, unsigned min_baseq) {
    Gpu g;
    g.create(0, op == OP_ADDBC, 1, false, fastq_in.plain_bytes());
    Stream rd, bc;
    rd.in = fastq_in;  // opened by the dispatcher, in the reference's order of errors
    g.init_stream(rd, "");
    if (op == OP_ADDBC) {
        bc.in = aux_in;
        g.init_stream(bc, "");
    }
    std::vector<uint8_t> last_bc;  // last barcode record: reused once the barcode file is exhausted (fasta_add_barcode.rs:20-27)
    bool first = true, bc_fastx = false;
    const int NL = g.lanes();
    std::vector<Batch> batches((size_t)NL);
    uint64_t bi = 0;

    auto launch = [&](int lane, uint64_t rec_limit) {
        sk_ctx *x = g.c(lane);
        const uint32_t slot = g.slot(lane);
        if (op == OP_TRIM) g.ck(x, sk_trim_by_quality(x, slot, min_baseq, rec_limit), "sk_trim_by_quality");
        else if (op == OP_MASK) g.ck(x, sk_mask_by_quality(x, slot, min_baseq, rec_limit), "sk_mask_by_quality");
        else g.ck(x, sk_add_barcode(x, slot, rec_limit), "sk_add_barcode");
    };
    auto submit = [&](int lane) -> bool {
        sk_ctx *x = g.c(lane);
        const uint32_t slot = g.slot(lane);
        Batch &B = batches[(size_t)lane];
        B = Batch();
        rd.top_up();
        if (first) {
            first = false;
            if (op == OP_ADDBC) {
                auto reframe = [](Stream &x) {  // '>' = FASTA framing, 2 lines per record (fasta_add_barcode.rs:25-27,39-40)
                    x.lpr = 2;
                    x.rec_ends.clear();
                    x.scanned = 0;
                    x.lines_mod = 0;
                    x.top_up();
                };
                if (rd.fill && rd.buf()[0] == '>') reframe(rd);
                bc.top_up();
                bc_fastx = bc.fill && (bc.buf()[0] == '@' || bc.buf()[0] == '>');
                if (bc.fill && bc.buf()[0] == '>') reframe(bc);
            }
        }
        size_t n = std::min<size_t>(rd.avail(), g.max_records);
        if (n == 0) {
            if (rd.drained()) return false;
            refuse("a record does not fit in one batch (%llu bytes); raise SK_BATCH_MB", (unsigned long long)g.batch_bytes);
        }
        // Barcode records of this batch (fasta_add_barcode.rs:20-27).  A barcode file that does not start
        // with '@' or '>' never yields a barcode; once the file is exhausted the last barcode is reused.
        size_t nb = 0;
        bool bc_reuse = false;
        if (op == OP_ADDBC && bc_fastx) {
            bc.top_up();
            if (bc.avail() == 0 && bc.in.eof) {
                bc_reuse = !last_bc.empty();
            } else {
                nb = std::min(n, bc.avail());
                if (nb < n && !bc.in.eof) n = nb;  // the rest of the barcode records is still to be read
                if (n == 0) refuse("a barcode record does not fit in one batch; raise SK_BATCH_MB");
            }
        }
        B.live = true;
        B.n = n;
        B.first_record = rd.records_done;
        B.bytes[SK_IN_R1] = rd.bytes_for(n);
        B.src[SK_IN_R1] = rd.buf();
        g.ck(x, sk_upload(x, slot, SK_IN_R1, B.src[SK_IN_R1], B.bytes[SK_IN_R1]), "sk_upload");
        if (op == OP_ADDBC) {
            B.src[SK_IN_AUX1] = bc.buf();
            if (bc_reuse) {
                memcpy(bc.buf(), last_bc.data(), last_bc.size());
                B.bytes[SK_IN_AUX1] = last_bc.size();
            } else if (nb) {
                B.bytes[SK_IN_AUX1] = bc.bytes_for(nb);
                const size_t s0 = nb >= 2 ? bc.bytes_for(nb - 1) : 0;
                last_bc.assign(bc.buf() + s0, bc.buf() + B.bytes[SK_IN_AUX1]);
            }
            g.ck(x, sk_upload(x, slot, SK_IN_AUX1, B.src[SK_IN_AUX1], B.bytes[SK_IN_AUX1]), "sk_upload");
            // the barcode ring moves in step with the read ring: a buffer stays untouched until its batch is done
            if (nb) bc.consume(nb, B.bytes[SK_IN_AUX1]);
            else bc.park();
        }
        launch(lane, 0);
        rd.consume(n, B.bytes[SK_IN_R1]);
        return true;
    };
    auto fetch_out = [&](int lane, uint64_t n) {
        sk_ctx *x = g.c(lane);
        const uint32_t slot = g.slot(lane);
        uint8_t *h = g.ensure_out(0, lane, n);
        if (n) g.ck(x, sk_download_out(x, slot, 0, h, n), "sk_download_out");
        g.ck(x, sk_wait(x, slot, nullptr), "sk_wait");
        write_all(1, h, n);
    };
    auto complete = [&](int lane) {
        sk_ctx *x = g.c(lane);
        const uint32_t slot = g.slot(lane);
        Batch &B = batches[(size_t)lane];
        sk_result r;
        g.ck(x, sk_wait(x, slot, &r), "sk_wait");
        refuse_status(r, B.first_record);
        if (r.status == SK_DATA_OK) {
            fetch_out(lane, r.out_bytes[0]);
            B.live = false;
            return;
        }
        // A record the reference stops at: everything before it has already been printed (replay).
        uint64_t off = 0;
        if (r.err_record) {
            launch(lane, r.err_record);
            sk_result rep;
            g.ck(x, sk_wait(x, slot, &rep), "sk_wait");
            fetch_out(lane, rep.out_bytes[0]);
            off = rep.consumed[SK_IN_R1];
        }
        const uint8_t *d = B.src[SK_IN_R1];
        const size_t nbytes = B.bytes[SK_IN_R1];
        const uint8_t *nl = (const uint8_t *)memchr(d + off, '\n', nbytes - off);
        const size_t hlen = nl ? (size_t)(nl - (d + off)) + 1 : nbytes - off;
        switch (r.status) {
            case SK_DATA_BAD_HEADER: fatal("Invalid FASTQ format encountered.");  // trim :20-22, mask :21-23
            case SK_DATA_LEN_MISMATCH: fatal("Read sequence and base qualities are of different length.");  // mask :35-37
            case SK_DATA_SEQ_SHORT:  // trim :47 slice panic; the header (:23) is already out
                write_all(1, d + off, hlen);
                panic101("byte index out of range of seq (fasta_trim_by_quality.rs:47)");
            case SK_DATA_BAD_FASTX_LINE: {  // add barcode :33 then :41-43
                write_all(1, d + off, trim_end_len(d + off, hlen));
                write_all(1, (const uint8_t *)" BC:", 4);
                // barcode of iteration err_record: sequence line of that barcode record (or the last one)
                const uint8_t *b = B.src[SK_IN_AUX1];
                const size_t bn = B.bytes[SK_IN_AUX1];
                if (bn && (b[0] == '@' || b[0] == '>')) {
                    const uint32_t lpr = b[0] == '>' ? 2 : 4;
                    std::vector<size_t> ls{0};
                    for (size_t i = 0; i < bn; i++)
                        if (b[i] == '\n' && i + 1 < bn) ls.push_back(i + 1);
                    const size_t nrec = (ls.size() + lpr - 1) / lpr;
                    const size_t i = std::min<size_t>(r.err_record, nrec - 1);
                    const size_t j = i * lpr + 1;
                    if (j < ls.size()) {
                        const size_t e = j + 1 < ls.size() ? ls[j + 1] : bn;
                        write_all(1, b + ls[j], trim_end_len(b + ls[j], e - ls[j]));
                    }
                }
                write_all(1, (const uint8_t *)"\n", 1);
                std::string h((const char *)d + off, hlen);
                fatal("Invalid FASTQ line:\n%s", h.c_str());
            }
            default:
                fprintf(stderr, "seqkit_b200: unexpected data status %d\n", r.status);
                finish(3);
        }
    };

    // Batches are submitted up to NL ahead and completed strictly in order: batch i on lane i % NL.
    uint64_t submitted = 0;
    bool more = true;
    while (more && submitted < (uint64_t)NL) {
        more = submit((int)(submitted % (uint64_t)NL));
        if (more) submitted++;
    }
    while (bi < submitted) {
        complete((int)(bi % (uint64_t)NL));
        bi++;
        if (more) {
            more = submit((int)(submitted % (uint64_t)NL));
            if (more) submitted++;
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// demultiplex (fasta_demultiplex.rs:30-265)
// ------------------------------------------------------------------------------------------------
struct Sample {
    std::string name, barcode;
    Sink *out[2] = {nullptr, nullptr};
    uint64_t total_reads = 0;
};

static bool bc_class(uint8_t c) {
    switch (c) {
        case 'A': case 'C': case 'G': case 'T': case 'N': case 'a': case 'c': case 'g': case 't': case 'n': case '+':
            return true;
    }
    return false;
}
// Regex::find(" BC:[ACGTNacgtn+]+") -- only used to quote the barcode in messages and the dry-run table.
static bool bc_find(const uint8_t *s, size_t n, size_t *st, size_t *en) {
    for (size_t i = 0; i + 5 <= n; i++)
        if (s[i] == ' ' && s[i + 1] == 'B' && s[i + 2] == 'C' && s[i + 3] == ':' && bc_class(s[i + 4])) {
            size_t e = i + 5;
            while (e < n && bc_class(s[e])) e++;
            *st = i;
            *en = e;
            return true;
        }
    return false;
}

static int run_demultiplex(int argc, char **argv) {
    bool parallel = false;
    std::string index1, index2, dry_s;
    bool have_dry = false;
    int fused_trim = -1;  // extension (off by default): --trim-by-quality=Q fuses `fasta trim by quality` into the pass
    std::vector<std::string> pos;
    for (int a = 2; a < argc; a++) {
        std::string s = argv[a];
        auto val = [&](const char *name, std::string &dst) -> bool {
            const size_t k = strlen(name);
            if (s.compare(0, k, name) != 0) return false;
            if (s.size() > k && s[k] == '=') {
                dst = s.substr(k + 1);
                return true;
            }
            if (s.size() == k && a + 1 < argc) {
                dst = argv[++a];
                return true;
            }
            return false;
        };
        std::string tq;
        if (s == "--parallel") parallel = true;
        else if (val("--index1", index1) || val("--index2", index2)) {}
        else if (val("--dry-run", dry_s)) have_dry = true;
        else if (val("--trim-by-quality", tq)) {
            unsigned q;
            if (!parse_u8(tq.c_str(), &q)) invalid_args(USAGE_DEMUX);
            fused_trim = (int)q;
        } else if (s.size() > 1 && s[0] == '-' && s != "-") invalid_args(USAGE_DEMUX);
        else pos.push_back(s);
    }
    if (pos.size() < 2 || pos.size() > 3) invalid_args(USAGE_DEMUX);
    uint64_t dry_run = 0;  // :33-36
    if (have_dry) {
        char *e = nullptr;
        errno = 0;
        dry_run = dry_s.empty() || dry_s[0] == '-' ? 0 : strtoull(dry_s.c_str(), &e, 10);
        if (errno || (e && *e)) dry_run = 0;
        if (dry_run == 0 && !dry_s.empty()) fatal("In --dry-run=N, N must be 64-bit positive integer.");
    }
    const bool paired = pos.size() == 3 && !pos[2].empty();

    // Readers are opened before the sheet is read (:41-55), so "Cannot open file" comes first -- each input
    // exactly once (a FIFO or process substitution cannot be opened twice); the pinned rings come with the
    // contexts below.
    Gpu g;
    Stream st[SK_N_INPUTS];
    Input sheet_in;
    // (open order of the reference: fastq_1, fastq_2, index1, index2, then the sheet)
    std::vector<std::pair<int, std::string>> to_open;
    to_open.push_back({SK_IN_R1, pos[1]});
    if (paired) to_open.push_back({SK_IN_R2, pos[2]});
    if (!index1.empty()) to_open.push_back({SK_IN_AUX1, index1});
    if (!index2.empty()) to_open.push_back({SK_IN_AUX2, index2});
    for (auto &o : to_open) st[o.first].in.open_path(o.second);

    fputs("Reading sample sheet...\n", stderr);  // :58
    sheet_in.open_path(pos[0]);
    std::vector<uint8_t> sheet;
    {
        uint8_t tmp[1 << 16];
        size_t k;
        while ((k = sheet_in.read_some(tmp, sizeof tmp)) > 0) sheet.insert(sheet.end(), tmp, tmp + k);
    }
    for (uint8_t c : sheet)
        if (c >= 0x80) refuse("non-ASCII bytes in the sample sheet");
    std::vector<Sample *> samples;
    size_t barcode_len = 0;
    // SK_GZIP=child: one `gzip -c` (--parallel: `pigz -c`) child per output file, as the reference does it
    const bool child_gzip = getenv("SK_GZIP") && strcmp(getenv("SK_GZIP"), "child") == 0;
    for (size_t p = 0; p < sheet.size();) {  // :63-95
        const uint8_t *nl = (const uint8_t *)memchr(sheet.data() + p, '\n', sheet.size() - p);
        const size_t end = nl ? (size_t)(nl - sheet.data()) + 1 : sheet.size();
        const uint8_t *line = sheet.data() + p;
        size_t n = end - p;
        p = end;
        if (n && line[0] == '#') continue;  // :64
        size_t off = 0;
        while (off < n && is_ws(line[off])) off++;  // line.trim() :65
        const uint8_t *t = line + off;
        const size_t tn = trim_end_len(t, n - off);
        const uint8_t *tab = (const uint8_t *)memchr(t, '\t', tn);
        if (!tab) continue;  // cols.len() < 2 :66
        const size_t name_n = (size_t)(tab - t);
        const uint8_t *c1 = tab + 1;
        const size_t rest = tn - name_n - 1;
        const uint8_t *tab2 = (const uint8_t *)memchr(c1, '\t', rest);
        const size_t bc_n = tab2 ? (size_t)(tab2 - c1) : rest;
        std::string name((const char *)t, name_n);
        if (bc_n == 0) fatal("Sample %s has no barcode.", name.c_str());  // :68
        if (barcode_len == 0) barcode_len = bc_n;                        // :69-73
        else if (bc_n != barcode_len) fatal("Barcodes in sample sheet must all be of same length.");
        Sample *s = new Sample();
        s->name = name;
        s->barcode.assign((const char *)c1, bc_n);
        samples.push_back(s);
        if (dry_run == 0) {  // outputs are created while the sheet is read (:77-87)
            auto make_sink = [&](const std::string &path) -> Sink * {
                if (child_gzip) {
                    ChildSink *k = new ChildSink();
                    k->open_path(path, parallel);
                    return k;
                }
                DeflateSink *k = new DeflateSink();
                k->open_path(path);
                return k;
            };
            if (paired) {
                s->out[0] = make_sink(name + "_1.fq.gz");
                s->out[1] = make_sink(name + "_2.fq.gz");
            } else {
                s->out[0] = make_sink(name + ".fq.gz");
            }
        }
    }
    for (size_t a = 0; a < samples.size(); a++)  // :98-104
        for (size_t b = a + 1; b < samples.size(); b++)
            if (samples[a]->name == samples[b]->name)
                fatal("Sample %s is listed multiple times in sample sheet.", samples[a]->name.c_str());
    fprintf(stderr, "Starting demultiplexing in %s end mode...\n", paired ? "paired" : "single");  // :106-107

    const uint32_t S = (uint32_t)samples.size();
    if (S == 0) refuse("empty sample sheet");
    const bool use_aux = !index1.empty() || !index2.empty();
    Phase *ph_init = new Phase(0);
    uint64_t plain = 0;
    for (auto &o : to_open) plain += st[o.first].in.plain_bytes();
    g.create(S, use_aux, paired ? 2 : 1, false, plain);
    for (auto &o : to_open) g.init_stream(st[o.first], o.second);
    {
        std::string flat;
        for (auto *s : samples) flat += s->barcode;
        for (sk_ctx *xc : g.ctxs) {  // every device holds the sheet
            int rc = sk_set_sheet(xc, (const uint8_t *)flat.data(), S, (uint32_t)barcode_len);
            if (rc == SK_E_UNSUPPORTED) refuse("%s", sk_last_error(xc));
            g.ck(xc, rc, "sk_set_sheet");
        }
    }
    const uint32_t use_index = (index1.empty() ? 0u : 1u) | (index2.empty() ? 0u : 2u);
    const int NL = g.lanes();
    // Per-sample output: one contiguous slice per sample, mate and batch from the device-side compaction
    // (sk_demux_compact).  Sheets beyond its limit (4096 samples) fall back to the per-record slice tables.
    const bool compact = S <= 4096 && !getenv("SK_NO_COMPACT");
    const uint32_t max_chunks = sk_max_chunks(g.ctx);
    std::vector<sk_slice *> slices_h[2];
    std::vector<sk_chunk_row *> rows_h[2];
    std::vector<sk_group *> groups_h[2];
    for (int m = 0; m < (paired ? 2 : 1); m++)
        for (int l = 0; l < NL; l++) {
            if (compact) {
                slices_h[m].push_back((sk_slice *)g.pinned(g.c(l), (uint64_t)(S + 1) * sizeof(sk_slice)));
            } else {
                rows_h[m].push_back((sk_chunk_row *)g.pinned(g.c(l), (uint64_t)max_chunks * sizeof(sk_chunk_row)));
                groups_h[m].push_back((sk_group *)g.pinned(g.c(l), g.max_records * sizeof(sk_group)));
            }
        }
    std::vector<uint64_t> counts_h(S + 2);
    std::vector<sk_event> events;
    std::vector<int16_t> assign_h;
    std::unordered_map<std::string, uint64_t> extra;  // dry run: barcodes matching no sample (:190-194)
    std::vector<std::string> extra_order;
    uint64_t total_reads = 0, identified_reads = 0;
    std::vector<Batch> batches((size_t)NL);
    uint64_t bi = 0;
    unsigned n_threads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));

    auto call = [&](int lane, uint64_t rec_limit) {
        sk_ctx *x = g.c(lane);
        sk_demux_opts o;
        memset(&o, 0, sizeof o);
        o.fused_trim_min_baseq = fused_trim;
        o.use_index = use_index;
        o.rec_limit = rec_limit;
        o.no_output = dry_run ? 1 : 0;
        g.ck(x, sk_demultiplex(x, g.slot(lane), &o), "sk_demultiplex");
        if (compact && !dry_run) g.ck(x, sk_demux_compact(x, g.slot(lane)), "sk_demux_compact");
    };
    auto submit = [&](int lane) -> bool {
        Phase ph(1);
        sk_ctx *x = g.c(lane);
        const uint32_t slot = g.slot(lane);
        Batch &B = batches[(size_t)lane];
        B = Batch();
        if (dry_run && st[SK_IN_R1].records_done >= dry_run) return false;  // :248
        for (auto &o : to_open) st[o.first].top_up();
        size_t n = std::min<size_t>(st[SK_IN_R1].avail(), g.max_records);
        if (n == 0) {
            if (st[SK_IN_R1].drained()) return false;
            refuse("a record does not fit in one batch (%llu bytes); raise SK_BATCH_MB", (unsigned long long)g.batch_bytes);
        }
        for (auto &o : to_open) {
            if (o.first == SK_IN_R1) continue;
            Stream &other = st[o.first];
            if (other.avail() < n) {
                if (other.in.eof) refuse("mate / index files hold fewer records than <fastq_1>");
                n = other.avail();
                if (n == 0) refuse("a record does not fit in one batch; raise SK_BATCH_MB");
            }
        }
        if (dry_run) n = (size_t)std::min<uint64_t>(n, dry_run - st[SK_IN_R1].records_done);
        B.live = true;
        B.n = n;
        B.first_record = st[SK_IN_R1].records_done;
        for (auto &o : to_open) {
            Stream &xs = st[o.first];
            B.bytes[o.first] = xs.bytes_for(n);
            B.src[o.first] = xs.buf();
            g.ck(x, sk_upload(x, slot, o.first, B.src[o.first], B.bytes[o.first]), "sk_upload");
        }
        if (!paired) g.ck(x, sk_set_input_len(x, slot, SK_IN_R2, 0), "sk_set_input_len");
        call(lane, 0);
        for (auto &o : to_open) st[o.first].consume(n, B.bytes[o.first]);
        return true;
    };
    // observed barcode of a record, for messages only
    auto index_barcode = [&](const Batch &B, uint32_t off1, uint32_t off2) {
        std::string bc;
        const int which[2] = {index1.empty() ? SK_IN_AUX2 : SK_IN_AUX1, SK_IN_AUX2};
        const uint32_t offs[2] = {off1, off2};
        for (int q = 0; q < 2; q++) {
            if (offs[q] == 0xFFFFFFFFu) continue;
            const uint8_t *d = B.src[which[q]];
            const size_t nb = B.bytes[which[q]];
            if (offs[q] > nb) continue;
            const uint8_t *nl = (const uint8_t *)memchr(d + offs[q], '\n', nb - offs[q]);
            const size_t len = trim_end_len(d + offs[q], nl ? (size_t)(nl - (d + offs[q])) : nb - offs[q]);
            if (q == 1 && !bc.empty()) bc += '+';
            bc.append((const char *)d + offs[q], len);
        }
        return bc;
    };
    // processes the results of a batch that ran with `r` (possibly a replay limited to the records before an error)
    auto drain = [&](int lane, const sk_result &r) {
        Phase ph(3);
        sk_ctx *x = g.c(lane);
        const uint32_t slot = g.slot(lane);
        Batch &B = batches[(size_t)lane];
        // counters stay on the device: added to the context's run totals, merged over the GPUs at the end (:169,177-178)
        g.ck(x, sk_counts_accumulate(x, slot), "sk_counts_accumulate");
        // WARNING lines in record order (:184-188)
        if (r.flags & SK_FLAG_EVENTS_OVERFLOW) refuse("too many ambiguous reads in one batch");
        events.resize(std::max<uint32_t>(r.n_events, 1));
        const int ne = sk_download_events(x, slot, events.data(), r.n_events);
        for (int k = 0; k < ne; k++) {
            const sk_event &e = events[k];
            std::string bc = use_index ? index_barcode(B, e.bc_off, e.bc_off2)
                                       : std::string((const char *)B.src[SK_IN_R1] + e.bc_off, barcode_len);
            const Sample *a = samples[e.best_sample], *b = samples[e.equally_fine_sample];
            fprintf(stderr,
                    "WARNING: Sequenced barcode %s was an equally good match (%u mismatches) for samples %s (%s) and %s "
                    "(%s), and was therefore not assigned to any sample.\n",
                    bc.c_str(), e.mismatches, a->name.c_str(), a->barcode.c_str(), b->name.c_str(), b->barcode.c_str());
        }
        if (dry_run) {
            // tally the barcodes of reads whose best match is worse than one mismatch (:190-194)
            assign_h.resize(std::max<uint64_t>(r.n_records, 1));
            g.ck(x, sk_download_assign(x, slot, assign_h.data(), r.n_records), "sk_download_assign");
            const uint8_t *d = B.src[SK_IN_R1];
            size_t p = 0;
            std::vector<size_t> ipos(2, 0);
            for (uint64_t i = 0; i < r.n_records; i++) {
                // record i of every stream starts where record i-1 ended: walk four lines
                auto next_record = [](const uint8_t *b, size_t n, size_t &q, size_t &l1, size_t &l1e) {
                    size_t line = 0;
                    const size_t s0 = q;
                    l1 = l1e = s0;
                    while (line < 4 && q < n) {
                        const uint8_t *nl = (const uint8_t *)memchr(b + q, '\n', n - q);
                        const size_t e = nl ? (size_t)(nl - b) + 1 : n;
                        if (line == 0) l1 = e;
                        if (line == 1) l1e = e;
                        q = e;
                        line++;
                    }
                    if (line < 2) l1e = l1;
                    return s0;
                };
                size_t l1, l1e;
                const size_t h0 = next_record(d, B.bytes[SK_IN_R1], p, l1, l1e);
                std::string bc;
                if (use_index) {
                    int k = 0;
                    for (int w : {SK_IN_AUX1, SK_IN_AUX2}) {
                        if (!st[w].active) continue;
                        size_t a1, a1e;
                        next_record(B.src[w], B.bytes[w], ipos[k], a1, a1e);
                        if (!bc.empty()) bc += '+';
                        bc.append((const char *)B.src[w] + a1, trim_end_len(B.src[w] + a1, a1e - a1));
                        k++;
                    }
                } else {
                    size_t a, b;
                    if (bc_find(d + h0, l1 - h0, &a, &b)) bc.assign((const char *)d + h0 + a + 4, b - a - 4);
                }
                if (assign_h[i] == -1) {
                    auto it = extra.find(bc);
                    if (it == extra.end()) {
                        extra.emplace(bc, 1);
                        extra_order.push_back(bc);
                    } else {
                        it->second++;
                    }
                }
            }
            return;
        }
        // per-sample output -> the samples' sinks, in batch order
        const int nm = paired ? 2 : 1;
        if (compact) {
            // one contiguous slice per sample and mate (sk_compact.cu): S appends per mate, not one per record
            for (int m = 0; m < nm; m++) {
                uint8_t *h = g.ensure_out(m, lane, r.out_extent[m]);
                if (r.out_extent[m]) g.ck(x, sk_download_compact(x, slot, (uint32_t)m, h, r.out_extent[m]), "sk_download_compact");
                g.ck(x, sk_download_slices(x, slot, (uint32_t)m, slices_h[m][(size_t)lane]), "sk_download_slices");
            }
            g.ck(x, sk_wait(x, slot, nullptr), "sk_wait");
            for (int m = 0; m < nm; m++) {
                const sk_slice *sl = slices_h[m][(size_t)lane];
                const uint8_t *base = g.out_h[m][(size_t)lane];
                for (uint32_t s = 0; s < S; s++)
                    if (sl[s].len) samples[s]->out[m]->append(base + sl[s].offset, (size_t)sl[s].len);
            }
            return;
        }
        for (int m = 0; m < nm; m++) {
            uint8_t *h = g.ensure_out(m, lane, r.out_extent[m]);
            if (r.out_extent[m]) g.ck(x, sk_download_out(x, slot, (uint32_t)m, h, r.out_extent[m]), "sk_download_out");
            g.ck(x, sk_download_demux_tables(x, slot, (uint32_t)m, rows_h[m][(size_t)lane], groups_h[m][(size_t)lane], r.n_records),
                 "sk_download_demux_tables");
        }
        g.ck(x, sk_wait(x, slot, nullptr), "sk_wait");
        // Bucket the groups by sample (one pass over the chunk rows), then append each sample's groups in chunk order.
        struct Piece {
            uint64_t off;
            uint32_t len;
        };
        std::vector<uint32_t> cnt(S + 1);
        std::vector<Piece> pieces;
        for (int m = 0; m < nm; m++) {
            const uint32_t nc = r.n_chunks[m];
            const sk_chunk_row *rows = rows_h[m][(size_t)lane];
            const sk_group *groups = groups_h[m][(size_t)lane];
            std::fill(cnt.begin(), cnt.end(), 0u);
            for (uint32_t c = 0; c < nc; c++)
                for (uint32_t k = 0; k < rows[c].n_groups; k++) cnt[groups[rows[c].first_group + k].sample + 1]++;
            for (uint32_t s = 0; s < S; s++) cnt[s + 1] += cnt[s];
            pieces.resize(cnt[S]);
            std::vector<uint32_t> cur(cnt.begin(), cnt.end() - 1);
            for (uint32_t c = 0; c < nc; c++) {
                uint64_t off = rows[c].base;
                for (uint32_t k = 0; k < rows[c].n_groups; k++) {
                    const sk_group &gr = groups[rows[c].first_group + k];
                    pieces[cur[gr.sample]++] = Piece{off, gr.len};
                    off += gr.len;
                }
            }
            std::atomic<uint32_t> next{0};
            auto worker = [&]() {
                std::vector<uint8_t> tmp;
                for (;;) {
                    const uint32_t s = next.fetch_add(1);
                    if (s >= S) break;
                    tmp.clear();
                    for (uint32_t k = cnt[s]; k < cnt[s + 1]; k++) {
                        const uint8_t *src = g.out_h[m][(size_t)lane] + pieces[k].off;
                        tmp.insert(tmp.end(), src, src + pieces[k].len);
                    }
                    if (!tmp.empty()) samples[s]->out[m]->append(tmp.data(), tmp.size());
                }
            };
            std::vector<std::thread> pool;
            const unsigned nt = std::min<unsigned>(n_threads, S);
            for (unsigned t = 1; t < nt; t++) pool.emplace_back(worker);
            worker();
            for (auto &t : pool) t.join();
        }
    };
    // Run totals: one grouped NCCL all-reduce over the GPUs' device-side counters, then a single download.
    auto finish_counts = [&]() {
        // NCCL writes its version banner (and, with NCCL_DEBUG, its log) to stdout, which belongs to the reference's
        // output: the descriptor points at /dev/null while the collective runs (failures come back as codes)
        fflush(stdout);
        const int saved = g.G > 1 ? dup(1) : -1;
        if (saved >= 0) {
            const int nul = open("/dev/null", O_WRONLY | O_CLOEXEC);
            if (nul >= 0) {
                dup2(nul, 1);
                close(nul);
            }
        }
        const int arc = sk_allreduce_totals(g.ctxs.data(), g.G);
        if (saved >= 0) {
            fflush(stdout);  // what the library left in the stdio buffer goes to /dev/null too
            dup2(saved, 1);
            close(saved);
        }
        g.ck(arc, "sk_allreduce_totals");
        g.ck(sk_download_totals(g.ctx, counts_h.data()), "sk_download_totals");
        for (uint32_t s = 0; s < S; s++) samples[s]->total_reads = counts_h[s];
        total_reads = counts_h[S];
        identified_reads = counts_h[S + 1];
    };
    auto complete = [&](int lane) {
        sk_ctx *x = g.c(lane);
        const uint32_t slot = g.slot(lane);
        Batch &B = batches[(size_t)lane];
        sk_result r;
        {
            Phase ph(2);
            g.ck(x, sk_wait(x, slot, &r), "sk_wait");
        }
        refuse_status(r, B.first_record);
        if (r.flags & SK_FLAG_MATE_COUNT) refuse("mate / index files hold fewer records than <fastq_1>");
        if (r.status == SK_DATA_OK) {
            drain(lane, r);
            B.live = false;
            return;
        }
        sk_result rep;
        memset(&rep, 0, sizeof rep);
        if (r.err_record) {
            call(lane, r.err_record);
            g.ck(x, sk_wait(x, slot, &rep), "sk_wait");
            drain(lane, rep);
        }
        const uint8_t *d = B.src[SK_IN_R1];
        const size_t nbytes = B.bytes[SK_IN_R1];
        const size_t off = r.err_record ? rep.consumed[SK_IN_R1] : 0;
        const uint8_t *nl = (const uint8_t *)memchr(d + off, '\n', nbytes - off);
        const size_t hlen = nl ? (size_t)(nl - (d + off)) + 1 : nbytes - off;
        std::string hdr((const char *)d + off, hlen);
        switch (r.status) {
            case SK_DATA_BAD_HEADER: fatal("Invalid FASTQ header line:\n%s", hdr.c_str());  // :118-120
            case SK_DATA_NO_BC: fatal("No BC:xxxx field found.");                           // :141
            case SK_DATA_BC_LEN: {                                                          // :148-150
                std::string bc;
                if (use_index) {
                    uint32_t offs2[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
                    int k = index1.empty() ? 1 : 0;
                    for (int w : {SK_IN_AUX1, SK_IN_AUX2}) {
                        if (!st[w].active) continue;
                        const size_t o = r.err_record ? rep.consumed[w] : 0;
                        const uint8_t *b = B.src[w];
                        const uint8_t *e = (const uint8_t *)memchr(b + o, '\n', B.bytes[w] - o);
                        offs2[k++] = e ? (uint32_t)(e - b) + 1 : (uint32_t)B.bytes[w];
                    }
                    bc = index_barcode(B, offs2[0], offs2[1]);
                } else {
                    size_t a, b;
                    if (bc_find((const uint8_t *)hdr.data(), hdr.size(), &a, &b)) bc = hdr.substr(a + 4, b - a - 4);
                }
                fatal("Sequenced barcode %s is of different length (%zu nt) than barcodes in the sample sheet (%zu nt).",
                      bc.c_str(), bc.size(), barcode_len);
            }
            case SK_DATA_INDEX_ASSERT: panic101("assertion failed (fasta_demultiplex.rs:130/134)");
            case SK_DATA_SEQ_SHORT: refuse("fused trim: sequence shorter than the kept quality prefix");
            default:
                fprintf(stderr, "seqkit_b200: unexpected data status %d\n", r.status);
                finish(3);
        }
    };

    delete ph_init;
    // Batches are submitted up to NL ahead (round-robin over the GPUs) and completed strictly in order.
    uint64_t submitted = 0;
    bool more = true;
    while (more && submitted < (uint64_t)NL) {
        more = submit((int)(submitted % (uint64_t)NL));
        if (more) submitted++;
    }
    while (bi < submitted) {
        complete((int)(bi % (uint64_t)NL));
        bi++;
        if (more) {
            more = submit((int)(submitted % (uint64_t)NL));
            if (more) submitted++;
        }
    }
    finish_counts();

    if (dry_run) {  // :251-261
        fflush(stdout);
        fprintf(stderr, "Dry run completed with %llu clusters. Barcodes found:\n", (unsigned long long)total_reads);
        std::vector<std::pair<std::string, uint64_t>> entries;
        for (auto *s : samples) entries.push_back({s->name, s->total_reads});
        for (auto &k : extra_order) entries.push_back({k, extra[k]});  // the reference's HashMap order is arbitrary
        std::stable_sort(entries.begin(), entries.end(),
                         [](const std::pair<std::string, uint64_t> &a, const std::pair<std::string, uint64_t> &b) {
                             return a.second < b.second;
                         });
        std::reverse(entries.begin(), entries.end());
        if (entries.size() < 100) panic101("range end index 100 out of range for slice (fasta_demultiplex.rs:258)");
        for (size_t q = 0; q < 100; q++) printf("- %s: %llu\n", entries[q].first.c_str(), (unsigned long long)entries[q].second);
        fflush(stdout);
    }
    if (total_reads == 0)
        fprintf(stderr, "%llu / %llu (NaN%%) clusters carried a barcode matching one of the provided samples.\n",
                (unsigned long long)identified_reads, (unsigned long long)total_reads);
    else
        fprintf(stderr, "%llu / %llu (%.1f%%) clusters carried a barcode matching one of the provided samples.\n",
                (unsigned long long)identified_reads, (unsigned long long)total_reads,
                (double)identified_reads / (double)total_reads * 100.0);  // :263-264
    return 0;
}


// ------------------------------------------------------------------------------------------------
// trim --first/--last, check, statistics, interleave, deinterleave, extract dual umi
// (fasta_trim.rs, fasta_check.rs, fasta_statistics.rs, fasta_interleave.rs, fasta_deinterleave.rs,
// fasta_extract_dual_umi.rs) on the line engine (sk_line_op)
// ------------------------------------------------------------------------------------------------
static const char *USAGE_TRIMFIX =
    "\nUsage:\n  fasta trim [options] <fastq_file>\n\nOptions:\n"
    "  --first=N          Remove first N bases of each read [default: 0].\n"
    "  --last=N           Remove last N bases of each read [default: 0].\n";
static const char *USAGE_CHECK =
    "\nUsage:\n  fasta check <fasta/fastq>\n\nDescription:\nChecks that the input FASTA or FASTQ file is correctly formatted, and reports\n"
    "the line number if any malformatted lines are found.\n";
static const char *USAGE_STATS = "\nUsage:\n  fasta statistics <fastq_file>\n";
static const char *USAGE_INTERLEAVE = "\nUsage:\n  fasta interleave <fastq_1> <fastq_2>\n";
static const char *USAGE_DEINTERLEAVE = "\nUsage:\n  fasta deinterleave <interleaved_fastq> <out_prefix>\n";
static const char *USAGE_DUALUMI =
    "\nUsage:\n  fasta extract dual umi [options] <interleaved_fastq>\n\nOptions:\n"
    "  --first-bases=N   First N bases of read contain UMI bases [default: 0]\n";

// usize::from_str: optional '+', decimal digits
static bool parse_usize(const char *s, uint64_t *v) {
    if (*s == '+') s++;
    if (!*s) return false;
    uint64_t x = 0;
    for (; *s; s++) {
        if (*s < '0' || *s > '9') return false;
        if (x > (~0ull - 9) / 10) return false;
        x = x * 10 + (uint64_t)(*s - '0');
    }
    *v = x;
    return true;
}
// bytes [0, e) of d that hold the next `lines` lines from offset `from` (fewer at the end of the data)
static size_t skip_lines(const uint8_t *d, size_t n, size_t from, unsigned lines) {
    size_t p = from;
    for (unsigned k = 0; k < lines && p < n; k++) {
        const uint8_t *nl = (const uint8_t *)memchr(d + p, '\n', n - p);
        p = nl ? (size_t)(nl - d) + 1 : n;
    }
    return p;
}

static int run_line_op(uint32_t op, const Input &in_a, const Input &in_b, uint64_t x, uint64_t y, const std::string &out_prefix) {
    const bool two_out = op == SK_LOP_DEINTERLEAVE, two_in = op == SK_LOP_INTERLEAVE;
    const bool pairs = op == SK_LOP_DEINTERLEAVE || op == SK_LOP_DUAL_UMI;
    Sink *gz[2] = {nullptr, nullptr};
    if (two_out) {  // GzipWriter::with_method creates both files before anything is read (fasta_deinterleave.rs:17-20)
        const bool child_gzip = getenv("SK_GZIP") && strcmp(getenv("SK_GZIP"), "child") == 0;
        for (int m = 0; m < 2; m++) {
            const std::string path = out_prefix + (m == 0 ? "_1.fq.gz" : "_2.fq.gz");
            if (child_gzip) {
                ChildSink *k = new ChildSink();
                k->open_path(path, false);
                gz[m] = k;
            } else {
                DeflateSink *k = new DeflateSink();
                k->open_path(path);
                gz[m] = k;
            }
        }
    }
    Gpu g;
    g.create(0, false, two_out ? 2 : 1, true, in_a.plain_bytes() + (two_in ? in_b.plain_bytes() : 0));
    Stream sa, sb;
    sa.in = in_a;
    g.init_stream(sa, "");
    if (two_in) {
        sb.in = in_b;
        g.init_stream(sb, "");
    }
    const int NL = g.lanes();
    std::vector<Batch> batches((size_t)NL);
    bool first = true;
    unsigned lpr = 4;        // lines per record of the data: '@' 4, '>' 2
    uint64_t total_records = 0;
    std::unordered_map<std::string, uint64_t> barcodes;  // statistics
    std::vector<sk_stat_entry> ents;
    std::vector<std::string> tail_lines;  // check: the last ten lines before the current batch (fasta_check.rs:32-35)
    uint64_t lines_before = 0;            // ... and their number

    auto submit = [&](int lane) -> bool {
        sk_ctx *xc = g.c(lane);
        const uint32_t slot = g.slot(lane);
        Batch &B = batches[(size_t)lane];
        B = Batch();
        sa.top_up();
        if (first) {
            first = false;
            if (sa.fill && sa.buf()[0] == '>') lpr = 2;
            const unsigned unit = pairs ? 2 * lpr : lpr;  // batches are cut at whole records (pairs)
            for (Stream *st : {&sa, &sb}) {
                if (!st->active) continue;
                st->lpr = st == &sb ? lpr : unit;
                st->rec_ends.clear();
                st->scanned = 0;
                st->lines_mod = 0;
                st->top_up();
            }
        }
        size_t n = std::min<size_t>(sa.avail(), g.max_records / 2);
        if (n == 0) {
            if (sa.drained()) return false;
            refuse("a record does not fit in one batch (%llu bytes); raise SK_BATCH_MB", (unsigned long long)g.batch_bytes);
        }
        size_t nb = 0;
        if (two_in) {  // the second file may run out early: the reference then fails at that record (:29-33)
            sb.top_up();
            nb = std::min(n, sb.avail());
            if (nb < n && !sb.in.eof) n = nb;
            if (n == 0) refuse("a record does not fit in one batch; raise SK_BATCH_MB");
            nb = std::min(n, sb.avail());
        }
        B.live = true;
        B.n = n;
        B.first_record = sa.records_done;
        B.bytes[SK_IN_R1] = sa.bytes_for(n);
        B.src[SK_IN_R1] = sa.buf();
        g.ck(xc, sk_upload(xc, slot, SK_IN_R1, B.src[SK_IN_R1], B.bytes[SK_IN_R1]), "sk_upload");
        if (two_in) {
            B.bytes[SK_IN_R2] = sb.bytes_for(nb);
            B.src[SK_IN_R2] = sb.buf();
            g.ck(xc, sk_upload(xc, slot, SK_IN_R2, B.src[SK_IN_R2], B.bytes[SK_IN_R2]), "sk_upload");
        } else {
            g.ck(xc, sk_set_input_len(xc, slot, SK_IN_R2, 0), "sk_set_input_len");
        }
        g.ck(xc, sk_line_op(xc, slot, op, (uint32_t)std::min<uint64_t>(x, 0xFFFFFFFFull), (uint32_t)std::min<uint64_t>(y, 0xFFFFFFFFull), 0),
             "sk_line_op");
        sa.consume(n, B.bytes[SK_IN_R1]);
        if (two_in) {
            if (nb) sb.consume(nb, B.bytes[SK_IN_R2]);
            else sb.park();
        }
        return true;
    };
    auto complete = [&](int lane) {
        sk_ctx *xc = g.c(lane);
        const uint32_t slot = g.slot(lane);
        Batch &B = batches[(size_t)lane];
        sk_result r;
        g.ck(xc, sk_wait(xc, slot, &r), "sk_wait");
        refuse_status(r, B.first_record);
        const uint8_t *d = B.src[SK_IN_R1];
        const size_t nbytes = B.bytes[SK_IN_R1];
        // output of the records before the failing one (all of them when nothing failed)
        if (op == SK_LOP_TRIM || op == SK_LOP_INTERLEAVE || op == SK_LOP_DUAL_UMI) {
            uint8_t *h = g.ensure_out(0, lane, r.out_bytes[0]);
            if (r.out_bytes[0]) g.ck(xc, sk_download_out(xc, slot, 0, h, r.out_bytes[0]), "sk_download_out");
            g.ck(xc, sk_wait(xc, slot, nullptr), "sk_wait");
            write_all(1, h, r.out_bytes[0]);
        } else if (op == SK_LOP_DEINTERLEAVE) {
            for (int m = 0; m < 2; m++) {
                uint8_t *h = g.ensure_out(m, lane, r.out_bytes[m]);
                if (r.out_bytes[m]) g.ck(xc, sk_download_out(xc, slot, (uint32_t)m, h, r.out_bytes[m]), "sk_download_out");
            }
            g.ck(xc, sk_wait(xc, slot, nullptr), "sk_wait");
            for (int m = 0; m < 2; m++) gz[m]->append(g.out_h[m][(size_t)lane], r.out_bytes[m]);
        } else if (op == SK_LOP_STATS) {
            uint32_t ne = 0;
            sk_download_stats(xc, slot, nullptr, 0, &ne);
            ents.resize(std::max<uint32_t>(ne, 1));
            g.ck(xc, sk_download_stats(xc, slot, ents.data(), ne, &ne), "sk_download_stats");
            for (uint32_t k = 0; k < ne; k++) barcodes[std::string((const char *)d + ents[k].off, ents[k].len)] += ents[k].count;
        }
        total_records += r.n_records;
        const size_t off = (size_t)r.consumed[SK_IN_R1];  // where the failing record (pair) starts
        const unsigned unit = pairs ? 2 * lpr : lpr;
        if (r.status != SK_DATA_OK) {
            const size_t hend = skip_lines(d, nbytes, off, 1);
            const std::string hdr((const char *)d + off, hend - off);
            switch (r.status) {
                case SK_DATA_BAD_HEADER:
                    if (op == SK_LOP_TRIM) fatal("Invalid FASTA/FASTQ format encountered.");                 // fasta_trim.rs:28-30
                    if (op == SK_LOP_STATS) fatal("Invalid FASTQ header:\n%s", hdr.c_str());                 // fasta_statistics.rs:36-38
                    if (op == SK_LOP_INTERLEAVE || op == SK_LOP_DEINTERLEAVE) fatal("Line is not FASTA/FASTQ format: %s", hdr.c_str());
                    if (op == SK_LOP_DUAL_UMI) fatal("Header is not valid FASTA/FASTQ:\n%s", hdr.c_str());   // :33-35
                    break;  // check: below
                case SK_DATA_QUAL_SHORT:
                    if (op == SK_LOP_TRIM) {  // the first print! of the record is out when &qual[..] panics (fasta_trim.rs:35,:44)
                        const size_t s1 = skip_lines(d, nbytes, off, 2);
                        const size_t sl = trim_end_len(d + hend, s1 - hend);
                        write_all(1, d + off, hend - off);
                        write_all(1, d + hend + x, sl - y - x);
                        write_all(1, (const uint8_t *)"\n", 1);
                    }
                    panic101("byte index out of range of qual");
                case SK_DATA_SEQ_SHORT: panic101("byte index out of range of seq (fasta_extract_dual_umi.rs:56-58)");
                case SK_DATA_INCONSISTENT:
                    if (op == SK_LOP_INTERLEAVE) {  // the record of <fastq_1> is out before <fastq_2>'s header is looked at (:24-29)
                        write_all(1, d + off, skip_lines(d, nbytes, off, lpr) - off);
                        fatal("Input files do not share a consistent format.");
                    }
                    if (op == SK_LOP_DEINTERLEAVE) {  // :26-35
                        gz[0]->append(d + off, skip_lines(d, nbytes, off, lpr) - off);
                        fatal("Interleaved FASTA records are not in consistent format.");
                    }
                    fatal(lpr == 4 ? "Invalid FASTQ record found in input file." : "Invalid FASTA record found in input file.");
                default: break;
            }
            if (op == SK_LOP_CHECK) {  // fasta_check.rs:58-66: the line number and the last ten lines read
                // lines of this batch up to and including the offending one
                std::vector<std::string> hist = tail_lines;
                uint64_t lines_read = lines_before;
                const size_t upto = r.status == SK_DATA_NO_PLUS ? skip_lines(d, nbytes, off, 3) : hend;
                for (size_t p = 0; p < upto;) {
                    const size_t e = skip_lines(d, nbytes, p, 1);
                    hist.push_back(std::string((const char *)d + p, e - p));
                    if (hist.size() > 10) hist.erase(hist.begin());
                    lines_read++;
                    p = e;
                }
                std::string msg = r.status == SK_DATA_NO_PLUS ? "Missing quality header prefix '+'" : "Missing header prefix '>' or '@'";
                msg += " on line " + std::to_string(lines_read) + ":\n";
                for (auto &l : hist) msg += l + "\n";
                msg += "\n";
                fflush(stdout);
                fputs("ERROR: ", stderr);
                fwrite(msg.data(), 1, msg.size(), stderr);
                fputc('\n', stderr);
                finish(255);
            }
            fprintf(stderr, "seqkit_b200: unexpected data status %d\n", r.status);
            finish(3);
        }
        if (op == SK_LOP_CHECK) {  // remember the batch's last ten lines
            size_t p = nbytes, cnt = 0;
            std::vector<std::string> last;
            while (p > 0 && cnt < 10) {
                size_t q = p - 1;  // start of the line that ends at p
                while (q > 0 && d[q - 1] != '\n') q--;
                last.insert(last.begin(), std::string((const char *)d + q, p - q));
                p = q;
                cnt++;
            }
            for (auto &l : last) {
                tail_lines.push_back(l);
                if (tail_lines.size() > 10) tail_lines.erase(tail_lines.begin());
            }
            lines_before += (uint64_t)r.n_lines[SK_IN_R1];
        }
        (void)unit;
        B.live = false;
    };

    uint64_t submitted = 0, bi = 0;
    bool more = true;
    while (more && submitted < (uint64_t)NL) {
        more = submit((int)(submitted % (uint64_t)NL));
        if (more) submitted++;
    }
    while (bi < submitted) {
        complete((int)(bi % (uint64_t)NL));
        bi++;
        if (more) {
            more = submit((int)(submitted % (uint64_t)NL));
            if (more) submitted++;
        }
    }
    if (op == SK_LOP_STATS) {  // fasta_statistics.rs:43-52
        printf("Total sequence records: %llu\n", (unsigned long long)total_records);
        printf("Most frequent sample barcodes:\n");
        fflush(stdout);
        std::vector<std::pair<std::string, uint64_t>> entries(barcodes.begin(), barcodes.end());
        // count descending; the reference leaves the order of equal counts to its HashMap: barcode descending here (and in the oracle)
        std::sort(entries.begin(), entries.end(), [](const std::pair<std::string, uint64_t> &a, const std::pair<std::string, uint64_t> &b) {
            if (a.second != b.second) return a.second > b.second;
            return a.first > b.first;
        });
        if (entries.size() < 100) panic101("range end index 100 out of range for slice (fasta_statistics.rs:50)");
        for (size_t q = 0; q < 100; q++) printf("- %s: %llu\n", entries[q].first.c_str(), (unsigned long long)entries[q].second);
        fflush(stdout);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// dispatcher (fasta_main.rs:42-82)
// ------------------------------------------------------------------------------------------------
int main(int argc, char **argv) {
    signal(SIGPIPE, SIG_IGN);
    auto is = [&](int i, const char *w) { return i < argc && strcmp(argv[i], w) == 0; };
    int rc = 0;
    if (argc >= 4 && is(1, "trim") && is(2, "by") && is(3, "quality")) {
        if (argc != 6 || argv[4][0] == '\0') invalid_args(USAGE_TRIM);
        unsigned q;
        Input in, none;  // FileReader::new precedes the parse of <min_baseq> (fasta_trim_by_quality.rs:12-13)
        in.open_path(argv[4]);
        if (!parse_u8(argv[5], &q)) panic101("min_baseq parse");
        rc = run_stream_op(OP_TRIM, in, none, q);
    } else if (argc >= 4 && is(1, "mask") && is(2, "by") && is(3, "quality")) {
        if (argc != 6) invalid_args(USAGE_MASK);
        unsigned q;
        Input in, none;
        in.open_path(argv[4]);
        if (!parse_u8(argv[5], &q)) panic101("min_baseq parse");
        rc = run_stream_op(OP_MASK, in, none, q);
    } else if (argc >= 3 && is(1, "add") && is(2, "barcode")) {
        if (argc != 5) invalid_args(USAGE_ADDBC);
        Input in, bcin;
        in.open_path(argv[3]);
        bcin.open_path(argv[4]);
        rc = run_stream_op(OP_ADDBC, in, bcin, 0);
    } else if (argc >= 2 && (is(1, "trim") || is(1, "check") || is(1, "statistics") || is(1, "interleave") || is(1, "deinterleave") ||
                             (argc >= 4 && is(1, "extract") && is(2, "dual") && is(3, "umi")))) {
        const bool umi = is(1, "extract");
        const uint32_t op = is(1, "trim") ? SK_LOP_TRIM : is(1, "check") ? SK_LOP_CHECK : is(1, "statistics") ? SK_LOP_STATS
                            : is(1, "interleave") ? SK_LOP_INTERLEAVE : is(1, "deinterleave") ? SK_LOP_DEINTERLEAVE : SK_LOP_DUAL_UMI;
        const char *usage = op == SK_LOP_TRIM ? USAGE_TRIMFIX : op == SK_LOP_CHECK ? USAGE_CHECK : op == SK_LOP_STATS ? USAGE_STATS
                            : op == SK_LOP_INTERLEAVE ? USAGE_INTERLEAVE : op == SK_LOP_DEINTERLEAVE ? USAGE_DEINTERLEAVE : USAGE_DUALUMI;
        std::vector<std::string> pos;
        std::string v_first = "0", v_last = "0", v_fb = "0";
        for (int a = umi ? 4 : 2; a < argc; a++) {
            const std::string s = argv[a];
            auto opt = [&](const char *name, std::string &dst) -> bool {
                const size_t k = strlen(name);
                if (s.compare(0, k, name) != 0) return false;
                if (s.size() > k && s[k] == '=') {
                    dst = s.substr(k + 1);
                    return true;
                }
                if (s.size() == k && a + 1 < argc) {
                    dst = argv[++a];
                    return true;
                }
                return false;
            };
            if (op == SK_LOP_TRIM && (opt("--first", v_first) || opt("--last", v_last))) continue;
            if (op == SK_LOP_DUAL_UMI && opt("--first-bases", v_fb)) continue;
            if (s.size() > 1 && s[0] == '-') invalid_args(usage);
            pos.push_back(s);
        }
        const size_t want = (op == SK_LOP_INTERLEAVE || op == SK_LOP_DEINTERLEAVE) ? 2 : 1;
        if (pos.size() != want) invalid_args(usage);
        Input a, b;  // FileReader::new comes before the options are parsed (fasta_trim.rs:17-22)
        a.open_path(pos[0]);
        if (op == SK_LOP_INTERLEAVE) b.open_path(pos[1]);
        uint64_t x = 0, y = 0;
        if (op == SK_LOP_TRIM) {
            if (!parse_usize(v_first.c_str(), &x)) fatal("N must be a non-negative integer in --first=N.");
            if (!parse_usize(v_last.c_str(), &y)) fatal("N must be a non-negative integer in --last=N.");
        } else if (op == SK_LOP_DUAL_UMI) {
            if (!parse_usize(v_fb.c_str(), &x)) fatal("N must be a non-negative integer in --first-bases=N.");
        }
        rc = run_line_op(op, a, b, x, y, op == SK_LOP_DEINTERLEAVE ? pos[1] : std::string());
    } else if (argc >= 2 && is(1, "demultiplex")) {
        rc = run_demultiplex(argc, argv);
    } else {
        // Subcommands outside the per-read batch path are not part of this drop-in.
        fprintf(stderr, "%s\n", TOP_USAGE);
        rc = 0;
    }
    finish(rc);
}
