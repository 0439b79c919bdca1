// pgb_export.cu — host pipeline behind pgb_export_gt_vcf / pgb_export_gt_vcf_mem.
//
// Replaces the body loop of Pfile::output_vcf (/root/reference/src/pfile.rs:149-192):
// instead of one lseek+read per variant and two buffered writes per genotype, the kept
// variants are cut into contiguous chunks; each chunk's records, prefixes and row indices
// are staged in page-locked memory, copied H2D in one transfer, indexed (K1), formatted
// (K2) and copied D2H into a page-locked ring that a writer thread drains to the sink.
// Contiguous chunk ranges are sharded over the requested devices; no collective is needed
// because line offsets are a closed-form function of the prefix offsets and K.
//
// There is no CPU formatting path in this file: without a CUDA device the calls fail with
// PGB_E_NO_DEVICE.
#include <cuda_runtime.h>
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/vfs.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "pgb_internal.h"

namespace {

constexpr int kSlots = 3;
constexpr uint64_t kAlign = 256;

inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

uint64_t env_u64(const char *name, uint64_t dflt) {
    const char *s = getenv(name);
    if (!s || !*s) return dflt;
    char *e = nullptr;
    unsigned long long v = strtoull(s, &e, 0);
    return (e && e != s) ? (uint64_t)v : dflt;
}

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            pgb_set_error("%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__);     \
            return e__ == cudaErrorMemoryAllocation ? PGB_E_NOMEM : PGB_E_CUDA;                      \
        }                                                                                            \
    } while (0)

struct Slot {
    uint8_t *h_in = nullptr, *d_in = nullptr;
    uint64_t cap_in = 0;
    uint8_t *h_out = nullptr, *d_out = nullptr;
    uint64_t cap_out = 0;
    pgb_line_meta *d_meta = nullptr;
    void *d_scratch = nullptr;
    uint64_t cap_lines = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_done = nullptr;
    bool busy = false; // guarded by DeviceCtx::mu
};

struct DeviceCtx {
    int device = -1;
    Slot slot[kSlots];
    uint8_t *d_mask = nullptr;
    uint8_t *h_mask = nullptr;
    uint32_t *d_kidx = nullptr;
    uint32_t *d_count = nullptr;
    uint64_t cap_mask = 0;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    std::mutex mu;
    std::condition_variable cv;
};

} // namespace

struct pgb_file {
    int fd = -1;
    const uint8_t *image = nullptr;
    bool image_pinned = false;
    uint64_t bytes = 0; // file or image size
    uint32_t M = 0, N = 0, R = 0;
    // standard-format (.pgen storage mode 0x10) handles: the per-variant record index from the
    // header walk (pgb_pgen10_index, src/pgen.rs:100-258); empty for fixed-width mode 0x02 files
    bool standard = false;
    std::vector<uint64_t> off10;
    std::vector<uint8_t> type10;
    std::vector<uint32_t> len10;
    std::mutex mu; // one export at a time per handle
    // file offset of variant v's record: pfile.rs:165 in u64, or the mode-0x10 index
    uint64_t rec_off(uint64_t v) const { return standard ? off10[v] : pgb_record_offset(v, R); }
};

namespace {

int parse_header(const uint8_t *h, size_t got, pgb_file *f) {
    // Order of checks follows Pfile::from_prefix, pfile.rs:44-69.
    if (got < 2) { pgb_set_error("short read of .pgen header"); return PGB_E_IO; }
    if (h[0] != 0x6C || h[1] != 0x1B) return PGB_E_MAGIC;
    if (got < 3) { pgb_set_error("short read of .pgen header"); return PGB_E_IO; }
    if (h[2] != 0x02) { pgb_set_error("storage mode 0x%02x", h[2]); return PGB_E_MODE; }
    if (got < 12) { pgb_set_error("short read of .pgen header"); return PGB_E_IO; }
    f->M = (uint32_t)h[3] | (uint32_t)h[4] << 8 | (uint32_t)h[5] << 16 | (uint32_t)h[6] << 24;
    f->N = (uint32_t)h[7] | (uint32_t)h[8] << 8 | (uint32_t)h[9] << 16 | (uint32_t)h[10] << 24;
    if (h[11] != 0x40) { pgb_set_error("header byte 11 = 0x%02x", h[11]); return PGB_E_FLAGS; }
    f->R = pgb_record_bytes(f->N);
    return PGB_OK;
}

void free_slot(Slot &s) {
    if (s.h_in) cudaFreeHost(s.h_in);
    if (s.d_in) cudaFree(s.d_in);
    if (s.h_out) cudaFreeHost(s.h_out);
    if (s.d_out) cudaFree(s.d_out);
    if (s.d_meta) cudaFree(s.d_meta);
    if (s.d_scratch) cudaFree(s.d_scratch);
    if (s.ev_k0) cudaEventDestroy(s.ev_k0);
    if (s.ev_k1) cudaEventDestroy(s.ev_k1);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = Slot();
}

void free_ctx(DeviceCtx *c) {
    if (!c) return;
    if (cudaSetDevice(c->device) == cudaSuccess) {
        for (auto &s : c->slot) free_slot(s);
        if (c->d_mask) cudaFree(c->d_mask);
        if (c->h_mask) cudaFreeHost(c->h_mask);
        if (c->d_kidx) cudaFree(c->d_kidx);
        if (c->d_count) cudaFree(c->d_count);
        if (c->ev_a) cudaEventDestroy(c->ev_a);
        if (c->ev_b) cudaEventDestroy(c->ev_b);
    }
    delete c;
}

// Per-device buffers (page-locked staging, device slots, streams, events) are process-wide
// and outlive the pgb_file handles: a CLI-style host that opens, exports and closes per call
// would otherwise pay ~30 ms of cudaHostAlloc/cudaMalloc on every export.  An export owns the
// contexts of its devices for its duration (in_use), so concurrent exports on one device
// queue up.  pgb_release_buffers() frees everything that is idle.
struct CtxRegistry {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<DeviceCtx *> all;
    std::vector<bool> in_use;
} g_registry;

// Acquires (creating on first use) the contexts of `devs` in ascending device order.
void acquire_ctxs(const std::vector<int> &devs, std::vector<DeviceCtx *> *out) {
    std::vector<int> order(devs);
    std::sort(order.begin(), order.end());
    std::unique_lock<std::mutex> lk(g_registry.mu);
    out->assign(devs.size(), nullptr);
    for (int d : order) {
        size_t idx = g_registry.all.size();
        for (size_t i = 0; i < g_registry.all.size(); i++)
            if (g_registry.all[i]->device == d) idx = i;
        if (idx == g_registry.all.size()) {
            DeviceCtx *c = new DeviceCtx();
            c->device = d;
            g_registry.all.push_back(c);
            g_registry.in_use.push_back(false);
        }
        g_registry.cv.wait(lk, [&] { return !g_registry.in_use[idx]; });
        g_registry.in_use[idx] = true;
        for (size_t g = 0; g < devs.size(); g++)
            if (devs[g] == d) (*out)[g] = g_registry.all[idx];
    }
}

void release_ctxs(const std::vector<DeviceCtx *> &ctxs) {
    {
        std::lock_guard<std::mutex> lk(g_registry.mu);
        for (DeviceCtx *c : ctxs)
            for (size_t i = 0; i < g_registry.all.size(); i++)
                if (g_registry.all[i] == c) g_registry.in_use[i] = false;
    }
    g_registry.cv.notify_all();
}

struct CtxLease {
    std::vector<DeviceCtx *> ctxs;
    ~CtxLease() { release_ctxs(ctxs); }
};

struct Chunk {
    uint64_t a, b;       // line range [a, b)
    uint64_t out_off;    // offset of the chunk in the body
    uint64_t out_bytes;
    uint64_t o0;         // file offset of the first staged record byte (dense)
    bool dense;          // true: stage the covering byte range; false: stage kept rows compactly
    uint64_t in_bytes;   // record bytes staged
    uint32_t max_pfx;
    // rows mode (prefixes built on the device from raw .pvar rows): the staged text
    uint64_t t0 = 0;         // offset in the .pvar image of the first staged byte (text_dense)
    uint64_t text_bytes = 0; // covering range (text_dense) or the kept rows packed back to back
    bool text_dense = false;
};

class WritePool;

struct Sink {
    int fd = -1;
    bool positional = false; // pwrite at absolute offsets
    uint64_t base = 0;
    uint8_t *mem = nullptr;
    bool mem_pinned = false;
    WritePool *pool = nullptr; // positional sinks only
    uint8_t *map = nullptr;    // regular-file sink mapped MAP_SHARED: map[0] is file offset `base`
    int dfd = -1;              // the same file opened O_DIRECT (PGB_ODIRECT=1): 4 KiB-aligned interiors bypass the page cache
    // ordered (non-positional) writes
    std::mutex mu;
    std::condition_variable cv;
    uint64_t next_seq = 0;
};

struct Job {
    pgb_file *f;
    const uint32_t *var_idx;
    uint64_t n_var;
    const uint32_t *sam_idx; // nullptr => all
    uint64_t K;
    const uint8_t *prefix_blob;
    const uint64_t *prefix_off; // blob mode: the caller's offsets; rows mode: cumulative prefix lengths (row_len + suffix)
    // rows mode: line i's prefix is text[row_off[i] .. + row_len[i]) followed by the suffix ("\tGT"), appended by K2
    const uint8_t *text = nullptr;
    const uint64_t *row_off = nullptr;
    const uint32_t *row_len = nullptr;
    uint32_t sfx = 0, sfx_len = 0;
    Sink *sink;
    int variant;
    bool trace = false;
    std::chrono::steady_clock::time_point t0;
    void lap(const char *what, size_t ci) const {
        if (trace)
            fprintf(stderr, "[pgb]   chunk %zu %-10s +%.3f ms\n", ci, what,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    std::atomic<int> status{PGB_OK};
    std::mutex err_mu;
    char err[512] = {0};
    void fail(int rc) {
        int expect = PGB_OK;
        if (status.compare_exchange_strong(expect, rc)) {
            std::lock_guard<std::mutex> g(err_mu);
            snprintf(err, sizeof err, "%s", pgb_last_error());
        }
    }
};

struct DeviceWork {
    DeviceCtx *ctx;
    std::vector<Chunk> chunks;
    uint64_t seq_base = 0;
    double device_ms = 0;
    uint64_t h2d = 0, d2h = 0, launches = 0;
};

int ensure_slot(Slot &s, uint64_t need_in, uint64_t need_out, uint64_t need_lines, bool need_h_out) {
    if (!s.stream) {
        CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&s.ev_k0));
        CU(cudaEventCreate(&s.ev_k1));
        CU(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    }
    if (need_in > s.cap_in) {
        if (s.h_in) cudaFreeHost(s.h_in);
        if (s.d_in) cudaFree(s.d_in);
        s.h_in = nullptr; s.d_in = nullptr; s.cap_in = 0;
        uint64_t cap = align_up(need_in + need_in / 8, 1 << 20);
        CU(cudaHostAlloc((void **)&s.h_in, cap, cudaHostAllocDefault));
        CU(cudaMalloc((void **)&s.d_in, cap + 64));
        s.cap_in = cap;
    }
    if (need_out > s.cap_out) {
        if (s.d_out) cudaFree(s.d_out);
        s.d_out = nullptr;
        if (s.h_out) cudaFreeHost(s.h_out);
        s.h_out = nullptr;
        s.cap_out = 0;
        uint64_t cap = align_up(need_out + need_out / 8, 1 << 20);
        CU(cudaMalloc((void **)&s.d_out, cap));
        s.cap_out = cap;
    }
    if (need_h_out && !s.h_out) CU(cudaHostAlloc((void **)&s.h_out, s.cap_out + 4096, cudaHostAllocDefault));
    if (need_lines > s.cap_lines) {
        if (s.d_meta) cudaFree(s.d_meta);
        if (s.d_scratch) cudaFree(s.d_scratch);
        s.d_meta = nullptr; s.d_scratch = nullptr; s.cap_lines = 0;
        uint64_t cap = need_lines + need_lines / 8 + 1024;
        CU(cudaMalloc((void **)&s.d_meta, (cap + 1) * sizeof(pgb_line_meta)));
        CU(cudaMalloc(&s.d_scratch, pgb_dev_index_scratch_bytes(cap)));
        s.cap_lines = cap;
    }
    return PGB_OK;
}

int read_fully(int fd, uint8_t *dst, uint64_t n, uint64_t off) {
    while (n) {
        ssize_t k = pread(fd, dst, n > (1u << 30) ? (1u << 30) : n, (off_t)off);
        if (k < 0) {
            if (errno == EINTR) continue;
            pgb_set_error("pread: %s", strerror(errno));
            return PGB_E_IO;
        }
        if (k == 0) { pgb_set_error("unexpected end of .pgen (read_exact, pfile.rs:170)"); return PGB_E_RANGE; }
        dst += k; off += (uint64_t)k; n -= (uint64_t)k;
    }
    return PGB_OK;
}

int write_fully(int fd, const uint8_t *src, uint64_t n, bool positional, uint64_t off) {
    while (n) {
        size_t want = n > (1u << 30) ? (1u << 30) : (size_t)n;
        ssize_t k = positional ? pwrite(fd, src, want, (off_t)off) : write(fd, src, want);
        if (k < 0) {
            if (errno == EINTR) continue;
            pgb_set_error("write: %s", strerror(errno));
            return PGB_E_IO;
        }
        src += k; off += (uint64_t)k; n -= (uint64_t)k;
    }
    return PGB_OK;
}

// Helper threads for positional sinks: a chunk that has landed in page-locked memory is cut
// into pieces that are pwrite()n concurrently (one writer thread per device cannot keep a
// tmpfs / page-cache sink busy: the copy into the page cache is CPU-bound).
class WritePool {
  public:
    explicit WritePool(int n_threads) {
        for (int i = 0; i < n_threads; i++) th_.emplace_back([this] { run(); });
    }
    ~WritePool() {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    // Writes [src, src+n) at file offset off (fd >= 0) or copies it to memory address off (fd < 0),
    // split into pieces; returns when all are done.
    int write(int fd, const uint8_t *src, uint64_t n, uint64_t off) {
        const uint64_t piece = std::max<uint64_t>(4ull << 20, align_up(n / (2 * (uint64_t)th_.size() + 1), 1 << 20));
        Batch b;
        for (uint64_t o = 0; o < n; o += piece) {
            Task t{fd, src + o, std::min(piece, n - o), off + o, &b};
            {
                std::lock_guard<std::mutex> g(mu_);
                b.pending++;
                q_.push_back(t);
            }
            cv_.notify_one();
        }
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return b.pending == 0; });
        if (b.rc) pgb_set_error("%s", b.err);
        return b.rc;
    }

  private:
    struct Batch {
        int pending = 0;
        int rc = PGB_OK;
        char err[160] = {0};
    };
    struct Task {
        int fd;
        const uint8_t *src;
        uint64_t n, off;
        Batch *batch;
    };
    void run() {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                t = q_.front();
                q_.pop_front();
            }
            pgb_clear_error();
            int rc = PGB_OK;
            if (t.fd < 0) memcpy((void *)(uintptr_t)t.off, t.src, t.n);
            else rc = write_fully(t.fd, t.src, t.n, true, t.off);
            {
                std::lock_guard<std::mutex> g(mu_);
                if (rc && !t.batch->rc) {
                    t.batch->rc = rc;
                    snprintf(t.batch->err, sizeof t.batch->err, "%s", pgb_last_error());
                }
                t.batch->pending--;
            }
            done_cv_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::deque<Task> q_;
    std::vector<std::thread> th_;
    bool stop_ = false;
};

// Stage layout inside h_in/d_in:  [records | pad16 | prefix bytes (or .pvar text) | pad16 | prefix_off u64[n+1] |
//                                   prefix_len u32[n] (rows mode, dense text) | var_row u32[n]]
struct Stage {
    uint64_t rec_bytes, pfx_pos, pfx_bytes, off_pos, len_pos, row_pos, total;
};

Stage stage_layout(const Job &j, const Chunk &c, bool records_inline) {
    Stage s;
    s.rec_bytes = records_inline ? c.in_bytes : 0;
    uint64_t n = c.b - c.a;
    s.pfx_pos = align_up(s.rec_bytes + 16, kAlign);
    s.pfx_bytes = j.text ? c.text_bytes : j.prefix_off[c.b] - j.prefix_off[c.a];
    s.off_pos = align_up(s.pfx_pos + s.pfx_bytes + 16, kAlign);
    s.len_pos = align_up(s.off_pos + (n + 1) * 8, kAlign);
    s.row_pos = align_up(s.len_pos + (j.text && c.text_dense ? n * 4 : 0), kAlign);
    // per-line record index input of K1: u32 row numbers (fixed-width) or u64 byte offsets (standard format)
    s.total = align_up(s.row_pos + (c.dense ? n * (j.f->standard ? 8 : 4) : 0), kAlign);
    return s;
}

// Writer side of one device: waits for a chunk's D2H, drains it to the sink, frees the slot.
struct Pending {
    int slot;
    size_t chunk;
    uint32_t shift; // the chunk sits at h_out + shift: file offset and memory address agree modulo 4 KiB (O_DIRECT)
};

void writer_loop(Job *job, DeviceWork *w, std::deque<Pending> *q, std::mutex *qmu, std::condition_variable *qcv,
                 bool *done) {
    DeviceCtx *c = w->ctx;
    cudaSetDevice(c->device);
    for (;;) {
        Pending p;
        {
            std::unique_lock<std::mutex> lk(*qmu);
            qcv->wait(lk, [&] { return !q->empty() || *done; });
            if (q->empty()) return;
            p = q->front();
            q->pop_front();
        }
        Slot &s = c->slot[p.slot];
        const Chunk &ch = w->chunks[p.chunk];
        cudaError_t e = cudaEventSynchronize(s.ev_done);
        if (e != cudaSuccess) {
            pgb_set_error("cudaEventSynchronize: %s", cudaGetErrorString(e));
            job->fail(PGB_E_CUDA);
        } else if (job->status.load() == PGB_OK) {
            job->lap("d2h done", p.chunk);
            float ms = 0;
            if (cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1) == cudaSuccess) w->device_ms += ms;
            Sink *sk = job->sink;
            int rc = PGB_OK;
            if (sk->mem) {
                if (!sk->mem_pinned) {
                    if (sk->pool) rc = sk->pool->write(-1, s.h_out, ch.out_bytes, (uint64_t)(uintptr_t)(sk->mem + ch.out_off));
                    else memcpy(sk->mem + ch.out_off, s.h_out, ch.out_bytes);
                }
            } else if (sk->map) {
                // page-cache pages are filled by parallel copies through the mapping: buffered pwrite()s
                // to one inode serialise on its lock (measured: 4-5 GB/s however many threads)
                if (sk->pool) rc = sk->pool->write(-1, s.h_out, ch.out_bytes, (uint64_t)(uintptr_t)(sk->map + ch.out_off));
                else memcpy(sk->map + ch.out_off, s.h_out, ch.out_bytes);
            } else if (sk->positional && sk->dfd >= 0) {
                // O_DIRECT output stage: the 4 KiB-aligned interior of the chunk goes from the page-locked slot straight
                // to the device queue (file offset, memory address and length all 4 KiB-aligned); the ragged ends
                // (shared 4 KiB blocks with the neighbouring chunks) go through the page cache
                const uint8_t *src = s.h_out + p.shift;
                const uint64_t fo = sk->base + ch.out_off, fe = fo + ch.out_bytes;
                const uint64_t a0 = align_up(fo, 4096), a1 = fe / 4096 * 4096;
                if (a0 < a1) {
                    rc = write_fully(sk->fd, src, a0 - fo, true, fo);
                    if (!rc) rc = sk->pool ? sk->pool->write(sk->dfd, src + (a0 - fo), a1 - a0, a0)
                                           : write_fully(sk->dfd, src + (a0 - fo), a1 - a0, true, a0);
                    if (!rc) rc = write_fully(sk->fd, src + (a1 - fo), fe - a1, true, a1);
                } else {
                    rc = write_fully(sk->fd, src, ch.out_bytes, true, fo);
                }
            } else if (sk->positional) {
                rc = sk->pool ? sk->pool->write(sk->fd, s.h_out, ch.out_bytes, sk->base + ch.out_off)
                              : write_fully(sk->fd, s.h_out, ch.out_bytes, true, sk->base + ch.out_off);
            } else {
                const uint64_t seq = w->seq_base + p.chunk;
                std::unique_lock<std::mutex> lk(sk->mu);
                sk->cv.wait(lk, [&] { return sk->next_seq == seq || job->status.load() != PGB_OK; });
                if (job->status.load() == PGB_OK) rc = write_fully(sk->fd, s.h_out, ch.out_bytes, false, 0);
                sk->next_seq = seq + 1;
                sk->cv.notify_all();
            }
            if (rc) job->fail(rc);
        }
        if (job->status.load() != PGB_OK && !job->sink->mem && !job->sink->positional) {
            // keep ordered writers from waiting forever on a failed predecessor
            std::lock_guard<std::mutex> lk(job->sink->mu);
            job->sink->next_seq = w->seq_base + p.chunk + 1;
            job->sink->cv.notify_all();
        }
        {
            std::lock_guard<std::mutex> lk(c->mu);
            s.busy = false;
        }
        c->cv.notify_all();
    }
}

int run_device_inner(Job *job, DeviceWork *w, std::deque<Pending> &q, std::mutex &qmu, std::condition_variable &qcv) {
    DeviceCtx *c = w->ctx;
    pgb_file *f = job->f;
    CU(cudaSetDevice(c->device));
    const uint32_t R = f->R;
    const uint64_t K = job->K;
    const bool gather = job->sam_idx != nullptr;
    const bool need_h_out = !(job->sink->mem && job->sink->mem_pinned);

    // ---- K0: keep-mask -> kept-sample index list (once per call and device) ----
    if (!c->ev_a) {
        CU(cudaEventCreate(&c->ev_a));
        CU(cudaEventCreate(&c->ev_b));
    }
    // make sure slot 0 has a stream
    int rc = ensure_slot(c->slot[0], 1, 1, 1, false);
    if (rc) return rc;
    if (gather) {
        const uint64_t n_mask = 4ull * R; // indices into padding bits are legal (pfile.rs:173 only bounds the byte)
        if (n_mask + 64 > c->cap_mask) {
            if (c->d_mask) cudaFree(c->d_mask);
            if (c->h_mask) cudaFreeHost(c->h_mask);
            if (c->d_kidx) cudaFree(c->d_kidx);
            c->d_mask = nullptr; c->h_mask = nullptr; c->d_kidx = nullptr; c->cap_mask = 0;
            CU(cudaMalloc((void **)&c->d_mask, n_mask + 64));
            CU(cudaHostAlloc((void **)&c->h_mask, n_mask + 64, cudaHostAllocDefault));
            CU(cudaMalloc((void **)&c->d_kidx, (n_mask + 64) * sizeof(uint32_t)));
            c->cap_mask = n_mask + 64;
        }
        if (!c->d_count) CU(cudaMalloc((void **)&c->d_count, sizeof(uint32_t)));
        memset(c->h_mask, 0, n_mask);
        for (uint64_t i = 0; i < K; i++) c->h_mask[job->sam_idx[i]] = 1;
        cudaStream_t st = c->slot[0].stream;
        CU(cudaMemcpyAsync(c->d_mask, c->h_mask, n_mask, cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(c->ev_a, st));
        rc = pgb_dev_compact_samples(c->d_mask, (uint32_t)n_mask, c->d_kidx, c->d_count, st);
        if (rc) return rc;
        CU(cudaEventRecord(c->ev_b, st));
        uint32_t cnt = 0;
        CU(cudaMemcpyAsync(&cnt, c->d_count, sizeof cnt, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, c->ev_a, c->ev_b));
        w->device_ms += ms;
        w->launches += 1;
        w->h2d += n_mask;
        if (cnt != K) { pgb_set_error("K0 count %u != %llu", cnt, (unsigned long long)K); return PGB_E_CUDA; }
    }

    for (size_t ci = 0; ci < w->chunks.size(); ci++) {
        if (job->status.load() != PGB_OK) break;
        const Chunk &ch = w->chunks[ci];
        const int si = (int)(ci % kSlots);
        Slot &s = c->slot[si];
        {
            std::unique_lock<std::mutex> lk(c->mu);
            c->cv.wait(lk, [&] { return !s.busy; });
            s.busy = true;
        }
        job->lap("slot free", ci);
        const bool rec_direct = f->image && f->image_pinned && ch.dense; // DMA records straight from the image
        const Stage lay = stage_layout(*job, ch, !rec_direct);
        const uint64_t n = ch.b - ch.a;
        const uint64_t rec_dev_bytes = ch.in_bytes;
        const uint64_t d_rec_region = rec_direct ? align_up(rec_dev_bytes + 16, kAlign) : 0;
        rc = ensure_slot(s, lay.total + d_rec_region, ch.out_bytes + 64, n, need_h_out);
        if (rc) return rc;

        // -- stage inputs in page-locked memory
        uint8_t *h = s.h_in;
        if (!rec_direct) {
            if (ch.dense) {
                const uint64_t off = ch.o0;
                if (f->image) memcpy(h, f->image + off, rec_dev_bytes);
                else if ((rc = read_fully(f->fd, h, rec_dev_bytes, off))) return rc;
            } else {
                for (uint64_t i = 0; i < n; i++) {
                    const uint64_t off = f->rec_off(job->var_idx ? job->var_idx[ch.a + i] : ch.a + i);
                    if (f->image) memcpy(h + i * R, f->image + off, R);
                    else if ((rc = read_fully(f->fd, h + i * R, R, off))) return rc;
                }
            }
            memset(h + lay.rec_bytes, 0, 16);
        }
        if (!job->text) {
            memcpy(h + lay.pfx_pos, job->prefix_blob + job->prefix_off[ch.a], lay.pfx_bytes);
            memcpy(h + lay.off_pos, job->prefix_off + ch.a, (n + 1) * 8);
        } else if (ch.text_dense) {
            // the covering range of the .pvar image as it is in the file; K1 gets (offset, length) per line
            memcpy(h + lay.pfx_pos, job->text + ch.t0, lay.pfx_bytes);
            uint64_t *po = (uint64_t *)(h + lay.off_pos);
            uint32_t *pl = (uint32_t *)(h + lay.len_pos);
            for (uint64_t i = 0; i < n; i++) {
                po[i] = job->row_off[ch.a + i] - ch.t0;
                pl[i] = job->row_len[ch.a + i];
            }
            po[n] = 0;
        } else {
            // sparse selection: the kept rows packed back to back (still without the suffix)
            uint64_t *po = (uint64_t *)(h + lay.off_pos);
            uint64_t at = 0;
            for (uint64_t i = 0; i < n; i++) {
                po[i] = at;
                memcpy(h + lay.pfx_pos + at, job->text + job->row_off[ch.a + i], job->row_len[ch.a + i]);
                at += job->row_len[ch.a + i];
            }
            po[n] = at;
        }
        memset(h + lay.pfx_pos + lay.pfx_bytes, 0, 16);
        if (ch.dense && f->standard) {
            uint64_t *ro = (uint64_t *)(h + lay.row_pos);
            for (uint64_t i = 0; i < n; i++) ro[i] = f->rec_off(job->var_idx ? job->var_idx[ch.a + i] : ch.a + i) - ch.o0;
        } else if (ch.dense) {
            uint32_t *vr = (uint32_t *)(h + lay.row_pos);
            const uint64_t v0 = (ch.o0 - 12) / (R ? R : 1);
            if (job->var_idx) for (uint64_t i = 0; i < n; i++) vr[i] = (uint32_t)(job->var_idx[ch.a + i] - v0);
            else for (uint64_t i = 0; i < n; i++) vr[i] = (uint32_t)(ch.a + i - v0);
        }

        job->lap("staged", ci);
        // -- H2D, K1, K2, D2H on the slot's stream
        cudaStream_t st = s.stream;
        uint8_t *d_stage = s.d_in + d_rec_region;
        const uint8_t *d_records;
        if (rec_direct) {
            CU(cudaMemcpyAsync(s.d_in, f->image + ch.o0, rec_dev_bytes, cudaMemcpyHostToDevice, st));
            CU(cudaMemsetAsync(s.d_in + rec_dev_bytes, 0, 16, st));
            d_records = s.d_in;
            w->h2d += rec_dev_bytes;
        } else {
            d_records = d_stage;
        }
        CU(cudaMemcpyAsync(d_stage, h, lay.total, cudaMemcpyHostToDevice, st));
        w->h2d += lay.total;
        CU(cudaEventRecord(s.ev_k0, st));
        {
            const bool rows_dense = job->text && ch.text_dense;
            const uint64_t *d_rec_off = ch.dense && f->standard ? (const uint64_t *)(d_stage + lay.row_pos) : nullptr;
            const uint32_t *d_var_row = ch.dense && !f->standard ? (const uint32_t *)(d_stage + lay.row_pos) : nullptr;
            rc = pgb_dev_index_lines_ex(d_var_row, d_rec_off, R, (const uint64_t *)(d_stage + lay.off_pos),
                                        rows_dense ? (const uint32_t *)(d_stage + lay.len_pos) : nullptr, job->sfx_len,
                                        job->text ? 0 : job->prefix_off[ch.a], n, (uint32_t)K, s.d_meta, s.d_scratch, st);
        }
        if (rc) return rc;
        rc = pgb_dev_format_lines_ex(d_records, R, s.d_meta, n, d_stage + lay.pfx_pos, job->sfx, job->sfx_len,
                                     gather ? c->d_kidx : nullptr, (uint32_t)K, ch.max_pfx, s.d_out, job->variant, st);
        if (rc) return rc;
        CU(cudaEventRecord(s.ev_k1, st));
        w->launches += 4;
        const uint32_t shift = job->sink->dfd >= 0 ? (uint32_t)((job->sink->base + ch.out_off) & 4095u) : 0u;
        uint8_t *dst = need_h_out ? s.h_out + shift : job->sink->mem + ch.out_off;
        CU(cudaMemcpyAsync(dst, s.d_out, ch.out_bytes, cudaMemcpyDeviceToHost, st));
        w->d2h += ch.out_bytes;
        CU(cudaEventRecord(s.ev_done, st));
        job->lap("enqueued", ci);
        {
            std::lock_guard<std::mutex> lk(qmu);
            q.push_back(Pending{si, ci, shift});
        }
        qcv.notify_one();
    }
    return PGB_OK;
}

void run_device(Job *job, DeviceWork *w) {
    std::deque<Pending> q;
    std::mutex qmu;
    std::condition_variable qcv;
    bool done = false;
    pgb_clear_error();
    std::thread writer(writer_loop, job, w, &q, &qmu, &qcv, &done);
    int rc = run_device_inner(job, w, q, qmu, qcv);
    if (rc) job->fail(rc);
    {
        std::lock_guard<std::mutex> lk(qmu);
        done = true;
    }
    qcv.notify_all();
    writer.join();
    // a failed producer may leave ordered writers of later devices waiting
    if (job->status.load() != PGB_OK && !job->sink->mem && !job->sink->positional) {
        std::lock_guard<std::mutex> lk(job->sink->mu);
        job->sink->next_seq = UINT64_MAX;
        job->sink->cv.notify_all();
    }
    if (cudaSetDevice(w->ctx->device) == cudaSuccess) {
        for (auto &s : w->ctx->slot)
            if (s.stream) cudaStreamSynchronize(s.stream);
    }
    // The context outlives this export (process-wide cache): a producer that failed between claiming
    // a slot and queueing it must not leave the slot marked busy for the next export.
    {
        std::lock_guard<std::mutex> lk(w->ctx->mu);
        for (auto &s : w->ctx->slot) s.busy = false;
    }
    if (job->status.load() != PGB_OK) cudaGetLastError(); // clear a sticky non-fatal error state
}

// Where the line prefixes (pfile.rs:157-161) come from: a blob of finished prefixes with n_var + 1 offsets, or the
// raw .pvar image with one (offset, length) per kept row — then the "\tGT" that ends every prefix is appended
// on the device and no per-row host pass builds a blob.
struct PrefixSrc {
    const uint8_t *blob = nullptr;
    const uint64_t *off = nullptr;
    const uint8_t *text = nullptr;
    uint64_t text_bytes = 0;
    const uint64_t *row_off = nullptr;
    const uint32_t *row_len = nullptr;
};

int export_impl(pgb_file *f, const uint32_t *var_idx, uint64_t n_var, const uint32_t *sam_idx, uint64_t n_sam,
                const PrefixSrc &pfx, Sink *sink, uint64_t out_cap, uint64_t *out_len, const int *device_ids, int n_devices,
                pgb_stats *stats) {
    const uint8_t *prefix_blob = pfx.blob;
    const uint64_t *prefix_off = pfx.off;
    const bool rows_mode = pfx.text != nullptr || pfx.row_off != nullptr;
    constexpr uint32_t kSfx = 0x00544709u, kSfxLen = 3; // "\tGT"
    std::vector<uint64_t> cum; // rows mode: cumulative prefix lengths stand in for prefix_off
    if (rows_mode) {
        if (n_var && (!pfx.row_off || !pfx.row_len || !pfx.text)) { pgb_set_error("row_off / row_len / pvar_text is NULL"); return PGB_E_ARG; }
        cum.resize(n_var + 1);
        uint64_t at = 0;
        for (uint64_t i = 0; i < n_var; i++) {
            if (pfx.row_off[i] > pfx.text_bytes || pfx.row_len[i] > pfx.text_bytes - pfx.row_off[i] || pfx.row_len[i] > 0x7ffffff0u) {
                pgb_set_error("row %llu lies outside the .pvar image", (unsigned long long)i);
                return PGB_E_RANGE;
            }
            cum[i] = at;
            at += (uint64_t)pfx.row_len[i] + kSfxLen;
        }
        cum[n_var] = at;
        prefix_off = cum.data();
    }
    const auto t_begin = std::chrono::steady_clock::now();
    const bool trace = env_u64("PGB_TRACE", 0) != 0;
    auto lap = [&](const char *what) {
        if (trace)
            fprintf(stderr, "[pgb] %-12s +%.3f ms\n", what,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    pgb_clear_error();
    if (stats) memset(stats, 0, sizeof *stats);
    if (!f) return PGB_E_ARG;
    if (n_var && !prefix_off) { pgb_set_error("prefix_off is NULL"); return PGB_E_ARG; }
    const uint32_t R = f->R;
    const uint64_t K = sam_idx ? n_sam : f->N;

    // ---- validation (the reference panics lazily at pfile.rs:170/173; we fail up front) ----
    if (sam_idx) {
        for (uint64_t i = 0; i < n_sam; i++) {
            if ((uint64_t)sam_idx[i] / 4 >= R) { pgb_set_error("sample index %u outside the record", sam_idx[i]); return PGB_E_RANGE; }
            if (i && sam_idx[i] <= sam_idx[i - 1]) { pgb_set_error("sam_idx must be strictly ascending"); return PGB_E_ARG; }
        }
    }
    if (K > 0xffffffffull / 4) return PGB_E_ARG;
    // A strictly ascending list of N indices ending at N-1 is the identity (what
    // filter_metadata returns for a None query, pfile.rs:321): take the keep-all path.
    if (sam_idx && n_sam == f->N && n_sam > 0 && sam_idx[n_sam - 1] == f->N - 1) sam_idx = nullptr;
    if (!var_idx && n_var > f->M) { pgb_set_error("n_var %llu > variants in file %u", (unsigned long long)n_var, f->M); return PGB_E_RANGE; }
    uint64_t total_pfx = 0;
    // the planner below places chunk boundaries by binary search when rows (and, in rows mode, row offsets) ascend —
    // what filter_metadata produces (pfile.rs:319,331); anything else gets the linear scan
    bool ascending = true;
    uint32_t maxp_all = 0;
    for (uint64_t i = 0; i < n_var; i++) {
        const uint64_t v = var_idx ? var_idx[i] : i;
        if (i && var_idx && var_idx[i] <= var_idx[i - 1]) ascending = false;
        if (i && rows_mode && pfx.row_off[i] < pfx.row_off[i - 1] + pfx.row_len[i - 1]) ascending = false;
        if (v >= f->M && f->standard) {
            pgb_set_error("variant row %llu outside the %u variants of the file", (unsigned long long)v, f->M);
            return PGB_E_RANGE;
        }
        if (f->standard && (f->type10[v] != 0 || f->len10[v] != R)) {
            pgb_set_error("variant row %llu is stored as record type %u, length %u: only plain 2-bit hardcall records "
                          "(type 0, %u bytes) can be decoded", (unsigned long long)v, f->type10[v], f->len10[v], R);
            return PGB_E_MODE;
        }
        if (f->rec_off(v) + R > f->bytes) {
            pgb_set_error("variant row %llu lies outside the .pgen (read_exact, pfile.rs:170)", (unsigned long long)v);
            return PGB_E_RANGE;
        }
        if (prefix_off[i + 1] < prefix_off[i] || prefix_off[i + 1] - prefix_off[i] > 0x7fffffffull) {
            pgb_set_error("prefix_off must be ascending");
            return PGB_E_ARG;
        }
        maxp_all = std::max(maxp_all, (uint32_t)(prefix_off[i + 1] - prefix_off[i]));
    }
    if (n_var) total_pfx = prefix_off[n_var] - prefix_off[0];
    if (total_pfx && !prefix_blob && !rows_mode) return PGB_E_ARG;
    const uint64_t fixed = 4ull * K + 1ull;
    const uint64_t total = total_pfx + n_var * fixed;
    if (out_len) *out_len = total;
    if (sink->mem && total > out_cap) return PGB_E_SPACE;
    if (stats) {
        stats->n_lines = n_var;
        stats->n_kept_samples = K;
        stats->genotypes = n_var * K;
        stats->bytes_out = total;
    }
    if (n_var == 0) return PGB_OK;

    lap("validated");
    // ---- devices ----
    int n_avail = 0;
    if (cudaGetDeviceCount(&n_avail) != cudaSuccess || n_avail <= 0) {
        cudaGetLastError();
        pgb_set_error("cudaGetDeviceCount found no device");
        return PGB_E_NO_DEVICE;
    }
    std::vector<int> devs;
    if (!device_ids || n_devices <= 0) devs.push_back(0);
    else devs.assign(device_ids, device_ids + n_devices);
    for (int d : devs)
        if (d < 0 || d >= n_avail) { pgb_set_error("device %d not present (%d visible)", d, n_avail); return PGB_E_NO_DEVICE; }
    const int G = (int)devs.size();

    for (int g = 0; g < G; g++)
        for (int h = 0; h < g; h++)
            if (devs[h] == devs[g]) { pgb_set_error("duplicate device id"); return PGB_E_ARG; }
    std::lock_guard<std::mutex> file_lock(f->mu);
    std::vector<DeviceWork> work(G);
    CtxLease lease;
    acquire_ctxs(devs, &lease.ctxs);
    for (int g = 0; g < G; g++) work[g].ctx = lease.ctxs[g];

    lap("ctx leased");
    // ---- plan: contiguous line ranges per device balanced by output bytes, then chunks ----
    // Output bytes per chunk: 128 MiB for big exports (64..1024 MiB measured within 3 % of each other),
    // smaller for small ones so that at least ~16 chunks keep H2D, kernels, D2H and the sink overlapped
    uint64_t chunk_out = env_u64("PGB_CHUNK_MB", 128) << 20;
    if (!getenv("PGB_CHUNK_MB")) chunk_out = std::min<uint64_t>(128ull << 20, std::max<uint64_t>(8ull << 20, total / 16));
    const uint64_t chunk_in = env_u64("PGB_CHUNK_IN_MB", 128) << 20;
    auto out_before = [&](uint64_t i) { return (prefix_off[i] - prefix_off[0]) + i * fixed; };
    uint64_t seq = 0;
    uint64_t line = 0;
    std::vector<uint64_t> shard_begin(G + 1), shard_byte(G + 1);
    pgb_shard_plan(n_var, K, prefix_off, G, shard_begin.data(), shard_byte.data());
    for (int g = 0; g < G; g++) {
        const uint64_t end = shard_begin[g + 1];
        work[g].seq_base = seq;
        while (line < end) {
            Chunk c;
            c.a = line;
            uint64_t b = line + 1;
            uint64_t omin = f->rec_off(var_idx ? var_idx[line] : line), omax = omin;
            uint32_t maxp = (uint32_t)(prefix_off[line + 1] - prefix_off[line]);
            uint64_t tmin = rows_mode ? pfx.row_off[line] : 0, tmax = rows_mode ? pfx.row_off[line] + pfx.row_len[line] : 0;
            if (ascending && !env_u64("PGB_PLAN_LINEAR", 0)) {
                // every limit is monotonic in the chunk end: largest b with [line, b) inside all of them
                auto fits = [&](uint64_t e) { // e > line
                    if (out_before(e) - out_before(line) > chunk_out) return false;
                    if ((e - line) * (uint64_t)R > chunk_in) return false;
                    const uint64_t o = f->rec_off(var_idx ? var_idx[e - 1] : e - 1);
                    return f->standard || (o - omin) / (R ? R : 1) < 0xffffffffull;
                };
                uint64_t lo = line + 1, hi = end; // fits(lo) by definition (a chunk holds at least one line)
                while (lo < hi) {
                    const uint64_t mid = lo + (hi - lo + 1) / 2;
                    if (fits(mid)) lo = mid; else hi = mid - 1;
                }
                b = lo;
                omax = f->rec_off(var_idx ? var_idx[b - 1] : b - 1);
                maxp = maxp_all; // the export's largest prefix: only sizes tiles / shared memory
                if (rows_mode) tmax = std::max(tmax, pfx.row_off[b - 1] + pfx.row_len[b - 1]);
            } else
            while (b < end) {
                const uint64_t ob = out_before(b + 1) - out_before(line);
                if (ob > chunk_out) break;
                const uint64_t o = f->rec_off(var_idx ? var_idx[b] : b);
                const uint64_t nmin = std::min(omin, o), nmax = std::max(omax, o);
                const uint64_t compact = (b + 1 - line) * (uint64_t)R; // bytes of the kept rows alone
                if (compact > chunk_in) break;
                if (!f->standard && (nmax - nmin) / (R ? R : 1) >= 0xffffffffull) break;
                omin = nmin; omax = nmax;
                maxp = std::max(maxp, (uint32_t)(prefix_off[b + 1] - prefix_off[b]));
                if (rows_mode) {
                    tmin = std::min(tmin, pfx.row_off[b]);
                    tmax = std::max(tmax, pfx.row_off[b] + pfx.row_len[b]);
                }
                b++;
            }
            c.b = b;
            c.o0 = omin;
            const uint64_t cover_bytes = omax - omin + R, kept_bytes = (b - line) * (uint64_t)R;
            // density >= 25 %: move the whole covering byte range with one DMA (at most 4 x chunk_in bytes,
            // free on the otherwise idle H2D direction) instead of gathering the kept rows on the CPU
            c.dense = cover_bytes <= 4 * kept_bytes;
            c.in_bytes = c.dense ? cover_bytes : kept_bytes;
            c.out_off = out_before(line);
            c.out_bytes = out_before(b) - c.out_off;
            c.max_pfx = maxp;
            if (rows_mode) {
                // rows of a dense selection: DMA the covering range of the .pvar image as it is (the rows between
                // kept rows and the line terminators travel along); sparse: the kept rows packed on the host
                const uint64_t kept_text = (prefix_off[b] - prefix_off[line]) - (b - line) * kSfxLen;
                c.text_dense = tmax - tmin <= 4 * kept_text + 4096;
                c.t0 = tmin;
                c.text_bytes = c.text_dense ? tmax - tmin : kept_text;
            }
            work[g].chunks.push_back(c);
            seq++;
            line = b;
        }
    }

    lap("planned");
    Job job;
    job.f = f; job.var_idx = var_idx; job.n_var = n_var; job.sam_idx = sam_idx; job.K = K;
    job.prefix_blob = prefix_blob; job.prefix_off = prefix_off; job.sink = sink;
    if (rows_mode) {
        job.text = pfx.text; job.row_off = pfx.row_off; job.row_len = pfx.row_len;
        job.sfx = kSfx; job.sfx_len = kSfxLen;
    }
    job.variant = (int)env_u64("PGB_K2_VARIANT", 0);
    job.trace = trace;
    job.t0 = t_begin;

    std::unique_ptr<WritePool> pool;
    struct Unmap {
        void *p = nullptr;
        size_t n = 0;
        ~Unmap() { if (p) munmap(p, n); }
    } unmap;
    if ((sink->fd >= 0 && sink->positional) || (sink->mem && !sink->mem_pinned && total > (64u << 20))) {
        if (sink->fd >= 0) {
            // regular file: size it once so that concurrent writers do not serialise on extending it
            struct stat st;
            const bool regular = fstat(sink->fd, &st) == 0 && S_ISREG(st.st_mode);
            if (regular && (uint64_t)st.st_size < sink->base + total) (void)!ftruncate(sink->fd, (off_t)(sink->base + total));
            // Measured (profiles/README.md): on tmpfs parallel copies through a mapping reach 7 GB/s vs 4 GB/s
            // for pwrite(); on ext4 the mapping's write faults make it slower (2 vs 4.7 GB/s) -> tmpfs only.
            struct statfs sfs;
            const bool tmpfs = fstatfs(sink->fd, &sfs) == 0 && (unsigned long)sfs.f_type == 0x01021994ul;
            // O_DIRECT output stage (opt-in, PGB_ODIRECT=1): measured on this pool's boxes (virtio disk, ext4) it is
            // slower than page-cache writes that never reach the disk inside the call (3.9 vs 5.5 GB/s, profiles/README.md);
            // on a box whose page cache cannot hold the VCF it is the path that keeps the export off the cache
            if (regular && !tmpfs && total >= (8u << 20) && env_u64("PGB_ODIRECT", 0)) {
                char link[64];
                snprintf(link, sizeof link, "/proc/self/fd/%d", sink->fd);
                sink->dfd = open(link, O_WRONLY | O_DIRECT);
            }
            if (sink->dfd < 0 && regular && total >= (64u << 20) && env_u64("PGB_MMAP_SINK", tmpfs ? 1 : 0)) {
                // map the body region; the caller's descriptor is usually write-only (File::create), which
                // mmap(PROT_WRITE, MAP_SHARED) rejects, so reopen the same file read-write through /proc
                char link[64];
                snprintf(link, sizeof link, "/proc/self/fd/%d", sink->fd);
                const int rw = open(link, O_RDWR);
                if (rw >= 0) {
                    const uint64_t page = (uint64_t)sysconf(_SC_PAGESIZE);
                    const uint64_t lo = sink->base / page * page;
                    const size_t len = (size_t)(sink->base + total - lo);
                    void *m = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, rw, (off_t)lo);
                    close(rw);
                    if (m != MAP_FAILED) {
                        unmap.p = m;
                        unmap.n = len;
                        sink->map = (uint8_t *)m + (sink->base - lo);
                    }
                }
            }
        }
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const int n_writers = (int)env_u64("PGB_WRITERS", std::min<uint64_t>(8, std::max(1u, hw / 2)));
        if (n_writers > 1) {
            pool.reset(new WritePool(n_writers));
            sink->pool = pool.get();
        }
    }
    if (G == 1) run_device(&job, &work[0]);
    else {
        std::vector<std::thread> th;
        for (int g = 0; g < G; g++) th.emplace_back(run_device, &job, &work[g]);
        for (auto &t : th) t.join();
    }
    lap("devices done");
    sink->pool = nullptr; // all three die with this call
    sink->map = nullptr;
    if (sink->dfd >= 0) {
        close(sink->dfd);
        sink->dfd = -1;
    }
    int rc = job.status.load();
    if (rc != PGB_OK) {
        pgb_set_error("%s", job.err);
        return rc;
    }
    if (sink->fd >= 0 && sink->positional) {
        if (lseek(sink->fd, (off_t)(sink->base + total), SEEK_SET) < 0) { pgb_set_error("lseek: %s", strerror(errno)); return PGB_E_IO; }
    }
    if (stats) {
        for (auto &w : work) {
            stats->device_ms = std::max(stats->device_ms, w.device_ms);
            stats->bytes_h2d += w.h2d;
            stats->bytes_d2h += w.d2h;
            stats->kernel_launches += w.launches;
            stats->n_chunks += (int32_t)w.chunks.size();
        }
        stats->n_devices = G;
        stats->e2e_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    }
    return PGB_OK;
}

bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

} // namespace

extern "C" int pgb_shard_plan(uint64_t n_var, uint64_t n_kept_samples, const uint64_t *prefix_off, int n_shards,
                              uint64_t *line_begin, uint64_t *byte_begin) {
    if (n_shards <= 0 || !line_begin || (n_var && !prefix_off)) return PGB_E_ARG;
    const uint64_t fixed = 4ull * n_kept_samples + 1ull;
    auto out_before = [&](uint64_t i) { return i ? (prefix_off[i] - prefix_off[0]) + i * fixed : 0; };
    const uint64_t total = out_before(n_var);
    uint64_t line = 0;
    line_begin[0] = 0;
    if (byte_begin) byte_begin[0] = 0;
    for (int g = 0; g < n_shards; g++) {
        uint64_t end;
        if (g == n_shards - 1) end = n_var;
        else {
            // first line boundary at or past g+1 equal shares of the output bytes
            const uint64_t target = total / (uint64_t)n_shards * (uint64_t)(g + 1);
            uint64_t lo = line, hi = n_var;
            while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (out_before(mid) < target) lo = mid + 1; else hi = mid; }
            end = lo;
        }
        line_begin[g + 1] = end;
        if (byte_begin) byte_begin[g + 1] = out_before(end);
        line = end;
    }
    return PGB_OK;
}

extern "C" int pgb_open(const char *pgen_path, pgb_file **out) {
    pgb_clear_error();
    if (!pgen_path || !out) return PGB_E_ARG;
    *out = nullptr;
    int fd = open(pgen_path, O_RDONLY);
    if (fd < 0) { pgb_set_error("open %s: %s", pgen_path, strerror(errno)); return PGB_E_IO; }
    uint8_t h[12];
    size_t got = 0;
    while (got < 12) {
        ssize_t k = pread(fd, h + got, 12 - got, (off_t)got);
        if (k < 0 && errno == EINTR) continue;
        if (k <= 0) break;
        got += (size_t)k;
    }
    pgb_file *f = new pgb_file();
    int rc = parse_header(h, got, f);
    if (rc) { close(fd); delete f; return rc; }
    struct stat st;
    if (fstat(fd, &st) != 0) { pgb_set_error("fstat: %s", strerror(errno)); close(fd); delete f; return PGB_E_IO; }
    f->fd = fd;
    f->bytes = (uint64_t)st.st_size;
    *out = f;
    return PGB_OK;
}

extern "C" int pgb_open_standard(const char *pgen_path, pgb_file **out) {
    pgb_clear_error();
    if (!pgen_path || !out) return PGB_E_ARG;
    *out = nullptr;
    pgb_pgen10_info info;
    int rc = pgb_pgen10_index(pgen_path, &info, nullptr, nullptr, nullptr);
    if (rc) return rc;
    pgb_file *f = new pgb_file();
    f->standard = true;
    f->M = info.n_variants;
    f->N = info.n_samples;
    f->R = pgb_record_bytes(f->N);
    f->off10.resize((size_t)f->M + 1);
    f->type10.resize(std::max<size_t>(f->M, 1));
    f->len10.resize(std::max<size_t>(f->M, 1));
    rc = pgb_pgen10_index(pgen_path, &info, f->off10.data(), f->type10.data(), f->len10.data());
    if (rc) { delete f; return rc; }
    const int fd = open(pgen_path, O_RDONLY);
    if (fd < 0) { pgb_set_error("open %s: %s", pgen_path, strerror(errno)); delete f; return PGB_E_IO; }
    struct stat st;
    if (fstat(fd, &st) != 0) { pgb_set_error("fstat: %s", strerror(errno)); close(fd); delete f; return PGB_E_IO; }
    f->fd = fd;
    f->bytes = (uint64_t)st.st_size;
    *out = f;
    return PGB_OK;
}

extern "C" int pgb_open_mem(const void *image, uint64_t image_bytes, pgb_file **out) {
    pgb_clear_error();
    if (!image || !out) return PGB_E_ARG;
    *out = nullptr;
    pgb_file *f = new pgb_file();
    int rc = parse_header((const uint8_t *)image, (size_t)std::min<uint64_t>(image_bytes, 12), f);
    if (rc) { delete f; return rc; }
    f->image = (const uint8_t *)image;
    f->bytes = image_bytes;
    f->image_pinned = is_pinned(image);
    *out = f;
    return PGB_OK;
}

extern "C" void pgb_dims(const pgb_file *f, uint32_t *n_variants, uint32_t *n_samples, uint32_t *record_bytes) {
    if (!f) return;
    if (n_variants) *n_variants = f->M;
    if (n_samples) *n_samples = f->N;
    if (record_bytes) *record_bytes = f->R;
}

extern "C" void pgb_close(pgb_file *f) {
    if (!f) return;
    if (f->fd >= 0) close(f->fd);
    delete f;
}

extern "C" void pgb_release_buffers(void) {
    std::vector<DeviceCtx *> idle;
    {
        std::lock_guard<std::mutex> lk(g_registry.mu);
        for (size_t i = 0; i < g_registry.all.size();) {
            if (!g_registry.in_use[i]) {
                idle.push_back(g_registry.all[i]);
                g_registry.all.erase(g_registry.all.begin() + (long)i);
                g_registry.in_use.erase(g_registry.in_use.begin() + (long)i);
            } else i++;
        }
    }
    for (DeviceCtx *c : idle) free_ctx(c);
}

extern "C" int pgb_export_gt_vcf(pgb_file *f, const uint32_t *var_idx, uint64_t n_var, const uint32_t *sam_idx,
                                 uint64_t n_sam, const uint8_t *prefix_blob, const uint64_t *prefix_off, int out_fd,
                                 const int *device_ids, int n_devices, pgb_stats *stats) {
    if (out_fd < 0) return PGB_E_ARG;
    Sink sink;
    sink.fd = out_fd;
    off_t cur = lseek(out_fd, 0, SEEK_CUR);
    int fl = fcntl(out_fd, F_GETFL);
    sink.positional = cur >= 0 && fl >= 0 && !(fl & O_APPEND);
    sink.base = cur >= 0 ? (uint64_t)cur : 0;
    PrefixSrc pfx;
    pfx.blob = prefix_blob;
    pfx.off = prefix_off;
    return export_impl(f, var_idx, n_var, sam_idx, n_sam, pfx, &sink, 0, nullptr, device_ids, n_devices, stats);
}

static void fd_sink(Sink *sink, int out_fd) {
    sink->fd = out_fd;
    off_t cur = lseek(out_fd, 0, SEEK_CUR);
    int fl = fcntl(out_fd, F_GETFL);
    sink->positional = cur >= 0 && fl >= 0 && !(fl & O_APPEND);
    sink->base = cur >= 0 ? (uint64_t)cur : 0;
}

extern "C" int pgb_export_gt_vcf_rows(pgb_file *f, const uint32_t *var_idx, uint64_t n_var, const uint32_t *sam_idx,
                                      uint64_t n_sam, const uint8_t *pvar_text, uint64_t pvar_bytes, const uint64_t *row_off,
                                      const uint32_t *row_len, int out_fd, const int *device_ids, int n_devices,
                                      pgb_stats *stats) {
    if (out_fd < 0) return PGB_E_ARG;
    Sink sink;
    fd_sink(&sink, out_fd);
    static const uint8_t empty = 0;
    PrefixSrc pfx;
    pfx.text = pvar_text ? pvar_text : &empty;
    pfx.text_bytes = pvar_text ? pvar_bytes : 0;
    pfx.row_off = row_off;
    pfx.row_len = row_len;
    return export_impl(f, var_idx, n_var, sam_idx, n_sam, pfx, &sink, 0, nullptr, device_ids, n_devices, stats);
}

extern "C" int pgb_export_gt_vcf_rows_mem(pgb_file *f, const uint32_t *var_idx, uint64_t n_var, const uint32_t *sam_idx,
                                          uint64_t n_sam, const uint8_t *pvar_text, uint64_t pvar_bytes,
                                          const uint64_t *row_off, const uint32_t *row_len, uint8_t *out_buf, uint64_t out_cap,
                                          uint64_t *out_len, const int *device_ids, int n_devices, pgb_stats *stats) {
    if (!out_buf && out_cap) return PGB_E_ARG;
    Sink sink;
    static uint8_t dummy;
    static const uint8_t empty = 0;
    sink.mem = out_buf ? out_buf : &dummy;
    sink.mem_pinned = out_buf && is_pinned(out_buf);
    PrefixSrc pfx;
    pfx.text = pvar_text ? pvar_text : &empty;
    pfx.text_bytes = pvar_text ? pvar_bytes : 0;
    pfx.row_off = row_off;
    pfx.row_len = row_len;
    return export_impl(f, var_idx, n_var, sam_idx, n_sam, pfx, &sink, out_cap, out_len, device_ids, n_devices, stats);
}

extern "C" int pgb_export_gt_vcf_mem(pgb_file *f, const uint32_t *var_idx, uint64_t n_var, const uint32_t *sam_idx,
                                     uint64_t n_sam, const uint8_t *prefix_blob, const uint64_t *prefix_off,
                                     uint8_t *out_buf, uint64_t out_cap, uint64_t *out_len, const int *device_ids,
                                     int n_devices, pgb_stats *stats) {
    if (!out_buf && out_cap) return PGB_E_ARG;
    Sink sink;
    static uint8_t dummy;
    sink.mem = out_buf ? out_buf : &dummy;
    sink.mem_pinned = out_buf && is_pinned(out_buf);
    PrefixSrc pfx;
    pfx.blob = prefix_blob;
    pfx.off = prefix_off;
    return export_impl(f, var_idx, n_var, sam_idx, n_sam, pfx, &sink, out_cap, out_len, device_ids, n_devices, stats);
}
