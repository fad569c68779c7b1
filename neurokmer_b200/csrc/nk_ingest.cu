// nk_ingest.cu — getting bytes that sit in PAGEABLE host memory (a plain malloc'ed batch, the page cache
// behind a file) onto the device at link speed, and parsing FASTA / FASTQ there.
//
// Replaces the reference's producer thread + unbounded channels (src/spiking_hash.rs:285-303, 405-422) and
// needletail's host parser (src/utils.rs:9-24).  A single cudaMemcpy from pageable memory runs at ~11 GB/s
// (the driver stages it on one thread); here a pool of host threads copies 2 MiB pieces into its own pinned
// slots (memcpy, or pread straight from the file) and each thread issues the H2D copy of its piece on its own
// stream — page-cache reads, the staging copies and PCIe overlap.  The pool's threads are bound to the CPUs
// that are local to the GPU (sysfs local_cpulist) and allocate their pinned slots themselves, so the staging
// memory sits on the GPU's NUMA node.
#include <fcntl.h>
#include <sched.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>

#include "nk_host.h"
#include "nk_internal.h"

using namespace nkd;

namespace nkd {

namespace {

size_t piece_bytes() {
    static const size_t v = [] {
        size_t kb = 2048;
        if (const char* e = getenv("NK_STAGE_PIECE_KB")) { const long t = atol(e); if (t >= 64 && t <= 65536) kb = (size_t)t; }
        return kb << 10;
    }();
    return v;
}

// CPUs local to the device's PCIe root ("0-15,32-47" style list from sysfs); empty if unknown
std::vector<int> local_cpus(int device) {
    std::vector<int> cpus;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return cpus; }
    for (char* p = bus; *p; ++p) *p = (char)tolower(*p);
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return cpus;
    char line[4096];
    if (fgets(line, sizeof line, f)) {
        for (char* p = line; *p && *p != '\n';) {
            char* e = nullptr;
            const long a = strtol(p, &e, 10);
            if (e == p) break;
            long b = a;
            p = e;
            if (*p == '-') { b = strtol(p + 1, &e, 10); p = e; }
            for (long c = a; c <= b && cpus.size() < 4096; ++c) cpus.push_back((int)c);
            if (*p == ',') ++p;
        }
    }
    fclose(f);
    return cpus;
}

}  // namespace

class StagePool {
public:
    StagePool(int device, unsigned nthreads) : device_(device) {
        const std::vector<int> cpus = getenv("NK_STAGE_NO_AFFINITY") ? std::vector<int>() : local_cpus(device);
        workers_.resize(nthreads);
        for (unsigned t = 0; t < nthreads; ++t) {
            workers_[t].th = std::thread([this, t, cpus] { run(t, cpus); });
        }
        // wait until every worker has its stream and slots (or failed)
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return ready_ == workers_.size(); });
    }
    ~StagePool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            quit_ = true;
            ++gen_;
            gen_a_.store(gen_, std::memory_order_release);
        }
        cv_job_.notify_all();
        for (auto& w : workers_) w.th.join();
    }
    bool ok() const { return !init_failed_; }
    unsigned threads() const { return (unsigned)workers_.size(); }
    static constexpr unsigned kCopyWorkers = 8;

    // dst_other != null: a PACK job — the workers turn their pieces of ASCII bases into 2-bit code words + `other` bits
    // (the host packer's SIMD bodies) in their pinned slots and copy those: 3/8 of the bytes on PCIe, and 3/8 of the
    // bytes written to host memory instead of a second copy of all of them
    int copy(const uint8_t* src, int fd, uint64_t off, uint64_t n, unsigned char* dst, cudaEvent_t after, cudaStream_t then,
             unsigned char* dst_other = nullptr) {
        if (n == 0) return NK_OK;
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = Job{src, fd, off, n, dst, after, dst_other,
                       dst_other ? (unsigned)workers_.size() : std::min<unsigned>((unsigned)workers_.size(), kCopyWorkers)};
            next_.store(0);
            npieces_ = (n + piece_bytes() - 1) / piece_bytes();
            finished_ = 0;
            finished_a_.store(0, std::memory_order_relaxed);
            failed_ = false;
            ++gen_;
            gen_a_.store(gen_, std::memory_order_release);
        }
        cv_job_.notify_all();
        // the workers finish within a millisecond or two of each other: spin for them (a sleeping thread of a VM costs
        // ~0.1-0.4 ms to wake, which a job of four 32 MiB chunks pays four times on each side), then sleep if they don't
        spin_until([&] { return finished_a_.load(std::memory_order_acquire) == workers_.size(); }, 20000);
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_done_.wait(lk, [&] { return finished_ == workers_.size(); });
        }
        if (failed_) return fail(NK_ERR_IO, "staging copy failed: %s", err_.c_str());
        for (auto& w : workers_) {
            if (!w.used) continue;
            cudaError_t e = cudaStreamWaitEvent(then, w.done, 0);
            if (e != cudaSuccess) return fail(NK_ERR_CUDA, "cudaStreamWaitEvent: %s", cudaGetErrorString(e));
        }
        return NK_OK;
    }

private:
    struct Job {
        const uint8_t* src; int fd; uint64_t off, n; unsigned char* dst; cudaEvent_t after; unsigned char* dst_other;
        unsigned active;   // workers that take pieces of this job (copies stop scaling at eight, packing does not)
    };
    struct Worker {
        std::thread th;
        cudaStream_t stream = nullptr;
        uint8_t* slot[2] = {nullptr, nullptr};
        cudaEvent_t ev[2] = {nullptr, nullptr};
        bool inflight[2] = {false, false};
        cudaEvent_t done = nullptr;
        bool used = false;
    };

    void run(unsigned t, const std::vector<int>& cpus) {
        Worker& w = workers_[t];
        if (!cpus.empty()) {
            cpu_set_t set;
            CPU_ZERO(&set);
            for (int c : cpus) if (c < CPU_SETSIZE) CPU_SET(c, &set);
            sched_setaffinity(0, sizeof set, &set);  // best effort
        }
        bool good = cudaSetDevice(device_) == cudaSuccess &&
                    cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking) == cudaSuccess &&
                    cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < 2 && good; ++i) {
            good = cudaMallocHost((void**)&w.slot[i], piece_bytes()) == cudaSuccess &&
                   cudaEventCreateWithFlags(&w.ev[i], cudaEventDisableTiming) == cudaSuccess;
            if (good) memset(w.slot[i], 0, piece_bytes());  // first touch on this (GPU-local) CPU
        }
        unsigned long long seen = 0;
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (!good) init_failed_ = true;
            ++ready_;
        }
        cv_done_.notify_all();
        for (;;) {
            Job job;
            // stay hot for a moment: the next chunk of the same batch / file is usually a fraction of a millisecond away
            spin_until([&] { return gen_a_.load(std::memory_order_acquire) != seen; }, 1500);
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_job_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (quit_) break;
                job = job_;
            }
            w.used = false;
            bool bad = !good;
            std::string why = "worker initialisation failed";
            if (!bad && job.after && cudaStreamWaitEvent(w.stream, job.after, 0) != cudaSuccess) { bad = true; why = "cudaStreamWaitEvent"; }
            int k = 0;
            while (!bad && t < job.active) {
                const uint64_t i = next_.fetch_add(1);
                if (i >= npieces_) break;
                const uint64_t o = i * piece_bytes(), len = std::min<uint64_t>(piece_bytes(), job.n - o);
                const int s = k & 1;
                ++k;
                if (w.inflight[s]) { cudaEventSynchronize(w.ev[s]); w.inflight[s] = false; }
                if (job.fd >= 0) {
                    uint64_t got = 0;
                    while (got < len) {
                        const ssize_t r = pread(job.fd, w.slot[s] + got, len - got, (off_t)(job.off + o + got));
                        if (r < 0 && errno == EINTR) continue;
                        if (r <= 0) { bad = true; why = r == 0 ? "unexpected end of file" : strerror(errno); break; }
                        got += (uint64_t)r;
                    }
                    if (bad) break;
                } else if (job.dst_other) {
                    // codes at the head of the slot (len/4 bytes), `other` bits behind them (len/8 bytes); a piece
                    // starts on a multiple of 64 bases, i.e. on a word of both arrays
                    unsigned char* const so = w.slot[s] + piece_bytes() / 4;
                    nk::host_pack_range(job.src + o, 0, len, reinterpret_cast<uint32_t*>(w.slot[s]), reinterpret_cast<uint32_t*>(so), 0);
                    const uint64_t cb = (len + 15) / 16 * 4, ob = (len + 31) / 32 * 4;
                    if (cudaMemcpyAsync(job.dst + o / 4, w.slot[s], cb, cudaMemcpyHostToDevice, w.stream) != cudaSuccess ||
                        cudaMemcpyAsync(job.dst_other + o / 8, so, ob, cudaMemcpyHostToDevice, w.stream) != cudaSuccess ||
                        cudaEventRecord(w.ev[s], w.stream) != cudaSuccess) { bad = true; why = "cudaMemcpyAsync"; break; }
                    w.inflight[s] = true;
                    w.used = true;
                    continue;
                } else {
                    memcpy(w.slot[s], job.src + o, len);
                }
                if (!job.dst) continue;  // host-only measurement of the read ceiling (nk_debug_stage_file)
                if (cudaMemcpyAsync(job.dst + o, w.slot[s], len, cudaMemcpyHostToDevice, w.stream) != cudaSuccess ||
                    cudaEventRecord(w.ev[s], w.stream) != cudaSuccess) { bad = true; why = "cudaMemcpyAsync"; break; }
                w.inflight[s] = true;
                w.used = true;
            }
            if (w.used) cudaEventRecord(w.done, w.stream);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (bad) { failed_ = true; err_ = why; }
                ++finished_;
                finished_a_.fetch_add(1, std::memory_order_release);
            }
            cv_done_.notify_all();
        }
        if (w.stream) cudaStreamSynchronize(w.stream);
        for (int i = 0; i < 2; ++i) {
            if (w.slot[i]) cudaFreeHost(w.slot[i]);
            if (w.ev[i]) cudaEventDestroy(w.ev[i]);
        }
        if (w.done) cudaEventDestroy(w.done);
        if (w.stream) cudaStreamDestroy(w.stream);
    }

    template <class Pred>
    static void spin_until(Pred done, long micros) {
        timespec t0;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (unsigned it = 0; !done(); ++it) {
            __builtin_ia32_pause();
            if ((it & 255u) == 255u) {
                timespec t1;
                clock_gettime(CLOCK_MONOTONIC, &t1);
                if ((t1.tv_sec - t0.tv_sec) * 1000000L + (t1.tv_nsec - t0.tv_nsec) / 1000L > micros) return;
            }
        }
    }

    int device_;
    std::atomic<unsigned long long> gen_a_{0};
    std::atomic<size_t> finished_a_{0};
    std::vector<Worker> workers_;
    std::mutex mu_;
    std::condition_variable cv_job_, cv_done_;
    unsigned long long gen_ = 0;
    bool quit_ = false, failed_ = false, init_failed_ = false;
    size_t ready_ = 0, finished_ = 0;
    Job job_{};
    std::atomic<uint64_t> next_{0};
    uint64_t npieces_ = 0;
    std::string err_;
};

static StagePool* pool_of(nk_counter* h) {
    if (!h->stage_pool) {
        unsigned n = std::thread::hardware_concurrency();
        // Measured on the 16-vCPU B200 boxes (tools/stage_sweep.py, profiles/r02_bench.md), 113 MB, 2 MiB pieces, workers
        // that stay hot between the chunks of a job.  COPY jobs (files: pread; pageable batches without packing): 3.9 /
        // 3.05 / 2.94 / 2.99 / 2.91 ms (batch) and 6.2 / 4.7 / 4.6 / 4.05 / 3.43 ms (file) with 3 / 4 / 5 / 6 / 8 threads,
        // no better beyond eight (StagePool::kCopyWorkers).  PACK jobs (pageable batches) are compute-bound per thread:
        // 3.95 / 3.17 / 3.09 / 2.40 / 2.36 / 2.38 / 2.33 ms with 4 / 6 / 8 / 10 / 12 / 14 / 15 threads — and 7.4 ms with
        // 16, when the caller's spinning thread has no core left.  Hence up to twelve workers, two cores always left
        // free, and never more than this GPU's share of the host's cores (one process per GPU on an 8-GPU box).
        int ngpu = 1;
        if (cudaGetDeviceCount(&ngpu) != cudaSuccess || ngpu < 1) { cudaGetLastError(); ngpu = 1; }
        const unsigned share = std::max(2u, n / (unsigned)ngpu);
        n = share >= 12u ? std::min(12u, share - 2u) : std::min(8u, share);
        if (const char* e = getenv("NK_STAGE_THREADS")) n = (unsigned)atoi(e);
        if (n > 32) n = 32;
        if (n < 1) n = 1;
        StagePool* p = new StagePool(h->cfg.device, n);
        if (!p->ok()) { delete p; return nullptr; }
        h->stage_pool = p;
    }
    return static_cast<StagePool*>(h->stage_pool);
}

int stage_to_device(nk_counter* h, const uint8_t* src, int fd, uint64_t off, uint64_t n, unsigned char* dst,
                    cudaEvent_t after, cudaStream_t then) {
    NvtxRange nvtx("nk:stage (host threads -> pinned slots -> H2D)");
    StagePool* p = pool_of(h);
    if (!p) return fail(NK_ERR_OOM, "cannot create the host staging pool (pinned memory / streams)");
    return p->copy(src, fd, off, n, dst, after, then);
}

// pageable ASCII bases -> packed code words + `other` bits on the device (see StagePool::copy)
int stage_pack_to_device(nk_counter* h, const uint8_t* src, uint64_t n, unsigned char* dst_codes, unsigned char* dst_other,
                         cudaEvent_t after, cudaStream_t then) {
    NvtxRange nvtx("nk:stage (host threads pack 2 bits per base -> pinned slots -> H2D)");
    StagePool* p = pool_of(h);
    if (!p) return fail(NK_ERR_OOM, "cannot create the host staging pool (pinned memory / streams)");
    return p->copy(src, -1, 0, n, dst_codes, after, then, dst_other);
}

// Packing pays from ten workers on (below that a plain copy of the ASCII bytes is faster).  A PINNED batch has the
// in-place read as its alternative, which costs the host nothing: it is packed only where this process has the host to
// itself (one GPU in the box: 1.92 against 2.37 ms) — two ranks packing at once share the host's memory bandwidth and
// end up level with the in-place read (2.50 against 2.44 ms at N = 2).
bool stage_pack_worthwhile(nk_counter* h, bool pinned_source) {
    StagePool* p = pool_of(h);
    if (!p || p->threads() < 10) return false;
    if (!pinned_source) return true;
    int ngpu = 1;
    if (cudaGetDeviceCount(&ngpu) != cudaSuccess) { cudaGetLastError(); return false; }
    return ngpu == 1;
}

void stage_pool_destroy(nk_counter* h) {
    delete static_cast<StagePool*>(h->stage_pool);
    h->stage_pool = nullptr;
}

void ingest_free(nk_counter* h) {
    stage_pool_destroy(h);
    cudaFree(h->d_raw);
    cudaFree(h->d_parse_scratch);
    cudaFree(h->d_line_end);
    cudaFree(h->d_parse_totals);
    if (h->h_parse_totals) cudaFreeHost(h->h_parse_totals);
    h->d_raw = nullptr; h->d_parse_scratch = nullptr; h->d_line_end = nullptr; h->d_parse_totals = nullptr; h->h_parse_totals = nullptr;
    h->raw_cap = h->parse_scratch_cap = h->line_end_cap = 0;
    h->fp_valid = false;
}

namespace {

int grow(void** p, unsigned long long* cap, unsigned long long need) {
    if (need <= *cap) return NK_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    const unsigned long long want = need + need / 8 + 256;
    NK_CUDA(cudaMalloc(p, want));
    *cap = want;
    return NK_OK;
}

}  // namespace

// count_pe != null: the windows are also COUNTED here (as part of a job the caller has begun) — for FASTA chunk by
// chunk while the rest of the file is still on its way; null: parse only (uniques pass, parity tap).
int parse_file_on_device(nk_counter* h, const char* path, bool* handled, bool* is_fastq, unsigned long long* nbases,
                         unsigned long long* nrec, std::string* err, PhaseEvents* count_pe) {
    *handled = false;
    if (const char* e = getenv("NK_GPU_PARSE")) if (atoi(e) == 0) return NK_OK;
    if (getenv("NK_FASTA_WINDOW")) return NK_OK;  // the tests of the host reader's parallel ingest pin that path
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return NK_OK;  // the host reader reports the error
    struct stat st;
    unsigned char magic[4] = {0, 0, 0, 0};
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 1 || ::pread(fd, magic, 4, 0) < 1 ||
        (magic[0] != '>' && magic[0] != '@')) {  // compressed, empty, not FASTA/FASTQ, a pipe: the host reader decides
        ::close(fd);
        return NK_OK;
    }
    const unsigned long long size_real = (unsigned long long)st.st_size;
    const bool fastq = magic[0] == '@';
    // raw bytes + stripped bases (+ bitmap) + FASTQ line table must fit beside the pool: else the host reader streams.
    // Only asked when buffers have to grow: cudaMemGetInfo is a resource-manager call with a long latency tail (it
    // showed up as 5-70 ms outliers of the file path when it ran for every file).
    if (size_real + 64 > h->raw_cap) {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const unsigned long long have = (unsigned long long)free_b + h->raw_cap + h->staged.bases_cap + h->line_end_cap * 8;
        if (size_real * (fastq ? 3ull : 2ull) + (size_real >> 2) + (256ull << 20) > have) { ::close(fd); return NK_OK; }
    }
    const bool ftrace = getenv("NK_FILE_TRACE") != nullptr;
    timespec ts0;
    clock_gettime(CLOCK_MONOTONIC, &ts0);
    auto since = [&]() { timespec b; clock_gettime(CLOCK_MONOTONIC, &b); return (b.tv_sec - ts0.tv_sec) * 1e3 + (b.tv_nsec - ts0.tv_nsec) * 1e-6; };
    const unsigned long long count_slice = (0xFFFFFFFFull / nk::COUNT_TILE - 1) * nk::COUNT_TILE;

    int rc = NK_OK;
    do {
        unsigned char last = '\n';
        if (::pread(fd, &last, 1, (off_t)(size_real - 1)) != 1) { rc = fail(NK_ERR_IO, "%s: read failed", path); break; }
        const bool add_nl = fastq && last != '\n';  // FASTQ: the end of the file ends the last line
        const unsigned long long size = size_real + (add_nl ? 1 : 0);
        if ((rc = grow((void**)&h->d_raw, &h->raw_cap, size + 64)) != NK_OK) break;
        if ((rc = grow(&h->d_parse_scratch, &h->parse_scratch_cap, nk::parse_scratch_bytes(size))) != NK_OK) break;
        if (!h->d_parse_totals) {
            if (cudaMalloc(&h->d_parse_totals, 8 * sizeof(unsigned long long)) != cudaSuccess ||
                cudaMallocHost(&h->h_parse_totals, 2 * 64 * sizeof(unsigned long long)) != cudaSuccess) {
                rc = fail(NK_ERR_OOM, "parse totals");
                break;
            }
        }
        h->fp_valid = false;
        // raw bytes -> device (the kernels of an earlier job may still read d_raw / staged: order behind them)
        cudaEvent_t prev = nullptr;
        if ((rc = get_event(h, &prev)) != NK_OK) break;
        if (cudaEventRecord(prev, h->stream) != cudaSuccess) { rc = fail(NK_ERR_CUDA, "cudaEventRecord"); break; }
        NvtxRange nvtx("nk:parse (FASTA/FASTQ records on the device)");
        unsigned long long nb = 0, nr = 0;
        h->last.h2d_bytes += size_real;
#define NK_B(expr) { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { rc = fail(e_ == cudaErrorMemoryAllocation ? NK_ERR_OOM : NK_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); break; } }
        unsigned long long* const carry = h->d_parse_totals + 5;
        NK_B(cudaMemsetAsync(carry, 0, 3 * sizeof(unsigned long long), h->stream));
        const unsigned long long SEGB = nk::parse_segment_bytes();
        const unsigned long long nseg_all = (size + SEGB - 1) / SEGB;
        // Overlapped FASTA job: the file arrives in super-chunks; the records of a chunk are parsed (the scan carries its
        // state from chunk to chunk) and its windows are counted while the next chunk is still being read.  A window is
        // counted once all of its k bases have been parsed: after chunk c the starts up to (bases so far) - (k-1), rounded
        // down to 16; the record that is open at the end of a chunk ends, provisionally, where the parsed bases end.
        // At most 62 chunks (their totals come back through a small pinned ring, read one chunk late).
        unsigned long long chunk = 32ull << 20;   // measured at 115 MB: 4 / 8 / 16 / 32 MiB chunks -> 12.6 / 5.8 / 6.4 / 4.6 ms per job
        if (const char* e = getenv("NK_FILE_CHUNK_MB")) { const unsigned long long t = strtoull(e, nullptr, 10); if (t >= 1 && t <= 4096) chunk = t << 20; }
        while ((size + chunk - 1) / chunk > 62) chunk *= 2;
        const bool overlapped = !fastq && count_pe != nullptr && size > 2 * chunk && !h->file_no_overlap;
        if (overlapped) {
            // upper bounds: every byte a base; records are discovered on the way (a file with more than the offsets buffer
            // holds falls back to the two-phase path below)
            if ((rc = ensure_devbuf(h->staged, size)) != NK_OK) break;
            if ((rc = ensure_offsets(&h->staged_offsets, &h->staged_offsets_cap, 1u << 20)) != NK_OK) break;
            const unsigned long long ocap = h->staged_offsets_cap;
            // an end that has not been parsed yet reads as "infinitely far": the write kernels fill in real starts (= the
            // previous record's end) as records appear, whatever is still open ends beyond every window counted so far
            NK_B(cudaMemsetAsync(h->staged_offsets, 0xFF, ocap * sizeof(unsigned long long), h->stream));
            const unsigned long long nchunks = (size + chunk - 1) / chunk;
            std::vector<cudaEvent_t> planned(nchunks, nullptr);
            unsigned long long counted = 0, seq_open = 0;   // window starts counted so far; index of the open record
            bool overflow = false;
            auto count_upto = [&](unsigned long long c, bool final) -> int {
                // totals of chunk c are on the host (its event has completed)
                const unsigned long long nbc = h->h_parse_totals[2 * c], nrc = h->h_parse_totals[2 * c + 1];
                if (nrc + 1 > ocap) { overflow = true; return NK_OK; }
                if (nrc == 0) return NK_OK;  // (cannot happen: the file starts with '>')
                unsigned long long upto = final ? nbc : (nbc >= h->cfg.k - 1 ? (nbc - (h->cfg.k - 1)) / 16 * 16 : 0);
                if (final) {  // the last record ends where the bases end
                    cudaError_t e = cudaMemcpyAsync(h->staged_offsets + nrc, h->h_parse_totals + 2 * c, sizeof(unsigned long long),
                                                    cudaMemcpyHostToDevice, h->stream);
                    if (e != cudaSuccess) return fail(NK_ERR_CUDA, "cudaMemcpyAsync(offsets end): %s", cudaGetErrorString(e));
                }
                // every record seen so far is handed to the kernels (the one that holds `counted` is somewhere among the
                // last few; empty records can pile up on one position); the short-read heuristic gets the records of
                // this range only
                h->nseq_hint = nrc - seq_open;
                while (counted < upto) {
                    const unsigned long long n = std::min(count_slice, upto - counted);
                    DevBuf view = h->staged;
                    view.bases = h->staged.bases + counted;
                    const int r = count_chunk(h, view, h->staged_offsets, 0, nrc, counted, n, n, count_pe, false);
                    if (r != NK_OK) { h->nseq_hint = 0; return r; }
                    counted += n;
                }
                h->nseq_hint = 0;
                seq_open = nrc - 1;
                return NK_OK;
            };
            cudaError_t ce = cudaSuccess;
            for (unsigned long long c = 0; c < nchunks && rc == NK_OK && !overflow; ++c) {
                const unsigned long long c0 = c * chunk, c1 = std::min(size, c0 + chunk);
                rc = stage_to_device(h, nullptr, fd, c0, std::min(c1, size_real) - c0, h->d_raw + c0, c == 0 ? prev : nullptr, h->stream);
                if (rc != NK_OK) break;
                const unsigned long long s0 = c0 / SEGB, s1 = c + 1 == nchunks ? nseg_all : c1 / SEGB;
                ce = nk::launch_fasta_plan(h->d_raw, size, s0, s1 - s0, h->d_parse_scratch, h->d_parse_totals, carry, h->stream);
                if (ce == cudaSuccess) ce = nk::launch_fasta_write(h->d_raw, size, s0, s1 - s0, h->d_parse_scratch, h->staged.bases, h->staged_offsets, ocap, h->stream);
                if (ce == cudaSuccess) ce = cudaMemcpyAsync(h->h_parse_totals + 2 * c, h->d_parse_totals, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream);
                if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&planned[c], cudaEventDisableTiming);
                if (ce == cudaSuccess) ce = cudaEventRecord(planned[c], h->stream);
                if (ce != cudaSuccess) break;
                if (c > 0) {  // the previous chunk's totals have long arrived: count its windows behind this chunk's parse
                    ce = cudaEventSynchronize(planned[c - 1]);
                    if (ce != cudaSuccess) break;
                    rc = count_upto(c - 1, false);
                }
            }
            if (rc == NK_OK && ce == cudaSuccess && !overflow) {
                ce = cudaEventSynchronize(planned[nchunks - 1]);
                if (ce == cudaSuccess) rc = count_upto(nchunks - 1, true);
            }
            if (ftrace) fprintf(stderr, "[file trace] overlapped: %llu chunks, last count enqueued at %.2f ms\n", nchunks, since());
            for (cudaEvent_t e : planned) if (e) cudaEventDestroy(e);
            if (ce != cudaSuccess) { rc = fail(NK_ERR_CUDA, "overlapped file parse: %s", cudaGetErrorString(ce)); break; }
            if (rc != NK_OK) break;
            if (overflow) {
                // more records than the offsets buffer holds: what was counted so far is void — the caller starts the job
                // again through the two-phase path (NK_ERR_STATE is caught by process_file_device)
                rc = NK_ERR_STATE;
                break;
            }
            nb = h->h_parse_totals[2 * (nchunks - 1)];
            nr = h->h_parse_totals[2 * (nchunks - 1) + 1];
        } else {
            if ((rc = stage_to_device(h, nullptr, fd, 0, size_real, h->d_raw, prev, h->stream)) != NK_OK) break;
            if (ftrace) fprintf(stderr, "[file trace] staging issued at %.2f ms\n", since());
            if (add_nl) NK_B(cudaMemsetAsync(h->d_raw + size_real, '\n', 1, h->stream));
            if (!fastq) {
                NK_B(nk::launch_fasta_plan(h->d_raw, size, 0, nseg_all, h->d_parse_scratch, h->d_parse_totals, carry, h->stream));
                NK_B(cudaMemcpyAsync(h->h_parse_totals, h->d_parse_totals, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
                NK_B(cudaStreamSynchronize(h->stream));
                nb = h->h_parse_totals[0];
                nr = h->h_parse_totals[1];
                if ((rc = ensure_devbuf(h->staged, nb)) != NK_OK) break;
                if ((rc = ensure_offsets(&h->staged_offsets, &h->staged_offsets_cap, nr + 1)) != NK_OK) break;
                NK_B(nk::launch_fasta_write(h->d_raw, size, 0, nseg_all, h->d_parse_scratch, h->staged.bases, h->staged_offsets,
                                            h->staged_offsets_cap, h->stream));
                // offsets[nrec] = number of bases (h_parse_totals[0] stays put until the next parse, which synchronises first)
                NK_B(cudaMemcpyAsync(h->staged_offsets + nr, h->h_parse_totals, sizeof(unsigned long long), cudaMemcpyHostToDevice, h->stream));
            } else {
                NK_B(nk::launch_fastq_lines(h->d_raw, size, h->d_parse_scratch, h->d_parse_totals, h->stream));
                NK_B(cudaMemcpyAsync(h->h_parse_totals + 2, h->d_parse_totals + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
                NK_B(cudaStreamSynchronize(h->stream));
                const unsigned long long nlines = h->h_parse_totals[2];
                unsigned long long le_bytes = h->line_end_cap * 8;
                if ((rc = grow((void**)&h->d_line_end, &le_bytes, (nlines + 1) * 8)) != NK_OK) break;
                h->line_end_cap = le_bytes / 8;
                // sequence lines are at most half of a well-formed file; a malformed one may keep more (dropped later)
                if ((rc = ensure_devbuf(h->staged, size)) != NK_OK) break;
                if ((rc = ensure_offsets(&h->staged_offsets, &h->staged_offsets_cap, nlines / 4 + 2)) != NK_OK) break;
                NK_B(nk::launch_fastq_write(h->d_raw, size, size_real, h->d_parse_scratch, h->staged.bases, h->staged_offsets, h->d_line_end,
                                            nlines, h->d_parse_totals, h->stream));
                NK_B(cudaMemcpyAsync(h->h_parse_totals, h->d_parse_totals, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
                NK_B(cudaStreamSynchronize(h->stream));
                nb = h->h_parse_totals[0];
                nr = h->h_parse_totals[1];
            }
            if (count_pe) {  // the whole file is parsed: count it like a staged batch
                for (unsigned long long c0 = 0; c0 < nb && nr > 0 && rc == NK_OK; c0 += count_slice) {
                    const unsigned long long n = std::min(count_slice, nb - c0);
                    DevBuf view = h->staged;
                    view.bases = h->staged.bases + c0;
                    rc = count_chunk(h, view, h->staged_offsets, 0, nr, c0, n, n, count_pe, false);
                }
                if (rc != NK_OK) break;
            }
        }
#undef NK_B
        *nbases = nb;
        *nrec = nr;
        *is_fastq = fastq;
        *handled = true;
        h->fp_valid = true;
        h->fp_path = path;
        h->fp_size = size_real;
        h->fp_mtime_ns = (unsigned long long)st.st_mtim.tv_sec * 1000000000ull + (unsigned long long)st.st_mtim.tv_nsec;
        h->fp_nbases = nb;
        h->fp_nrec = nr;
    } while (0);
    ::close(fd);
    if (rc != NK_OK && err) *err = g_err;
    return rc;
}

}  // namespace nkd

extern "C" {

// Measurement tap: how fast can this host hand the file's bytes over at all?  host_only_ms: the staging pool
// preads the whole file into its pinned slots and drops the bytes (no device work); with_h2d_ms: the same with
// the H2D copies into the raw buffer, waited for.  The file path of nk_process_file cannot beat the second.
int nk_debug_stage_file(nk_counter* h, const char* path, double* host_only_ms, double* with_h2d_ms) {
    if (!h || !path || !host_only_ms || !with_h2d_ms) return fail(NK_ERR_BAD_ARG, "null argument");
    if (is_group(h)) return nk_debug_stage_file(h->group[0], path, host_only_ms, with_h2d_ms);
    NK_CUDA(cudaSetDevice(h->cfg.device));
    NK_TRY(resolve(h));
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return fail(NK_ERR_IO, "cannot open %s", path);
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 1) { ::close(fd); return fail(NK_ERR_IO, "%s: empty or unreadable", path); }
    const unsigned long long size = (unsigned long long)st.st_size;
    int rc = grow((void**)&h->d_raw, &h->raw_cap, size + 64);
    h->fp_valid = false;
    auto now = [] { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; };
    double best0 = 1e30, best1 = 1e30;
    for (int rep = 0; rep < 4 && rc == NK_OK; ++rep) {
        cudaStreamSynchronize(h->stream);
        double t0 = now();
        rc = stage_to_device(h, nullptr, fd, 0, size, nullptr, nullptr, h->stream);
        double t1 = now();
        if (rc != NK_OK) break;
        rc = stage_to_device(h, nullptr, fd, 0, size, h->d_raw, nullptr, h->stream);
        cudaStreamSynchronize(h->stream);
        double t2 = now();
        if (rep) { best0 = std::min(best0, t1 - t0); best1 = std::min(best1, t2 - t1); }
    }
    ::close(fd);
    *host_only_ms = best0;
    *with_h2d_ms = best1;
    return rc;
}

// Parity tap: parse a plain FASTA / FASTQ file ON THE DEVICE and digest what comes out exactly like
// nk_debug_fastx_digest digests the host reader's records (FNV-1a-64 over every record's bases + one 0xFF).
int nk_debug_parse_file(nk_counter* h, const char* path, uint64_t* nrecords, uint64_t* nbases, uint64_t* fnv1a) {
    if (!h || !path) return fail(NK_ERR_BAD_ARG, "null argument");
    if (is_group(h)) return nk_debug_parse_file(h->group[0], path, nrecords, nbases, fnv1a);
    NK_CUDA(cudaSetDevice(h->cfg.device));
    NK_TRY(resolve(h));
    NK_CUDA(cudaStreamSynchronize(h->stream));
    bool handled = false, fq = false;
    unsigned long long nb = 0, nr = 0;
    std::string err;
    NK_TRY(parse_file_on_device(h, path, &handled, &fq, &nb, &nr, &err, nullptr));
    if (!handled) return fail(NK_ERR_UNSUPPORTED, "%s: not a plain regular FASTA/FASTQ file the device parser takes", path);
    std::vector<uint8_t> bases(nb ? nb : 1);
    std::vector<uint64_t> offs(nr + 1);
    NK_CUDA(cudaMemcpyAsync(bases.data(), h->staged.bases, nb, cudaMemcpyDeviceToHost, h->stream));
    NK_CUDA(cudaMemcpyAsync(offs.data(), h->staged_offsets, (nr + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
    NK_CUDA(cudaStreamSynchronize(h->stream));
    uint64_t hsh = 0xcbf29ce484222325ull;
    if (offs[0] != 0 || offs[nr] != nb) return fail(NK_ERR_CUDA, "device parser: inconsistent offsets (%llu .. %llu of %llu)",
                                                    (unsigned long long)offs[0], (unsigned long long)offs[nr], nb);
    for (uint64_t r = 0; r < nr; ++r) {
        if (offs[r + 1] < offs[r]) return fail(NK_ERR_CUDA, "device parser: offsets decrease at record %llu", (unsigned long long)r);
        if (fnv1a) {
            for (uint64_t i = offs[r]; i < offs[r + 1]; ++i) { hsh ^= bases[i]; hsh *= 0x100000001b3ull; }
            hsh ^= 0xFFu; hsh *= 0x100000001b3ull;
        }
    }
    if (nrecords) *nrecords = nr;
    if (nbases) *nbases = nb;
    if (fnv1a) *fnv1a = hsh;
    return NK_OK;
}

}  // extern "C"
