// nk_internal.h — the handle behind the C ABI and the host-side helpers shared by nk_api.cu (single-GPU
// entry points) and nk_multi.cu (single-process multi-GPU groups).  Not installed: the public surface is
// include/neurokmer.h only.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/neurokmer.h"
#include "nk_host.h"
#include "nk_kernels.cuh"

namespace nkd {

// NVTX range over a host-side phase (header-only NVTX v3: a no-op unless a profiler is attached).  The ranges
// bracket the ENQUEUE of a phase's work; nsys / ncu correlate the kernels launched inside them.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

extern thread_local std::string g_err;
int fail(int code, const char* fmt, ...);


#define NK_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(e_ == cudaErrorMemoryAllocation ? NK_ERR_OOM : NK_ERR_CUDA, "%s: %s", #expr, \
                        cudaGetErrorString(e_));                                               \
    } while (0)

#define NK_TRY(expr)              \
    do {                          \
        int rc_ = (expr);         \
        if (rc_ != NK_OK) return rc_; \
    } while (0)

constexpr size_t kMaxTimedChunks = 1024;
constexpr unsigned long long kChunkBytes = 32ull << 20;  // host->device pipeline granule (multiple of COUNT_TILE)
static_assert(kChunkBytes % nk::COUNT_TILE == 0, "chunks must be whole tiles");

struct DevBuf {
    unsigned char* bases = nullptr;
    unsigned long long bases_cap = 0;
    unsigned int* invalid = nullptr;
    unsigned long long invalid_cap = 0;  // words
    cudaEvent_t copy_done = nullptr, compute_done = nullptr;
    // pre-packed input (nk_*_packed): 2-bit code words and `other` bits of the chunk
    unsigned char* codes = nullptr;
    unsigned long long codes_cap = 0;
    unsigned char* other = nullptr;
    unsigned long long other_cap = 0;
    bool has_other = false;  // this chunk's `other` array was supplied
    // zero-copy views: readable bytes (multiples of 16) of bases/codes and other from the view's start; 0 = padded
    unsigned long long bases_bytes = 0, other_bytes = 0;
};

// how a host batch push ends: the file driver double-buffers its own pinned batches (never blocks, never reads
// in place); a plain push returns when the caller's buffers are reusable; a deferred push (multi-GPU fan-out)
// leaves that wait to the caller so that every GPU's work is enqueued before anything blocks
enum PushSync { kPushFileDriver = 0, kPushSync = 1, kPushDeferred = 2 };

struct PhaseEvents {
    std::vector<cudaEvent_t> mark0, count0, count1;
    cudaEvent_t begin = nullptr, copy0 = nullptr, copy1 = nullptr, fold0 = nullptr, fold1 = nullptr,
                lif1 = nullptr, end = nullptr;
};

}  // namespace nkd

using nkd::DevBuf;
using nkd::PhaseEvents;

struct nk_counter {
    nk_config cfg{};
    nk::FastMod fm{};
    cudaStream_t stream = nullptr, copy_stream = nullptr;

    // pool state (device)
    unsigned int* acc = nullptr;           // u32 batch accumulators (RED target)
    unsigned long long* currents = nullptr;
    float* v = nullptr;
    unsigned int* r = nullptr;
    unsigned long long* spikes = nullptr;
    unsigned long long* scalars = nullptr;  // [0] spikes fired by last LIF, [1] max cumulative spikes, [2] kmers
    unsigned int* tile_counter = nullptr;
    unsigned long long* h_scalars = nullptr;  // pinned mirror

    // carried-state LIF by memoisation (allocated at first use)
    nk::LifMemo memo{};
    // LIF per-count table
    nk::LifTable table{};
    unsigned long long table_cap = 0;

    // top-N scratch
    nk::TopNScratch topn{};
    unsigned long long topn_cap = 0;
    nk_top_entry* h_top = nullptr;  // pinned, topn_cap rows (built from two arrays)

    // host mirrors (EnergyTracker, src/models.rs:145-173)
    unsigned long long total_spikes = 0, energy_fixed = 0;
    bool fresh = true;      // every neuron still has v = 0, r = 0
    // after nk_reset the pool arrays (currents, v, r, spikes) are LOGICALLY zero but not yet written:
    // the fused fold+LIF kernel of the next job overwrites all four, anything else materialises first
    bool lazy_zero = false;
    int force_direct = 0;
    bool streaming = false;
    bool acc_dirty = false;
    unsigned long long acc_kmers = 0;  // windows added to acc since the last fold (u32 overflow guard)
    bool currents_valid_overwrite = true;  // next fold overwrites currents (first fold of a call)
    // u32 overflow guard: fold `acc` away before more than fold_limit window starts could have been added to it
    // (nk_debug_set_fold_limit lowers it for tests).  On a handle that is part of a sharded pool the counts go to
    // the SPILL array behind the mailbox (peers read it) instead of into this rank's private `currents`.
    unsigned long long fold_limit = 0xFFFFFFFFull;
    unsigned long long* spill = nullptr;   // pool_size u64, tail of the `acc` allocation
    bool spill_dirty = false;

    // staging
    DevBuf buf[2];
    int cur_buf = 0;
    // offsets of host batches: two buffers alternate so that batch i+1's copy never waits for batch i's kernels
    unsigned long long* d_offsets2[2] = {nullptr, nullptr};
    unsigned long long offsets_cap2[2] = {0, 0};
    cudaEvent_t offsets_done[2] = {nullptr, nullptr};
    int cur_off = 0;
    // device-resident staged batch (nk_stage_reserve)
    DevBuf staged;
    uint8_t* file_batch[2] = {nullptr, nullptr};  // pinned, 32 MiB each: the file driver's double buffer
    DevBuf zc;  // zero-copy pushes: just the invalid-start bitmap of the body (the bases stay in host memory)
    unsigned long long* staged_offsets = nullptr;
    unsigned long long staged_offsets_cap = 0;

    std::vector<cudaEvent_t> evpool;
    size_t ev_used = 0;
    nk_timings last{};

    // a process/stream call returns with its read-back (new spikes, k-mers) and event timings
    // still in flight on `stream`; resolve() waits for them the first time anything observes them
    bool pending = false, pending_lif = false, pending_timings = false;
    PhaseEvents pend_pe;
    PhaseEvents stream_pe;  // mark/count event pairs of the pushes between stream_begin and stream_finish
    // host-side upper bound of the largest cumulative spike count (bounds the top-N radix passes
    // without a device round trip): each LIF call adds at most ceil(steps / (refractory + 1))
    unsigned long long spike_bound = 0;
    // fused post kernel (fold + LIF table + top-N): scratch, result pack, cached rows
    int post_grid = 0;
    unsigned long long* post_zero = nullptr;   // [8 u64 ctrl][8*256 u32 hist] zeroed before each launch
    unsigned long long* d_pack = nullptr;      // PACK_MAX_U64
    unsigned long long* h_pack = nullptr;      // pinned mirror
    unsigned long long* h_pack_dev = nullptr;  // the same buffer as the device sees it (mapped): single-GPU jobs write it directly
    bool pack_direct = false;                  // the pending pack was written into h_pack by the kernel (no D2H copy)
    unsigned long long topn_hint = 20;         // rows computed speculatively by the fused kernel (CLI: 20)
    unsigned long long top_cached_n = 0;       // rows of the last fused launch (valid until state changes)
    bool top_cache_valid = false, pending_pack = false;
    // multi-GPU sharded-pool mode (nk_dist_*): peer mappings of every rank's accumulators
    int dist_rank = 0, dist_world = 0;
    const unsigned int* dist_peer[16] = {};
    bool dist_ipc_opened[16] = {};
    unsigned long long dist_lo = 0, dist_len = 0, dist_n_each = 0;
    unsigned long long* d_merged = nullptr;
    // peer-signalled mode (nk_dist_run): every rank's mailbox (tail of its accumulator allocation)
    unsigned char* dist_mail[16] = {};
    unsigned long long dist_epoch = 0;
    bool dist_failed = false;
    // ---- host -> device staging by a pool of host threads, and record parsing on the device (nk_ingest.cu) ----
    void* stage_pool = nullptr;             // nkd::StagePool*, created at first use
    unsigned char* d_raw = nullptr;         // raw file bytes
    unsigned long long raw_cap = 0;
    void* d_parse_scratch = nullptr;
    unsigned long long parse_scratch_cap = 0;
    unsigned long long* d_line_end = nullptr;  // FASTQ: position of every newline
    unsigned long long line_end_cap = 0;
    unsigned long long* d_parse_totals = nullptr;  // 8 u64
    unsigned long long* h_parse_totals = nullptr;  // pinned mirror
    // the parsed file that still sits in `staged` / `staged_offsets` (the uniques pass re-uses it instead of a second read)
    bool file_no_overlap = false;           // retry of a file whose records outgrew the overlapped path's offsets buffer
    unsigned long long nseq_hint = 0;       // sequences that actually lie in the range being counted (heuristics only)
    bool fp_valid = false;
    std::string fp_path;
    unsigned long long fp_size = 0, fp_mtime_ns = 0, fp_nbases = 0, fp_nrec = 0;
    // scalars[0] (spikes fired) / scalars[2] (k-mers) are known to be zero on the device: the fused post kernel
    // zeroes what it consumed, so a job of resident kernels needs no memset at all
    bool fired_clean = false, kmers_clean = false;
    bool uniques_whole_input = false;  // group[0] of a multi-GPU group: its uniques pass is handed the whole input again
    bool slice_only = false;  // after a sharded-pool job: currents / v / r / spikes are only defined inside this rank's slice
    bool last_push_zc = false;  // the last host batch was read in place (zero-copy): its kernels hold the caller's buffer
    // ---- single-process multi-GPU group (nk_create_multi, nk_multi.cu): this handle is the LEADER, a thin
    // dispatcher without device state of its own; group[r] is an ordinary single-GPU handle on devices[r]
    std::vector<nk_counter*> group;
    int group_state = 0;                    // where the neuron state lives: 0 nowhere yet (fresh), 1 sliced, 2 on group[0]
    bool group_streaming = false;
    bool group_counted = false;             // at least one window start was pushed since the stream began
    bool group_can_peer = true;
    unsigned long long* m_gathered = nullptr;  // device of group[0]: one result pack slot per member
    std::vector<cudaEvent_t> ev_counted, ev_posted;
    cudaEvent_t ev_leader = nullptr;
    std::vector<std::vector<uint64_t>> shard_offsets;  // per member: piece offsets of the batch being pushed
    nk_timings group_last{};
    bool dist_job = false;   // the pending result pack comes from a sharded-pool job (its time stamps describe the exchange)
    // exact side tables (opt-in, nk_enable_exact_counts)
    bool exact = false;
    nk::ExactTable xt;
    // member of a multi-GPU group: the merged table of THIS GPU's neuron slice (xt holds its own windows' table)
    nk::ExactTable xs;
    // group leader: kmer_per_neuron of the whole pool on the host (for the `uniques` column), filled on demand
    std::vector<uint32_t> group_uniques;
    bool group_uniques_valid = false;
    unsigned int* d_top_uniques = nullptr;
    // uniques pass (nk_uniques_*): `uniques` of the top rows by a second pass over the input, without the
    // O(windows) exact table — the words that map to the rows' neurons are collected, sorted and counted
    nk::ExactTable ut;
    unsigned int* d_filter = nullptr;        // pool_size bits: neurons of the fixed rows
    unsigned long long* d_rows = nullptr;    // their indices (device), row order
    unsigned long long d_rows_cap = 0;
    std::vector<unsigned long long> row_idx; // rows fixed by nk_uniques_begin
    std::vector<unsigned int> row_uniques;   // filled by nk_uniques_end
    bool rows_valid = false, uniques_open = false;
    unsigned long long ut_count = 0;         // host mirror of the append cursor
    unsigned long long file_uniques = 0;     // nk_set_file_uniques: rows nk_process_file resolves by re-reading the file
    bool table_valid = false, table_inflight = false;
    cudaEvent_t table_ready = nullptr;
    nk_config table_cfg{};
    unsigned long long table_n = 0;
};

namespace nkd {

// ---- helpers defined in nk_api.cu, shared with nk_multi.cu -------------------------------------------
int get_event(nk_counter* h, cudaEvent_t* out);
int materialize_zero(nk_counter* h);
int fold_now(nk_counter* h);
unsigned char* own_mail(nk_counter* h);
int zero_kmers(nk_counter* h);
int count_host_batch(nk_counter* h, const uint8_t* bases, const uint64_t* offsets, uint64_t nseq, PhaseEvents* pe, int sync);
int count_host_batch_packed(nk_counter* h, const uint32_t* codes, const uint32_t* other, const uint64_t* offsets,
                            uint64_t nseq, PhaseEvents* pe, int sync);
unsigned long long saturation_count(const nk_config& c);
int simulate(nk_counter* h, bool skip_zero, bool with_topn);
int finish_call(nk_counter* h, bool had_lif, const PhaseEvents* pe);
int resolve(nk_counter* h);
void begin_call(nk_counter* h);
void collect_timings(nk_counter* h, const PhaseEvents& pe);
float ev_ms(cudaEvent_t a, cudaEvent_t b);
int fold_and_simulate(nk_counter* h, bool skip_zero, PhaseEvents& pe, bool with_topn);
int dist_build_post(nk_counter* h, nk::PostParams& q, unsigned long long* n_top_out);
unsigned long long* merged_pack_out(nk_counter* h);
int dist_finish(nk_counter* h, unsigned long long n_out);
int copy_out(nk_counter* h, void* dst, const void* src, size_t bytes);

// ---- parallel staging + device-side record parsing (nk_ingest.cu) ---------------------------------------
// host memory [src, src+n) (fd < 0) or file bytes [off, off+n) of fd -> device dst, by the handle's pool of
// host threads (memcpy / pread into pinned slots, async H2D on per-thread streams).  Returns when every
// copy has been ENQUEUED and the source bytes have been consumed; `after` (may be null): the copies wait
// for this event on the device; `then`: this stream waits for all of them.
int stage_to_device(nk_counter* h, const uint8_t* src, int fd, uint64_t off, uint64_t n, unsigned char* dst,
                    cudaEvent_t after, cudaStream_t then);
int stage_pack_to_device(nk_counter* h, const uint8_t* src, uint64_t n, unsigned char* dst_codes, unsigned char* dst_other,
                         cudaEvent_t after, cudaStream_t then);
bool stage_pack_worthwhile(nk_counter* h, bool pinned_source);
void stage_pool_destroy(nk_counter* h);
// Whole plain FASTA / FASTQ file -> h->staged (bases) + h->staged_offsets, parsed on the device.
// *handled = false (and NK_OK): this path does not apply (compressed, not a regular file, too large, disabled)
// count_pe != null: the caller has begun a job and the windows are counted here as well (FASTA: overlapped with the read)
int parse_file_on_device(nk_counter* h, const char* path, bool* handled, bool* is_fastq, unsigned long long* nbases,
                         unsigned long long* nrec, std::string* err, PhaseEvents* count_pe);
void ingest_free(nk_counter* h);
int uniques_chunk(nk_counter* h, DevBuf& b, const unsigned long long* d_offsets, unsigned long long seq_lo,
                  unsigned long long seq_hi, unsigned long long origin, unsigned long long nstarts, bool packed);
int count_chunk(nk_counter* h, DevBuf& b, const unsigned long long* d_offsets, unsigned long long seq_lo,
                unsigned long long seq_hi, unsigned long long origin, unsigned long long nstarts,
                unsigned long long max_windows, PhaseEvents* pe, bool packed);
int ensure_devbuf(DevBuf& b, unsigned long long nbytes);
int ensure_offsets(unsigned long long** p, unsigned long long* cap, unsigned long long n);

// ---- single-process multi-GPU groups (nk_multi.cu) ---------------------------------------------------
inline bool is_group(const nk_counter* h) { return h && !h->group.empty(); }
int group_unsupported(const char* what);
int group_destroy(nk_counter* g);
int group_reset(nk_counter* g);
int group_set_steps(nk_counter* g, uint64_t steps);
int group_begin(nk_counter* g);
// ASCII when bases != null, else the pre-packed form
int group_push(nk_counter* g, const uint8_t* bases, const uint32_t* codes, const uint32_t* other, const uint64_t* offsets,
               uint64_t nseq);
int group_end(nk_counter* g, bool skip_zero);
int group_process_batch(nk_counter* g, const uint8_t* bases, const uint32_t* codes, const uint32_t* other,
                        const uint64_t* offsets, uint64_t nseq);
int group_simulate(nk_counter* g);
int group_top_n(nk_counter* g, uint64_t top_n, nk_top_entry* out, uint64_t* n_out);
int group_copy(nk_counter* g, int which, void* out);  // 0 currents, 1 spike counts, 2 voltages, 3 refractory ticks
int group_timings(nk_counter* g, nk_timings* out);
int group_synchronize(nk_counter* g);
// exact side tables of a group: every GPU builds the table of its own windows, GPU d merges the records of its neuron slice
int group_enable_exact(nk_counter* g, int on);
int group_get_count(nk_counter* g, uint64_t kmer, uint32_t* count, int32_t* found);
int group_exact_table_size(nk_counter* g, uint64_t* n);
int group_copy_exact_table(nk_counter* g, uint64_t* keys, uint32_t* counts);
int group_copy_uniques(nk_counter* g, uint32_t* out);
// shared with the single-GPU entry points (nk_api.cu)
int exact_lookup_in(nk_counter* h, nk::ExactTable& t, uint64_t kmer, uint32_t* count, int32_t* found);
int exact_copy_table_of(nk_counter* h, nk::ExactTable& t, uint64_t* keys, uint32_t* counts);

}  // namespace nkd
