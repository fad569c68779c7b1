// nk_kernels.cuh — host-callable launchers of the sm_100a kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "nk_device.cuh"

namespace nk {

// ---- geometry of the counting kernel ---------------------------------------
// A CTA of 8 warps consumes tiles of COUNT_TILE consecutive window-start
// positions.  Warp w owns the span [w*SPAN, (w+1)*SPAN) of the tile and walks it
// in chunks of 512 positions: lane L loads the 16 bytes [16L, 16L+16) of the
// chunk (one LDS.128 from the TMA-staged tile), turns them into 2-bit code words,
// and receives the words of lanes L+1, L+2 by shuffle (the k-1 <= 31 overlap).
#ifndef NK_COUNT_THREADS
#define NK_COUNT_THREADS 256
#endif
constexpr int COUNT_THREADS = NK_COUNT_THREADS;
constexpr int COUNT_WARPS = COUNT_THREADS / 32;
constexpr int COUNT_CHUNK = 512;                         // positions per warp iteration
#ifndef NK_CHUNKS_PER_SPAN
#define NK_CHUNKS_PER_SPAN 1
#endif
constexpr int COUNT_CHUNKS_PER_SPAN = NK_CHUNKS_PER_SPAN;
constexpr int COUNT_SPAN = COUNT_CHUNK * COUNT_CHUNKS_PER_SPAN;
constexpr int COUNT_TILE = COUNT_WARPS * COUNT_SPAN;     // 4096 positions: small tiles keep the grid's tail short
constexpr int COUNT_HALO = 32;                           // >= k-1, multiple of 16
constexpr int COUNT_STAGES = 2;

struct CountParams {
    const unsigned char* bases;   // device, 16-B aligned, readable up to ntiles*TILE + HALO
                                  // packed != 0: the 2-bit code words, readable up to ntiles*TILE/4 + 16
    const unsigned char* other;   // packed only: `other` bits, readable up to ntiles*TILE/8 + 16; may be null
    int packed;                   // 0: ASCII bases, 1: pre-packed input (include/neurokmer.h, "nk2" layout)
    // Readable bytes of `bases` / `other`, multiples of 16; 0 = padded device buffers, every tile copy is in
    // bounds.  Set for zero-copy input (the caller's host arrays end where they end): the last tile's bulk
    // copies are clamped to them.  A 16-byte aligned chunk never straddles a page, so rounding the array size
    // up to 16 stays inside the page that holds its last byte.
    unsigned long long bases_bytes;
    unsigned long long other_bytes;
    const unsigned int* invalid;  // bit p set => no window starts at p; ntiles*TILE/32 words (+pad)
    unsigned int* acc;            // pool_size u32 batch accumulators
    unsigned int* tile_counter;   // [0] dynamic tile scheduler cursor, [1] CTAs that finished: both zero before a
                                  // launch; the last CTA to leave zeroes them again (no memset between launches)
    // MODE 5 (long sequences, no invalid-start bitmap): the kernel finds the sequence ends itself.
    //   offsets[0..nseq]: batch-absolute sequence bounds of the sequences that overlap this chunk;
    //   window starts [origin, origin + nstarts) belong to the chunk (tile 0 starts at `origin`);
    //   *kmers_out += number of valid window starts counted.
    const unsigned long long* offsets;
    unsigned long long nseq, origin, nstarts;
    unsigned long long* kmers_out;
    unsigned long long ntiles;
    FastMod fm;
    RotMul rm;                    // make_rotmul(): opaque 2^B multipliers (see nk_device.cuh)
    unsigned int k;
    // exact side table (MODE 2 instantiation only): every counted word is appended here
    unsigned long long* words;
    unsigned int* widx;           // the neuron index of every appended word, parallel to `words`
    unsigned long long* words_cursor;
    // uniques pass (MODE 4 instantiation only): nothing is counted; the words of the windows whose neuron
    // has its bit set in `filter` (pool_size bits) are appended while the cursor stays below words_cap
    // (the cursor keeps counting past it, so the host learns how much room a re-run needs)
    const unsigned int* filter;
    unsigned long long words_cap;
    // debug taps (EMIT instantiation only); any may be null
    unsigned long long* out_fwd;
    unsigned long long* out_rc;
    unsigned long long* out_word;
    unsigned long long* out_idx;
};

size_t count_smem_bytes();
// size a bases buffer must have so that every tile copy is in bounds
inline unsigned long long count_ntiles(unsigned long long nbytes) {
    return (nbytes + COUNT_TILE - 1) / COUNT_TILE;
}
inline unsigned long long count_padded_bases(unsigned long long nbytes) {
    return count_ntiles(nbytes) * COUNT_TILE + COUNT_HALO + 64;
}
// pre-packed input: device bytes of the code / `other` arrays of a chunk of nbytes window starts
inline unsigned long long count_padded_codes(unsigned long long nbytes) {
    return count_ntiles(nbytes) * (COUNT_TILE / 4) + 16 + 64;
}
inline unsigned long long count_padded_other(unsigned long long nbytes) {
    return count_ntiles(nbytes) * (COUNT_TILE / 8) + 16 + 64;
}
inline unsigned long long count_bitmap_words(unsigned long long nbytes) {
    return count_ntiles(nbytes) * (COUNT_TILE / 32) + 16;
}

// mode 0: count; 1: emit the debug taps instead of counting; 2: count and append words (exact side
// table); 3: count with warp-level compaction of the valid window starts (short-read batches);
// 4: uniques pass (no counting: append the words that map to the neurons of `filter`);
// 5: count WITHOUT the invalid-start bitmap (long sequences: each tile looks its sequence ends up in `offsets`).
// The grid is persistent: min(ntiles, SMs x resident CTAs of the instantiation).
cudaError_t launch_count(const CountParams& p, bool canonical, int mode, cudaStream_t s);

// invalid-start bitmap: zero, then mark the last k-1 starts of every sequence
// in [seq_lo, seq_hi) (positions relative to `origin`) and everything in
// [nbytes, ntiles*TILE).  *kmers (device) += windows starting inside the chunk.
cudaError_t launch_mark_invalid(unsigned int* invalid, const unsigned long long* offsets,
                                unsigned long long seq_lo, unsigned long long seq_hi,
                                unsigned long long origin, unsigned long long nbytes, unsigned k,
                                unsigned long long* kmers, cudaStream_t s, uint64_t* launches);

// acc (u32) -> currents (u64): currents[i] = (overwrite ? 0 : currents[i]) + acc[i]; acc[i] = 0
cudaError_t launch_fold(unsigned int* acc, unsigned long long* currents, unsigned long long pool,
                        bool overwrite, cudaStream_t s);

// spill (u64) -> currents (u64): currents[i] = (overwrite ? 0 : currents[i]) + spill[i]
cudaError_t launch_fold64(const unsigned long long* spill, unsigned long long* currents, unsigned long long pool,
                          bool overwrite, cudaStream_t s);

// multi-GPU group, carried-state path: currents[i] = (overwrite ? 0 : currents[i]) + sum over members of their u32
// counts (+ u64 spill where spill_mask has the member's bit); *kmers_out = sum of the members' k-mer counters
struct PeerSumParams {
    const unsigned int* acc[16];
    const unsigned long long* spill[16];
    const unsigned long long* kmers[16];
    unsigned spill_mask;
    int n;
};
cudaError_t launch_peer_sum(const PeerSumParams& ps, unsigned long long* currents, unsigned long long pool, bool overwrite,
                            unsigned long long* kmers_out, cudaStream_t s);

struct LifParams {
    unsigned long long* currents;
    unsigned int* acc;               // u32 batch accumulators to fold in first (fold_mode != 0); zeroed
    int fold_mode;                   // 0: currents as stored; 1: currents += acc; 2: currents = acc (overwrite)
    int zero_state;                  // 1: v, r, spikes are logically zero and must be WRITTEN for every neuron, never read
    float* v;
    unsigned int* r;
    unsigned long long* spikes;
    unsigned long long* total_new;   // += spikes fired by this launch
    unsigned long long* max_spikes;  // max over neurons of cumulative spike count (atomicMax)
    unsigned long long pool;
    unsigned long long steps;
    float thr, leak;
    unsigned int period;
    int skip_zero;  // 1: in-memory driver, 0: streaming (SIMD-semantics) driver
};
// direct simulation: one thread per neuron, `steps` ticks
cudaError_t launch_lif(const LifParams& p, cudaStream_t s);
// uniform-fresh-state fast path: simulate once per distinct count value (table of
// `table_n` entries, counts >= table_n-1 behave like table_n-1), then one lookup per neuron.
struct LifTable {
    unsigned int* spikes;  // table_n
    float* v;              // table_n
    unsigned int* r;       // table_n
};
cudaError_t launch_lif_table_build(const LifParams& p, const LifTable& t, unsigned long long table_n, cudaStream_t s);
cudaError_t launch_lif_table_apply(const LifParams& p, const LifTable& t, unsigned long long table_n, cudaStream_t s);
// Carried-state fast path: memoised simulation.  A neuron's evolution over a call depends only on
// (voltage, refractory ticks, count): after a fresh job the state is itself a function of the first
// count, so a pool of millions of neurons holds only a few thousand distinct (state, count) keys.  The keys
// are deduplicated in a device hash set, each distinct key is simulated ONCE, and every neuron looks its
// result up.  If the pool turns out to be diverse (more than LIF_MEMO_SLOTS/2 distinct keys) nothing is
// applied and the direct kernel — launched behind it with `only_if` set — does the work instead.
constexpr unsigned long long LIF_MEMO_SLOTS = 1ull << 20;
struct LifMemo {
    unsigned long long* keys;      // LIF_MEMO_SLOTS, all-ones = empty
    unsigned int* dense;           // LIF_MEMO_SLOTS / 2: occupied slots, compacted
    float* res_v;                  // per slot
    unsigned int* res_r;
    unsigned int* res_f;
    unsigned int* slot_of;         // pool entries: the neuron's slot (all-ones: neuron skipped)
    unsigned long long* ctrl;      // [0] distinct keys, [1] overflow flag, [2] dense cursor
};
// counts >= sat share one trajectory (needs leak >= 0: see nk_lif.cu); requires p.period < 2^31, sat < 2^21
cudaError_t launch_lif_memo(const LifParams& p, const LifMemo& m, unsigned long long sat, cudaStream_t s, uint64_t* launches);
// the direct kernel, running only if *only_if != 0 (device-side fallback of the memo path; p.fold_mode must be 0)
cudaError_t launch_lif_if(const LifParams& p, const unsigned long long* only_if, cudaStream_t s);
// one LIF tick per neuron with the raw count as input; currents zeroed (process_sequence)
cudaError_t launch_lif_single_tick(const LifParams& p, unsigned long long* currents_rw, cudaStream_t s);

// top-N by (spikes desc, idx asc).  See nk_topn.cu.
struct TopNScratch {
    unsigned int* hist;              // 256 bins
    unsigned long long* ctrl;        // control block, 8 u64
    unsigned int* block_counts;      // ceil(pool/TOPN_BLOCK_ITEMS)+1
    unsigned long long* out_idx;     // capacity >= n_cap
    unsigned long long* out_spikes;  // capacity >= n_cap
};
constexpr int TOPN_BLOCK_ITEMS = 4096;
constexpr int POST_EXACT_BINS = 2048;  // = the 8 x 256 radix bins of PostParams::hist
constexpr int POST_MAX_GRID = 1024;   // >= 148 SMs x 6 resident CTAs
constexpr size_t POST_SCRATCH_BYTES = 8 * sizeof(unsigned long long) + 8 * 256 * sizeof(unsigned int) + POST_MAX_GRID * sizeof(unsigned int);
constexpr int POST_SEG_ITEMS = 1024;  // segment of the fused post kernel's ordered phases (block_counts is sized for it)
constexpr unsigned long long TOPN_MAX_N = 1ull << 20;
cudaError_t launch_topn(const unsigned long long* spikes, unsigned long long pool, unsigned long long n,
                        unsigned long long max_spikes, const TopNScratch& sc, cudaStream_t s,
                        uint64_t* launches);

cudaError_t launch_hash_words(const unsigned long long* words, unsigned long long n, FastMod fm,
                              unsigned long long* hashes, unsigned long long* idx, cudaStream_t s);

// fused fold + LIF(table) + top-N, one cooperative launch (nk_post.cu)
struct PostParams {
    LifParams lif;
    LifTable table;
    unsigned long long table_n;
    unsigned long long n;            // rows wanted, 1..2048 and <= pool
    int passes;                      // radix digits to visit (from the host-side spike bound)
    int single_pass;                 // 1: every total is < POST_EXACT_BINS: one exact histogram instead of radix passes
    // scratch, all-zero at launch and left all-zero by the kernel (POST_SCRATCH_BYTES, see post_scratch_*):
    unsigned int* hist;              // 8 * 256 bins
    unsigned long long* ctrl;        // 8 u64: [0] gather cursor, [1] error, [2] finished blocks, [4..6] block 0's time stamps
    unsigned int* block_ties;        // POST_MAX_GRID: per block, (its number of ties) + 1 once published
    int trace;                       // NK_POST_TRACE: extra time stamps in pack[8..11] (single GPU only)
    unsigned int* seg_counts;        // ceil(pool/4096)
    unsigned long long* out_idx;     // >= 2048
    unsigned long long* out_spikes;  // >= 2048
    unsigned long long* pack;        // PACK_HDR + 2n u64: {fired, error, kmers, n, 8 time stamps, idx[n], spikes[n]}
    const unsigned long long* kmers; // device k-mer counter of the call
    // multi-GPU "reduce-scatter fused into LIF": this rank owns neurons [slice_lo, slice_lo + lif.pool);
    // its counts are the SUM over all ranks' accumulators, read through NVLink peer mappings.
    // (lif.currents / v / r / spikes are pre-offset by slice_lo; lif.fold_mode must be 2.)
    int npeers;                      // 0: single GPU; else world size (self included)
    const unsigned int* peer_acc[16];
    const unsigned long long* peer_spill[16];  // every rank's u64 spill array (read only for ranks whose spill flag is set)
    unsigned long long slice_lo;
    // peer-signalled mode (nk_dist_run): no host / NCCL barrier around this kernel.
    //   wait_flags != null: before touching any peer's counts every block waits until wait_flags[r] >= epoch
    //   for all r < npeers (rank r's "counting finished" signal, written into THIS rank's mailbox).
    //   peer_mail[r] (set whenever npeers != 0: the spill flags live there); with wait_flags != null block 0 also
    //   delivers the result pack into rank r's mailbox slot `rank` and then raises rank r's "pack delivered"
    //   flag — which also tells r that this rank is done reading r's counts.
    const unsigned long long* wait_flags;
    unsigned long long epoch;
    unsigned long long timeout_ns;
    int rank;
    unsigned char* peer_mail[16];
};

// Mailbox appended to every rank's accumulator allocation (so that the ONE IPC handle of nk_dist_export
// maps it into the peers): flags[0][r] "rank r finished counting", flags[1][r] "rank r's pack delivered"
// (epoch numbers, written by rank r), flags[2][0] "the owner spilled counts into its u64 spill array during
// this job" (written by the owner, read by the peers' slice kernels), then 16 pack slots.  Behind the mailbox
// the allocation carries the owner's SPILL array (pool_size u64): when a stream could overflow the u32
// accumulators (> 2^32-1 window starts on one rank) they are folded into it instead of into `currents`, so
// that the peers — which only ever see this allocation — still read every count.
constexpr int DIST_MAX_WORLD = 16;
constexpr int DIST_FLAG_ROWS = 3;
constexpr unsigned long long PACK_HDR = 12;             // 4 scalars + 8 time stamps before the rows of a result pack
constexpr unsigned long long PACK_MAX_U64 = PACK_HDR + 2 * 2048;
constexpr unsigned long long DIST_PACK_SLOT_U64 = PACK_MAX_U64;
constexpr unsigned long long DIST_MAIL_BYTES =
    ((DIST_FLAG_ROWS * DIST_MAX_WORLD + DIST_MAX_WORLD * DIST_PACK_SLOT_U64) * 8 + 255) / 256 * 256;
inline unsigned long long dist_mail_offset(unsigned long long pool) { return (pool * 4 + 255) / 256 * 256; }
inline unsigned long long dist_spill_offset(unsigned long long pool) { return dist_mail_offset(pool) + DIST_MAIL_BYTES; }
inline unsigned long long dist_alloc_bytes(unsigned long long pool) { return dist_spill_offset(pool) + pool * 8; }
inline unsigned long long* dist_mail_flags(unsigned char* mail, int which) {
    return reinterpret_cast<unsigned long long*>(mail) + which * DIST_MAX_WORLD;
}
inline unsigned long long* dist_mail_slot(unsigned char* mail, int r) {
    return reinterpret_cast<unsigned long long*>(mail) + DIST_FLAG_ROWS * DIST_MAX_WORLD + r * DIST_PACK_SLOT_U64;
}
// raise flags[which][rank] = epoch in every peer's mailbox (one tiny kernel after the count kernels)
cudaError_t launch_dist_signal(unsigned char* const* peer_mail, int world, int rank, int which, unsigned long long epoch,
                               cudaStream_t s);
// wait for every rank's "pack delivered" flag in the local mailbox, then merge the packs found there
cudaError_t launch_merge_mailbox(unsigned char* mail, int world, int rank, unsigned long long n_each, unsigned long long n_out,
                                 unsigned long long epoch, unsigned long long timeout_ns, unsigned long long* pack_out,
                                 cudaStream_t s);
cudaError_t post_max_grid(int device, int* grid);
cudaError_t launch_post(const PostParams& q, int max_grid, cudaStream_t s);
// merge `world` result packs (stride 4+2*n_each u64) into one: fired and k-mers summed, rows re-sorted
cudaError_t launch_merge_packs(const unsigned long long* gathered, int world, int rank, unsigned long long n_each,
                               unsigned long long n_out, unsigned long long* pack_out, cudaStream_t s);

// exact side tables (nk_exact.cu)
struct ExactSlot {            // one record of the table (16 bytes, moved as one uint4)
    unsigned long long key;   // the k-mer word
    unsigned int count;       // occurrences (wraps at 2^32 like the reference's AtomicU32)
    unsigned int idx;         // its neuron index
};
struct BucketPlan {
    unsigned int neurons_per_bucket = 1;   // consecutive neurons that share a bucket ...
    unsigned int splits = 1;               // ... or sub-buckets per neuron (pools with few neurons), by a mix of the word
    unsigned long long nbuckets = 0;
};
struct ExactTable {
    // appended by the count kernel (mode 2 / 4): the word and the neuron index of every window
    unsigned long long* words = nullptr;
    unsigned int* widx = nullptr;
    unsigned long long words_cap = 0, words_bound = 0;
    unsigned long long* cursor = nullptr;  // [0] append cursor, [1] distinct total, [2] overflow flag, [3] windows inside runs of N
                                           // (summed by the count kernel), [4..5] their one weighted record (an ExactSlot)
    // the table: bucket b's distinct records are recs[bucket_start[b] .. + bucket_distinct[b])
    ExactSlot* recs = nullptr;          unsigned long long recs_cap = 0;
    unsigned long long n_keys = 0;      // distinct words
    BucketPlan plan;
    unsigned int* bucket_count = nullptr;        unsigned long long bucket_count_cap = 0;
    unsigned int* bucket_distinct = nullptr;     unsigned long long bucket_distinct_cap = 0;
    unsigned long long* bucket_start = nullptr;  unsigned long long bucket_start_cap = 0;
    unsigned long long* bucket_cursor = nullptr; unsigned long long bucket_cursor_cap = 0;
    unsigned long long* dense_start = nullptr;   unsigned long long dense_start_cap = 0;
    unsigned int* uniques = nullptr;    // per neuron: kmer_per_neuron
    unsigned int* flags = nullptr;
    bool valid = false;
};
cudaError_t exact_reserve_words(ExactTable& t, unsigned long long extra, cudaStream_t s);
cudaError_t exact_clear(ExactTable& t, unsigned long long pool, bool tables_too, cudaStream_t s);
// pool_counts (may be null): per-neuron counts of exactly the appended windows (bucket sizes without a pass over the records)
cudaError_t exact_finalize(ExactTable& t, const FastMod& fm, unsigned long long pool, const unsigned long long* pool_counts,
                           bool merge, cudaStream_t s);
cudaError_t exact_lookup(const ExactTable& t, const FastMod& fm, unsigned long long key, unsigned long long* d_out2, cudaStream_t s);
// the table as dense device arrays in bucket order (any output may be null; n_keys entries each)
cudaError_t exact_dense_copy(ExactTable& t, unsigned long long* out_keys, unsigned int* out_counts, ExactSlot* out_recs,
                             cudaStream_t s);
// multi-GPU groups: the dense table plus the dense ranges of n_ranges neuron ranges; a table from weighted records
cudaError_t exact_dense_export(ExactTable& t, ExactSlot* out_recs, const unsigned long long* bounds, int n_ranges,
                               unsigned long long* first, unsigned long long* last, cudaStream_t s);
cudaError_t exact_build_from_records(ExactTable& t, unsigned long long pool, const ExactSlot* recs, unsigned long long n_recs,
                                     unsigned int flo, unsigned int fhi, cudaStream_t s);
cudaError_t exact_gather_uniques(const ExactTable& t, const unsigned long long* idx, unsigned long long n,
                                 unsigned int* out, cudaStream_t s);
void exact_free(ExactTable& t);
// uniques pass helpers: grow the word array to `cap` entries keeping the first `keep`; set filter bits
cudaError_t exact_grow_words(ExactTable& t, unsigned long long cap, unsigned long long keep, cudaStream_t s);
cudaError_t launch_filter_set(unsigned int* filter, const unsigned long long* idx, unsigned long long n, cudaStream_t s);

// record parsing on the device (nk_parse.cu).  `file`: raw bytes, 16-byte aligned, readable up to size + 16.
// totals (device, 8 u64): [0] bases, [1] records, [2] newlines, [3] kept bytes incl. dropped records, [4] first bad record
size_t parse_scratch_bytes(unsigned long long size);
// FASTA: segments [seg_first, seg_first + nseg) of the raw bytes (units of parse_segment_bytes(); a whole file: 0, all).
// carry (device, 3 u64, zero for a file's first range): what the earlier segments left behind; updated for the next range.
unsigned long long parse_segment_bytes();
cudaError_t launch_fasta_plan(const unsigned char* file, unsigned long long size, unsigned long long seg_first,
                              unsigned long long nseg, void* scratch, unsigned long long* totals, unsigned long long* carry,
                              cudaStream_t s);
cudaError_t launch_fasta_write(const unsigned char* file, unsigned long long size, unsigned long long seg_first,
                               unsigned long long nseg, void* scratch, unsigned char* bases, unsigned long long* offsets,
                               unsigned long long offsets_cap, cudaStream_t s);
cudaError_t launch_fastq_lines(const unsigned char* file, unsigned long long size, void* scratch,
                               unsigned long long* totals, cudaStream_t s);
cudaError_t launch_fastq_write(const unsigned char* file, unsigned long long size, unsigned long long size_real, void* scratch,
                               unsigned char* bases, unsigned long long* offsets, unsigned long long* line_end,
                               unsigned long long nlines, unsigned long long* totals, cudaStream_t s);

// peak calibration (roofline denominators)
cudaError_t launch_int_peak(int mode, unsigned int* out, int blocks, unsigned iters, cudaStream_t s);
unsigned long long int_peak_ops_per_iter(int mode);
// idx[i] = SipHash-1-3(seed + i) % pool: the address stream of real k-mer traffic, for the RED ceiling
cudaError_t launch_hashed_idx(unsigned int* idx, unsigned long long n, unsigned long long seed, FastMod fm, cudaStream_t s);
cudaError_t launch_red_peak(unsigned int* acc, const unsigned int* idx, unsigned long long n, int blocks, cudaStream_t s);

cudaError_t launch_mod_words(const unsigned long long* h, unsigned long long n, FastMod fm, unsigned long long* out,
                             int which, cudaStream_t s);

cudaError_t launch_synth(unsigned char* out, unsigned long long seed, unsigned long long start,
                         unsigned long long n, unsigned flags, cudaStream_t s);

}  // namespace nk
