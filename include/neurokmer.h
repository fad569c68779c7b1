/*
 * neurokmer.h — C ABI of libneurokmer (B200 / sm_100a).
 *
 * This is the drop-in boundary for NeuroKmer's counting hot path.  The
 * reference has no FFI layer of its own: the seam is the Rust type
 * `SpikingKmerCounter` (reference src/spiking_hash.rs:16-37, impl :39-715).
 * Each entry point below names the reference item it replaces; the Rust
 * `extern "C"` block a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - every function returns an nk_status (0 = ok); nk_last_error() returns a
 *    thread-local human-readable message for the last failure on this thread;
 *  - the caller owns every input and output buffer, the library owns the
 *    opaque handle and all device / pinned memory behind it;
 *  - one thread at a time per handle (the reference takes `&mut self`);
 *    distinct handles are independent;
 *  - a batch of sequences is passed as one concatenated byte array `bases`
 *    plus `offsets[nseq+1]` (offsets[0] = 0, sequence i = bases[offsets[i] ..
 *    offsets[i+1])), i.e. the flattened form of the reference's `&[Vec<u8>]`;
 *  - there is NO CPU fallback: without a usable sm_100 device nk_create fails.
 */
#ifndef NEUROKMER_H
#define NEUROKMER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NK_API __attribute__((visibility("default")))
#else
#define NK_API
#endif

typedef enum nk_status {
    NK_OK = 0,
    NK_ERR_BAD_ARG = 1,     /* k outside [1,32], pool_size == 0, null pointer, bad offsets … */
    NK_ERR_IO = 2,          /* file could not be opened / read (reference: Err from parse_fastx_file) */
    NK_ERR_CUDA = 3,        /* CUDA runtime error; message carries cudaGetErrorString */
    NK_ERR_NO_DEVICE = 4,   /* no CUDA device, or device is not compute capability 10.x */
    NK_ERR_OOM = 5,         /* device or pinned-host allocation failed */
    NK_ERR_STATE = 6,       /* call sequence error (e.g. stream_push without stream_begin) */
    NK_ERR_UNSUPPORTED = 7  /* documented limit exceeded (e.g. pool_size >= 2^32) */
} nk_status;

typedef struct nk_counter nk_counter; /* opaque; replaces `SpikingKmerCounter` */

/* Arguments of SpikingKmerCounter::new (reference src/spiking_hash.rs:40-48)
 * plus `steps` (field, default 1000, :70) and the CUDA device ordinal. */
typedef struct nk_config {
    uint32_t k;            /* 1..32 */
    uint32_t refractory;   /* refractory period in ticks */
    uint64_t pool_size;    /* 1 .. 2^32-1 neurons */
    uint64_t steps;        /* LIF ticks per simulation */
    double   spike_cost;
    float    threshold;
    float    leak;
    int32_t  use_canonical; /* 0 = pack_kmer path, 1 = rolling canonical path */
    int32_t  device;        /* CUDA device ordinal */
} nk_config;

/* One row of top_abundant_neurons()'s Vec<(usize, u64, u32)>
 * (reference src/spiking_hash.rs:661-673). `uniques` needs the exact-count side
 * table: without nk_enable_exact_counts(h, 1) it is reported as
 * NK_UNIQUES_NOT_COMPUTED, never faked. */
typedef struct nk_top_entry {
    uint64_t idx;
    uint64_t spikes;
    uint32_t uniques;
    uint32_t _pad;
} nk_top_entry;
#define NK_UNIQUES_NOT_COMPUTED 0xFFFFFFFFu

/* Device-side time of the last process/simulate call, CUDA events, milliseconds. */
typedef struct nk_timings {
    float h2d_ms;        /* host→device copies (0 for staged/device-resident input) */
    float mark_ms;       /* invalid-start bitmap: memset + sequence-end marking kernels */
    float count_ms;      /* the fused windowing + SipHash + mod + pool-update kernel(s) */
    float fold_ms;       /* u32 batch pool → u64 currents */
    float lif_ms;        /* LIF simulation */
    float topn_ms;       /* last nk_top_n */
    float total_ms;      /* first kernel/copy → last kernel of the call */
    uint64_t kmers;      /* windows counted by the call */
    uint64_t launches;   /* kernels launched by the call */
    uint64_t h2d_bytes;  /* bytes copied host→device by the call */
    uint64_t d2h_bytes;  /* bytes copied device→host by the call */
    int32_t  lif_path;   /* 0 none, 1 direct simulation, 2 per-count table (uniform fresh state), 3 table fused with
                          * top-N, 4 sharded-pool slice kernel, 5 memoised (one simulation per distinct state+count) */
    int32_t  _pad;
    uint64_t topn_launches; /* kernels launched by the last nk_top_n */
    /* The fused fold + LIF + top-N kernel and, for sharded-pool multi-GPU jobs, the EXCHANGE reported on its
     * own (north star: "the allreduce cost reported separately").  Measured by the kernels themselves with
     * %globaltimer on this GPU; all 0 when the fused kernel did not run. */
    float post_ms;          /* fused kernel: start -> result pack written */
    float exch_wait_ms;     /* waiting for the peers: their "finished counting" flags (slice kernel) + their result
                             * packs (merge kernel) — rank skew, not bytes */
    float exch_reduce_ms;   /* the phase that reads this rank's neuron slice of EVERY rank's counts over NVLink peer
                             * memory (reduce-scatter) fused with the LIF look-up and the top-N histogram; on a single
                             * GPU the same phase over local memory (the baseline the exchange is compared with) */
    float merge_ms;         /* merging the ranks' result packs */
    uint64_t exch_bytes;    /* bytes this rank read from peer memory for the job */
} nk_timings;

/* ---- lifecycle ---------------------------------------------------------- */
/* CLI defaults of the reference: k=31, pool 1,000,000, canonical off,
 * threshold 1.0, leak 0.95, refractory 2, spike_cost 1.0 (src/main.rs:10-37),
 * steps 1000 (src/spiking_hash.rs:70). */
NK_API int nk_config_default(nk_config* cfg);
/* SpikingKmerCounter::new — src/spiking_hash.rs:40-77 */
NK_API int nk_create(const nk_config* cfg, nk_counter** out);
NK_API int nk_destroy(nk_counter* h);
/* Back to the state nk_create leaves (all neurons v=0, r=0, spike_count=0, energy 0,
 * currents 0): what constructing a fresh SpikingKmerCounter gives (:49-76). */
NK_API int nk_reset(nk_counter* h);
NK_API const char* nk_last_error(void);
NK_API const char* nk_version(void);

/* ---- one input sharded over several GPUs of THIS process --------------------------------------------
 * The reference spreads one input over the cores of one process (rayon fold/reduce over per-thread current
 * vectors, src/spiking_hash.rs:94-154; worker threads, :292-403).  nk_create_multi is the same thing with
 * GPUs: the returned handle is used exactly like one from nk_create — nk_process_batch, nk_stream_*,
 * nk_process_file, the packed variants, nk_top_n, totals, nk_copy_*, nk_uniques_*, the exact side tables
 * (nk_enable_exact_counts ... nk_copy_uniques) — and every batch is cut
 * by window start into one contiguous range per device (sequences are cut wherever a range ends; the owner of
 * starts [a, b) reads bases [a, min(b + k-1, end of sequence)), so every window is counted exactly once).
 * Each device counts into its own accumulators; the exchange is a reduce-scatter fused into the LIF/top-N
 * kernel over NVLink peer memory (device r owns neurons [r*P/N, (r+1)*P/N)), ordered by CUDA events.
 * Results are bit-identical to a single-GPU counter.  `devices` = NULL means ordinals 0..n_devices-1;
 * cfg->device is ignored.  Needs peer access between the devices (NK_ERR_UNSUPPORTED otherwise).
 * Not available on a group handle (NK_ERR_UNSUPPORTED): nk_process_sequence (sequential by definition:
 * replicas only), device-resident staging, the nk_dist_* plumbing. */
NK_API int nk_device_count(int32_t* n);
NK_API int nk_create_multi(const nk_config* cfg, const int32_t* devices, int32_t n_devices, nk_counter** out);
/* number of devices behind the handle (1 for nk_create) */
NK_API int nk_group_size(const nk_counter* h, int32_t* n);
/* Host-only view of the shard plan (needs no device): member `rank` of `world` reads the batch from base
 * *start on and counts the pieces piece_offsets[0 .. *n_pieces] (relative to *start; capacity nseq+1). */
NK_API int nk_debug_shard(const uint64_t* offsets, uint64_t nseq, uint32_t k, int32_t world, int32_t rank,
                          uint64_t* start, uint64_t* piece_offsets, uint64_t* n_pieces);

/* set_steps / get_steps — src/spiking_hash.rs:688-695 */
NK_API int nk_set_steps(nk_counter* h, uint64_t steps);
NK_API int nk_get_steps(const nk_counter* h, uint64_t* steps);

/* ---- batch entry points --------------------------------------------------- */
/* process_parallel — src/spiking_hash.rs:84-201.  Counts every window of every
 * sequence, OVERWRITES the per-neuron currents with this call's totals, then
 * runs the in-memory LIF driver (zero-current neurons skipped, :187-200).
 * `bases`/`offsets` are host pointers (pinned memory is copied directly,
 * pageable memory goes through the library's pinned staging ring). */
NK_API int nk_process_batch(nk_counter* h, const uint8_t* bases, const uint64_t* offsets, uint64_t nseq);

/* process_file_streaming minus parsing — src/spiking_hash.rs:277-486.
 * begin: zero the accumulators; push: count a batch; end: totals OVERWRITE currents, then the
 * streaming LIF driver runs (every neuron stepped, :544-659).
 * When does push return?  Always once the caller's buffers may be reused, and no later:
 *  - pageable memory: the batch is staged through the library's pinned ring in 32 MiB chunks; push returns
 *    when the last chunk has been copied, with that chunk's kernels still running (the next push's copies
 *    overlap them).  Batches of 8 MiB and more are staged by a pool of host threads; where that pool has ten or
 *    more workers (a host with cores to spare) the threads PACK their pieces to 2 bits per base + `other` bits
 *    on the way, so that 3/8 of the bytes cross the link (NK_STAGE_PACK=0 / 1 forces plain copies / packing);
 *  - on a single-GPU host with such a pool, pinned batches of 8 MiB and more go the same way (packing them on the
 *    way beats reading them in place across PCIe: 1.9-2.1 against 2.4 ms for 113 MB);
 *  - otherwise pinned, device-mapped memory (nk_host_alloc, cudaHostAlloc, cudaHostRegister) at a 16-byte aligned
 *    address: the count kernel reads the batch IN PLACE across PCIe (no staging copy), so push returns when
 *    that kernel has finished — synchronous, but the copy it replaces would have taken as long.  The kernel's
 *    16-byte bulk reads may touch up to 15 bytes past offsets[nseq] (never past the end of the page that
 *    holds the last base).  NK_ZEROCOPY=0 in the environment forces the staged path. */
NK_API int nk_stream_begin(nk_counter* h);
NK_API int nk_stream_push(nk_counter* h, const uint8_t* bases, const uint64_t* offsets, uint64_t nseq);
NK_API int nk_stream_end(nk_counter* h);

/* ---- pre-packed input (2 bits per base) -------------------------------------
 * The north star's "ASCII or pre-packed bases": callers that already hold 2-bit data (or pack it
 * once with nk_pack_bases and count it many times) move 3/8 of the bytes over PCIe and skip the
 * kernel's byte classifier.  Results are bit-identical to the ASCII entry points on the bytes the
 * packed form was made from (tests/test_parity_gpu.py::test_packed_*).  Layout ("nk2"), for the
 * concatenated batch of nbases = offsets[nseq] bases (sequences are NOT word-aligned):
 *   codes[i], i < nk_packed_code_words(nbases):  bases 16i .. 16i+15, base 16i in bits 31:30;
 *       A,a=0 C,c=1 G,g=2 T,t=3 (base_to_bits, src/models.rs:231-239), anything else 0;
 *   other[w], w < nk_packed_other_words(nbases): bit (p & 31) of word p >> 5 set iff byte p was not
 *       one of ACGTacgt: code 0 on BOTH strands in canonical mode (src/models.rs:237,249), skipped by
 *       pack_kmer (src/utils.rs:35).  NULL = no such byte in the batch (no array is read or copied).
 * Unused bits of the last words must be present (readable) but are ignored. */
NK_API uint64_t nk_packed_code_words(uint64_t nbases);
NK_API uint64_t nk_packed_other_words(uint64_t nbases);
/* Host-side packer (SIMD, `threads` host threads; <= 0: all).  *n_other = bytes that were not ACGTacgt
 * (0 => `other` may be passed as NULL below).  `other` may be NULL if the caller does not want it. */
NK_API int nk_pack_bases(const uint8_t* bases, uint64_t nbases, uint32_t* codes, uint32_t* other, int threads,
                         uint64_t* n_other);
/* process_parallel (src/spiking_hash.rs:84-201) / the push of process_file_streaming (:277-486) on
 * pre-packed input; same semantics, state rules and asynchrony as nk_process_batch / nk_stream_push. */
NK_API int nk_process_batch_packed(nk_counter* h, const uint32_t* codes, const uint32_t* other,
                                   const uint64_t* offsets, uint64_t nseq);
NK_API int nk_stream_push_packed(nk_counter* h, const uint32_t* codes, const uint32_t* other,
                                 const uint64_t* offsets, uint64_t nseq);

/* Whole-file drivers: main.rs:170-177.  streaming != 0 → process_file_streaming
 * (:277), else stream_sequences().collect() + process_parallel (main.rs:175-176).
 * FASTA/FASTQ record rules follow src/utils.rs:9-24.
 * Plain regular files that fit the device are copied there raw by a pool of host threads (pread -> pinned
 * slots -> async H2D) and split into records ON THE DEVICE (no host pass over the bytes); compressed input,
 * pipes and larger files go through the host reader (NK_GPU_PARSE=0 forces it).  Same results either way.
 * Limit of the host reader only: a FASTQ record whose sequence is longer than its 32 MiB batch buffer fails with
 * NK_ERR_UNSUPPORTED (FASTA records of any length are cut with a k-1 overlap; the device parser has no such limit). */
NK_API int nk_process_file(nk_counter* h, const char* path, int streaming);

/* process_sequence — src/spiking_hash.rs:203-273 (per-sequence API: one LIF tick
 * per call with the raw count as input current; currents zeroed afterwards). */
NK_API int nk_process_sequence(nk_counter* h, const uint8_t* seq, uint64_t len);

/* simulate_spikes_auto — src/spiking_hash.rs:697-714 (streaming-driver semantics
 * on the currently stored currents). */
NK_API int nk_simulate(nk_counter* h);

/* ---- read-outs ------------------------------------------------------------ */
/* top_abundant_neurons — src/spiking_hash.rs:661-673: spike_count descending,
 * ties by ascending neuron index; writes min(top_n, pool_size) rows. */
NK_API int nk_top_n(nk_counter* h, uint64_t top_n, nk_top_entry* out, uint64_t* n_out);
/* energy.total_spikes() — src/models.rs:166-168 */
NK_API int nk_total_spikes(const nk_counter* h, uint64_t* out);
/* energy_used() — src/spiking_hash.rs:684-686, src/models.rs:170-172 */
NK_API int nk_energy_used(const nk_counter* h, double* out);
/* Exact side tables — opt-in because they cost O(windows) device memory:
 *   counts: DashMap<u64, AtomicU32>          src/spiking_hash.rs:27 (filled :110,124,157-165,333,343,442-447)
 *   kmer_per_neuron: DashMap<usize, u32>     :26, :167-172, :467-473 (the `uniques` column of nk_top_n)
 * Enable BEFORE processing.  A batch/stream/file call then replaces both tables with this call's
 * (counts.clear(), :157/:426); nk_process_sequence adds to them (:218-221, :262-264).  Limits:
 * 28 B of device memory per window of a call while its table is built.  On a multi-GPU group (nk_create_multi) every
 * GPU builds the table of its own windows and GPU d merges the records of its neuron slice from all of them: ONE table
 * over the whole input, like the reference's map; handles of different PROCESSES (nk_dist_*) keep per-rank tables.
 * Off by default: then
 * nk_get_count returns NK_ERR_UNSUPPORTED and nk_top_n reports NK_UNIQUES_NOT_COMPUTED. */
NK_API int nk_enable_exact_counts(nk_counter* h, int on);
/* get_count — src/spiking_hash.rs:675-678: *found = 0 where the reference returns None. */
NK_API int nk_get_count(nk_counter* h, uint64_t kmer, uint32_t* count, int32_t* found);
/* the whole `counts` table (python.rs:31-40 walks it).  Like the reference's DashMap it has no defined order: the
 * records come out grouped by neuron range (that is how the table is built: one partition by neuron index, then a
 * shared-memory hash table per bucket — no sort anywhere). */
NK_API int nk_exact_table_size(nk_counter* h, uint64_t* n);
NK_API int nk_copy_exact_table(nk_counter* h, uint64_t* keys /* n */, uint32_t* counts /* n */);
/* kmer_per_neuron for every neuron (0 where the reference's map has no entry) */
NK_API int nk_copy_uniques(nk_counter* h, uint32_t* out /* pool_size */);

/* The `uniques` column for the TOP ROWS ONLY, without the exact table: a second pass over the same input
 * that keeps just the windows mapped to those rows' neurons (memory O(matches), not O(windows); this is
 * all the reference's CLI prints, src/main.rs:54-61).
 *   nk_uniques_begin(h, top_n)   fixes the top_n (<= 2048) rows of the current state and arms the filter;
 *   nk_uniques_push[_packed]     re-supply the batches that were counted, in any batching (synchronous);
 *   nk_uniques_end(h)            afterwards nk_top_n(h, n <= top_n, ...) reports `uniques` for its rows,
 *                                until the next job changes the state.
 * nk_set_file_uniques(h, top_n): nk_process_file does this itself by reading the file a second time
 * (top_n = 0 turns it off again; ignored while the exact table is enabled). */
NK_API int nk_uniques_begin(nk_counter* h, uint64_t top_n);
NK_API int nk_uniques_push(nk_counter* h, const uint8_t* bases, const uint64_t* offsets, uint64_t nseq);
NK_API int nk_uniques_push_packed(nk_counter* h, const uint32_t* codes, const uint32_t* other, const uint64_t* offsets,
                                  uint64_t nseq);
NK_API int nk_uniques_end(nk_counter* h);
NK_API int nk_set_file_uniques(nk_counter* h, uint64_t top_n);

/* ---- parity taps (debug; not on the timed path) ---------------------------- */
/* Words and neuron indices of every window of one sequence, in order: the
 * values `packed`/`idx` take at src/spiking_hash.rs:108-111,121-125 (canonical)
 * or :132-135 (pack_kmer).  Any of fwd/rc/words/idx may be NULL.  *n_out =
 * max(0, len-k+1); buffers must hold that many u64. fwd/rc are only defined
 * in canonical mode (RollingKmerHash::forward/reverse_complement). */
NK_API int nk_debug_kmers(nk_counter* h, const uint8_t* seq, uint64_t len, uint64_t* fwd, uint64_t* rc,
                          uint64_t* words, uint64_t* idx, uint64_t* n_out);
/* the same taps for one sequence in pre-packed form */
NK_API int nk_debug_kmers_packed(nk_counter* h, const uint32_t* codes, const uint32_t* other, uint64_t len,
                                 uint64_t* fwd, uint64_t* rc, uint64_t* words, uint64_t* idx, uint64_t* n_out);
/* nk_pack_bases with one named body on one thread: 1 portable, 2 AVX2, 3 AVX-512BW
 * (NK_ERR_UNSUPPORTED if this CPU lacks it) — the CPU tests compare every body with a numpy twin. */
NK_API int nk_debug_pack_body(const uint8_t* bases, uint64_t nbases, uint32_t* codes, uint32_t* other, int body,
                              uint64_t* n_other);
/* Host-only check of the record reader behind nk_process_file (src/utils.rs:9-24 rules; plain, gzip,
 * bzip2, xz, zstd input): number of records, total sequence bytes and FNV-1a-64 over every record's
 * sequence bytes followed by one 0xFF byte.  Needs no device. */
NK_API int nk_debug_fastx_digest(const char* path, uint64_t* nrecords, uint64_t* nbases, uint64_t* fnv1a);
/* The same digest of what the DEVICE-side record parser (nk_parse.cu: the path nk_process_file takes for plain
 * regular FASTA / FASTQ files) yields; NK_ERR_UNSUPPORTED where that path does not apply (compressed input, ...). */
NK_API int nk_debug_parse_file(nk_counter* h, const char* path, uint64_t* nrecords, uint64_t* nbases, uint64_t* fnv1a);
/* Measurement tap for the file path: milliseconds the staging pool needs to pread the whole file into its pinned
 * slots with no device work (host_only_ms), and with the H2D copies into device memory, completed (with_h2d_ms). */
NK_API int nk_debug_stage_file(nk_counter* h, const char* path, double* host_only_ms, double* with_h2d_ms);
/* the same digest (plain FASTA only) through the parallel ingest's window planner + window parser, with windows
 * of `window` bytes (>= 64), run serially: checks the code the multi-threaded file path is made of */
NK_API int nk_debug_fasta_windows_digest(const char* path, uint64_t window, uint64_t* nrecords, uint64_t* nbases,
                                         uint64_t* fnv1a);
/* SipHash-1-3(keys 0,0) of LE64(word) and word-hash % pool_size for a host array. */
NK_API int nk_debug_hash(nk_counter* h, const uint64_t* words, uint64_t n, uint64_t* hashes, uint64_t* idx);
/* values[i] % pool_size on the device for ANY pool_size in [1, 2^32) without allocating a pool:
 * which 0 = the two-stage FP64-pipe routine (any pool size), 1 = the integer (Moeller-Granlund) routine,
 * 2 = the routine the count kernel picks for this pool size (one FP64 stage where the size allows it). */
NK_API int nk_debug_mod(const uint64_t* values, uint64_t n, uint64_t pool_size, int which, uint64_t* out);
NK_API int nk_copy_currents(nk_counter* h, uint64_t* out /* pool_size */);
NK_API int nk_copy_spike_counts(nk_counter* h, uint64_t* out /* pool_size */);
NK_API int nk_copy_voltages(nk_counter* h, float* out /* pool_size */);
NK_API int nk_copy_refractory(nk_counter* h, uint32_t* out /* pool_size */);
NK_API int nk_last_timings(const nk_counter* h, nk_timings* out);
/* LIF path selection for tests: 0 = automatic (per-count table while every neuron is still in its
 * initial state; afterwards one simulation per distinct (state, count) key — "memoised" — where the
 * parameters allow it, else one per neuron), 1 = always one simulation per neuron (direct),
 * 2 = never the per-count table (memoised even from the initial state). */
NK_API int nk_debug_set_lif_path(nk_counter* h, int mode);
/* The u32 batch accumulators are folded away (into the u64 currents; on a sharded-pool handle into the u64 spill
 * array its peers read) before more than `limit` window starts could have been added to them.  Default and
 * maximum 2^32-1; tests lower it to exercise the guard without 4.3 Gbase of input. */
NK_API int nk_debug_set_fold_limit(nk_counter* h, uint64_t limit);

/* Roofline denominators measured on this device (micro-kernels, CUDA events, best of 3):
 *   which 0: 32-bit ALU-pipe ops/s of independent LOP3+SHF chains (xor/rotate: what SipHash
 *            cannot move off the ALU pipe);
 *   which 1: 32-bit integer ops/s of register-only SipRound chains (SipHash's own
 *            add:xor:rotate mix, 24 ops per round) — the integer-pipe ceiling of step 2;
 *   which 2: RED.ADD.U32 per second at the addresses real k-mer traffic produces (SipHash-1-3 of
 *            consecutive words % pool_size, precomputed) into THIS handle's pool
 *            (the pool-update ceiling of step 3).  Needs an idle counter; leaves it unchanged. */
NK_API int nk_calibrate(nk_counter* h, int which, double* out);

/* ---- device-resident input and multi-GPU plumbing --------------------------- */
/* Reserve library-owned DEVICE buffers for a batch the caller will fill itself
 * (cudaMemcpyAsync, its own kernel, or nk_synth_fill): *dev_bases holds nbytes
 * (+ padding the kernels may read), *dev_offsets holds nseq+1 u64. */
NK_API int nk_stage_reserve(nk_counter* h, uint64_t nbytes, uint64_t nseq, void** dev_bases, void** dev_offsets);
/* Count the staged batch.  mode 0: process_parallel semantics (count, overwrite
 * currents, in-memory LIF); mode 1: accumulate only (inside stream_begin/end). */
NK_API int nk_process_staged(nk_counter* h, uint64_t nbytes, uint64_t nseq, int mode);
/* The same for pre-packed input: *dev_codes holds nk_packed_code_words(nbases) u32 (+ padding),
 * *dev_other nk_packed_other_words(nbases) u32 (+ padding); has_other = 0 ignores the `other` array. */
NK_API int nk_stage_reserve_packed(nk_counter* h, uint64_t nbases, uint64_t nseq, void** dev_codes, void** dev_other,
                                   void** dev_offsets);
NK_API int nk_process_staged_packed(nk_counter* h, uint64_t nbases, uint64_t nseq, int mode, int has_other);
/* Sharded (one process per GPU) runs: every rank accumulates its shard between
 * nk_stream_begin and nk_stream_accumulated, sum-reduces the u64 array at
 * *dev_currents (pool_size elements) across ranks with its own collective
 * (e.g. ncclAllReduce ncclUint64 / torch.distributed), then calls
 * nk_stream_finish.  nk_stream_end == accumulated + finish. */
NK_API int nk_stream_accumulated(nk_counter* h, void** dev_currents);
NK_API int nk_stream_finish(nk_counter* h);
/* Sharded-pool multi-GPU mode: the reduce-scatter of the per-rank counts is fused into the LIF
 * kernel through NVLink peer mappings of every rank's accumulators (no 16 MB all-reduce).
 *   setup (once):  nk_dist_export on every rank -> exchange the 64-byte IPC handles ->
 *                  nk_dist_setup(rank, world, handles, NULL)   [same process: raw pointers instead]
 *   per job:       nk_reset; nk_stream_begin; push / process_staged(mode 1);
 *                  <cross-rank barrier on nk_cuda_stream()>;
 *                  nk_dist_post -> all-gather the returned packs on the stream -> nk_dist_complete
 * Afterwards nk_total_spikes / nk_energy_used / nk_top_n answer for the WHOLE pool on every rank;
 * nk_copy_* return data that is only defined inside this rank's slice (nk_dist_slice).
 * Fresh-state jobs only (one job per nk_reset); otherwise nk_dist_post returns NK_ERR_UNSUPPORTED
 * and the all-reduce flow above applies. */
NK_API int nk_dist_export(nk_counter* h, void* ipc_handle_64_bytes, void** raw_acc);
NK_API int nk_dist_setup(nk_counter* h, int rank, int world, const void* ipc_handles /* world*64 B */,
                         void* const* raw_ptrs /* world pointers, same process */);
NK_API int nk_dist_post(nk_counter* h, void** dev_pack, uint64_t* pack_u64s, uint64_t* n_each);
NK_API int nk_dist_complete(nk_counter* h, const void* dev_gathered_packs, uint64_t n_each);
NK_API int nk_dist_slice(const nk_counter* h, uint64_t* lo, uint64_t* len);
/* The same exchange with NO host or NCCL synchronisation per job (one handle per GPU):
 *   per job:  nk_reset; nk_stream_begin; push / process_staged(mode 1); nk_dist_run
 * nk_dist_run signals "this rank finished counting" into every peer's memory (NVLink stores), its slice
 * kernel waits for all ranks' signals before reading their counts, delivers its result pack into every
 * rank's mailbox, and a one-block kernel waits for all packs and merges them — nothing but kernels on
 * the handle's stream.  nk_dist_setup is a collective call: a cross-rank barrier must separate it from the
 * first nk_dist_run.  A rank that never arrives surfaces as NK_ERR_STATE from the first call that observes
 * the result, after NK_DIST_TIMEOUT_MS (environment, default 30000). */
NK_API int nk_dist_run(nk_counter* h);
/* The CUDA stream (cudaStream_t) the handle's kernels run on. */
NK_API int nk_cuda_stream(nk_counter* h, void** stream);
NK_API int nk_synchronize(nk_counter* h);

/* Position-addressable synthetic base generator: writes bases
 * [start, start+n) of stream `seed` to device memory `dev_out`.
 * flags bit0: sparse N-runs, bit1: 1 % lower-case blocks. Runs on the handle's stream. */
NK_API int nk_synth_fill(nk_counter* h, void* dev_out, uint64_t seed, uint64_t start, uint64_t n, uint32_t flags);

/* Pinned host memory helpers (so callers can parse straight into DMA-able memory). */
NK_API int nk_host_alloc(void** ptr, uint64_t nbytes);
NK_API int nk_host_free(void* ptr);

/* non-canonical helper: pack_kmer — src/utils.rs:26-39 (pure host arithmetic,
 * exported because python.rs:50-52 exposes it as pack_kmer_py). */
NK_API uint64_t nk_pack_kmer(const uint8_t* kmer, uint64_t len);

#ifdef __cplusplus
}
#endif
#endif /* NEUROKMER_H */
