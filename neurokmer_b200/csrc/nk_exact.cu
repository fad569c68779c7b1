// nk_exact.cu — the exact side tables of the reference counter, opt-in.  Hand-written
// kernels only (round 1 called cub::DeviceRadixSort / RunLengthEncode / ReduceByKey here).
//
// Replaces, when nk_enable_exact_counts(h, 1) was called:
//   counts: DashMap<u64, AtomicU32>             src/spiking_hash.rs:27, filled :110,124,157-165,333,343,442-447
//   get_count                                   src/spiking_hash.rs:675-678
//   kmer_per_neuron: DashMap<usize, u32>        :26, rebuilt :167-172 / :467-473 (= #distinct k-mers per neuron,
//                                               the `uniques` column of top_abundant_neurons :667)
//   process_sequence's variants                 :218-221, 259-263 (# sequences that touched the neuron)
//
// The reference's table is a hash map (DashMap): unordered, keyed by the k-mer word.  Nothing in it needs a
// global sort — what is needed is "equal words meet" and "words of one neuron meet".  Both follow from ONE
// partition by neuron index, which the count kernel has already computed for every window:
//   0. the count kernel appends (word, neuron index) of every window               [nk_count.cu, mode 2 / 4]
//   1. sizes of the BUCKETS of consecutive neurons (~1.5 K windows per bucket): for a whole-call table they are sums
//      of the per-neuron counts the call has produced anyway (no pass over the records); otherwise one RED per
//      record.  Pools with few neurons split every neuron into sub-buckets by a mix of the word.
//   2. one-block exclusive scan of the bucket sizes -> one 64-bit write cursor per bucket
//   3. scatter: per record ONE atomic on its bucket's cursor and ONE 16-byte store {word, weight, index}
//   4. one CTA per bucket: a shared-memory hash table (64-bit atomicCAS) merges equal words and counts them,
//      every NEW word adds one to its neuron's `uniques`; the distinct records are written back IN PLACE at the
//      head of the bucket's own segment
// get_count(word) = hash -> neuron -> bucket -> one warp scans that bucket's distinct records.
// The table is therefore grouped by neuron range and unordered inside a bucket (like the reference's map);
// nk_copy_exact_table compacts the buckets into dense arrays.  No 2^31 limit: every cursor is 64-bit.
// Memory: 12 B per window appended + 16 B per window partitioned.  Counts wrap at 2^32 like the reference's
// AtomicU32::fetch_add.
//
// Measured on the bench workload (113 M windows, 112.4 M distinct, B200; profiles/r02_exact.md): round 1's CUB
// pipeline 12.9 ms per job.  Two designs were tried before this one in round 2: the same partition with three
// separate arrays and a separate start[] + fill[] per bucket (five L2 transactions per record: scatter 7.0 ms),
// and ONE open-addressing table in HBM with a 64-bit CAS per record (no partition at all: 16.5 ms for the insert
// kernel — 113 M random read-modify-writes of DRAM sectors run at 6.8 G/s).
#include <algorithm>
#include <atomic>

#include "nk_kernels.cuh"

namespace nk {

namespace {

constexpr int XT = 256;                       // threads per block
constexpr unsigned TABLE_SLOTS = 2048;        // shared-memory hash table of one bucket (power of two)
constexpr unsigned long long BUCKET_TARGET = 768;   // windows per bucket aimed at (table load <= ~0.4 when all distinct)
constexpr unsigned long long EMPTY = ~0ull;   // no k < 32 word and no canonical word equals it; see dedup kernel

#define NKX(expr)                        \
    do {                                 \
        cudaError_t e_ = (expr);         \
        if (e_ != cudaSuccess) return e_; \
    } while (0)

cudaError_t ensure(void** p, unsigned long long* cap, unsigned long long bytes) {
    if (bytes <= *cap) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    NKX(cudaMalloc(p, bytes));
    *cap = bytes;
    return cudaSuccess;
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {  // splitmix64 finaliser
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

__device__ __forceinline__ unsigned long long bucket_of(unsigned idx, unsigned long long word, const BucketPlan& bp) {
    const unsigned long long g = idx / bp.neurons_per_bucket;
    return bp.splits > 1 ? g * bp.splits + (mix64(word) >> 32) % bp.splits : g;
}

// bucket sizes without touching the records: a whole-call table holds exactly the windows the call counted, so the
// size of a bucket of consecutive neurons is the sum of their counts (`pool_counts` = the call's u64 currents)
__global__ void bucket_hist_from_pool_kernel(const unsigned long long* __restrict__ pool_counts, unsigned long long pool,
                                             BucketPlan bp, unsigned int* __restrict__ bucket_count) {
    const unsigned long long b = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (b >= bp.nbuckets) return;
    const unsigned long long lo = b * bp.neurons_per_bucket;
    const unsigned long long hi = lo + bp.neurons_per_bucket < pool ? lo + bp.neurons_per_bucket : pool;
    unsigned long long s = 0;
    for (unsigned long long i = lo; i < hi; ++i) s += pool_counts[i];
    bucket_count[b] = (unsigned int)s;
}

// (flo, fhi): only records of the neurons [flo, fhi) take part (a slice table built from other GPUs' whole buckets)
__global__ void bucket_hist_kernel(const unsigned long long* __restrict__ words, const unsigned int* __restrict__ widx,
                                   const ExactSlot* __restrict__ recs, unsigned long long n_recs, unsigned long long n, BucketPlan bp,
                                   unsigned int flo, unsigned int fhi, unsigned int* __restrict__ bucket_count) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n + n_recs; i += stride) {
        unsigned ix;
        unsigned long long w = 0;
        if (i < n_recs) { ix = recs[i].idx; if (bp.splits > 1) w = recs[i].key; }
        else { ix = widx[i - n_recs]; if (bp.splits > 1) w = words[i - n_recs]; }
        if (ix < flo || ix >= fhi) continue;
        atomicAdd(bucket_count + bucket_of(ix, w, bp), 1u);
    }
}

// one block: start[b] = sum of count[< b], start[nb] = total, cursor[b] = start[b] (the scatter's write cursors)
__global__ void __launch_bounds__(1024) bucket_scan_kernel(const unsigned int* __restrict__ count, unsigned long long nb,
                                                            unsigned long long* __restrict__ start, unsigned long long* __restrict__ cursor) {
    __shared__ unsigned long long s_sum[1024];
    const unsigned t = threadIdx.x;
    const unsigned long long per = (nb + 1023) / 1024;
    const unsigned long long a = (unsigned long long)t * per, b = a + per < nb ? a + per : nb;
    unsigned long long s = 0;
    for (unsigned long long i = a; i < b; ++i) s += count[i];
    s_sum[t] = s;
    __syncthreads();
    if (t == 0) {
        unsigned long long run = 0;
        for (int q = 0; q < 1024; ++q) { const unsigned long long x = s_sum[q]; s_sum[q] = run; run += x; }
        start[nb] = run;
    }
    __syncthreads();
    s = s_sum[t];
    for (unsigned long long i = a; i < b; ++i) {
        start[i] = s;
        if (cursor) cursor[i] = s;
        s += count[i];
    }
}

// per record: one atomic on the bucket's cursor, one 16-byte store.  The kernel is latency-bound (ncu: 68 % of the
// stall samples wait for the atomic's return value, 4.6 % of the issue slots used), so every thread keeps
// SCATTER_BATCH records in flight: all loads, then all atomics, then all stores.
constexpr int SCATTER_BATCH = 4;
__global__ void bucket_scatter_kernel(const unsigned long long* __restrict__ words, const unsigned int* __restrict__ widx,
                                      const ExactSlot* __restrict__ recs, unsigned long long n_recs, unsigned long long n,
                                      BucketPlan bp, unsigned int flo, unsigned int fhi,
                                      unsigned long long* __restrict__ cursor, ExactSlot* __restrict__ out) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x, total = n + n_recs;
    for (unsigned long long i0 = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i0 < total; i0 += stride * SCATTER_BATCH) {
        ExactSlot r[SCATTER_BATCH];
        bool take[SCATTER_BATCH];
#pragma unroll
        for (int j = 0; j < SCATTER_BATCH; ++j) {
            const unsigned long long i = i0 + (unsigned long long)j * stride;
            take[j] = i < total;
            r[j].key = 0ull; r[j].idx = 0u; r[j].count = 1u;
            if (take[j]) {
                if (i < n_recs) r[j] = recs[i];
                else { r[j].key = words[i - n_recs]; r[j].idx = widx[i - n_recs]; }
                take[j] = r[j].idx >= flo && r[j].idx < fhi;
            }
        }
        unsigned long long pos[SCATTER_BATCH];
#pragma unroll
        for (int j = 0; j < SCATTER_BATCH; ++j)
            pos[j] = take[j] ? atomicAdd(cursor + bucket_of(r[j].idx, r[j].key, bp), 1ull) : 0ull;
#pragma unroll
        for (int j = 0; j < SCATTER_BATCH; ++j)
            if (take[j]) reinterpret_cast<uint4*>(out)[pos[j]] = *reinterpret_cast<const uint4*>(&r[j]);
    }
}

// One CTA per bucket.  Equal words are merged in a shared-memory hash table (linear probing, 64-bit atomicCAS
// on the key).  The all-ones word (only pack_kmer of 32 T's, k = 32, non-canonical) is the table's EMPTY marker
// and therefore counted on the side.  The distinct records overwrite the head of the bucket's segment.
// uniques != null: +1 per distinct word on its neuron (kmer_per_neuron).  *overflow is raised if a bucket holds
// more distinct words than the table takes (the caller re-partitions into more buckets).
__global__ void __launch_bounds__(XT) bucket_dedup_kernel(ExactSlot* __restrict__ recs, const unsigned long long* __restrict__ start,
                                                           const unsigned long long* __restrict__ filled,
                                                           unsigned int* __restrict__ distinct, unsigned int* __restrict__ uniques,
                                                           unsigned long long* __restrict__ n_distinct_total,
                                                           unsigned int* __restrict__ overflow, BucketPlan bp) {
    extern __shared__ __align__(16) unsigned char dedup_smem[];  // keys | counts | neuron indices | list of claimed slots
    unsigned long long* s_key = reinterpret_cast<unsigned long long*>(dedup_smem);
    unsigned int* s_cnt = reinterpret_cast<unsigned int*>(s_key + TABLE_SLOTS);
    unsigned int* s_ix = s_cnt + TABLE_SLOTS;
    unsigned short* s_list = reinterpret_cast<unsigned short*>(s_ix + TABLE_SLOTS);  // TABLE_SLOTS entries
    __shared__ unsigned int s_ones_cnt, s_ones_ix, s_over, s_nd;
    // kmer_per_neuron: a bucket of whole neurons counts its distinct words per neuron in shared memory and adds each
    // neuron's number once (2 M global atomics per job instead of one per distinct word: 113 M)
    constexpr unsigned LOCAL_UNI = 64;
    __shared__ unsigned int s_uni[LOCAL_UNI];
    const bool local_uni = uniques != nullptr && bp.splits == 1u && bp.neurons_per_bucket <= LOCAL_UNI;
    const unsigned long long b = blockIdx.x;
    // `filled` = the scatter's cursors: a segment sized from the neurons' counts is longer than its records when the
    // count kernel summed a run of N into one weighted record
    const unsigned long long lo = start[b], hi = filled[b];
    const unsigned tid = threadIdx.x;
    if (lo == hi) {
        if (tid == 0) distinct[b] = 0u;
        return;
    }
    // keys <- EMPTY, counts <- 0 (16-byte stores; the two arrays are adjacent)
    {
        uint4* k4 = reinterpret_cast<uint4*>(s_key);
        uint4* c4 = reinterpret_cast<uint4*>(s_cnt);
        for (unsigned i = tid; i < TABLE_SLOTS / 2; i += XT) k4[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
        for (unsigned i = tid; i < TABLE_SLOTS / 4; i += XT) c4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) { s_ones_cnt = 0u; s_ones_ix = 0u; s_over = 0u; s_nd = 0u; }
    if (tid < LOCAL_UNI) s_uni[tid] = 0u;
    const unsigned neuron0 = (unsigned)(b * bp.neurons_per_bucket);   // (local_uni only)
    __syncthreads();
    for (unsigned long long i = lo + tid; i < hi; i += XT) {
        const uint4 q = reinterpret_cast<const uint4*>(recs)[i];
        const unsigned long long w = ((unsigned long long)q.y << 32) | q.x;
        const unsigned wt = q.z, ix = q.w;
        if (w == EMPTY) { atomicAdd(&s_ones_cnt, wt); s_ones_ix = ix; continue; }
        unsigned slot = (unsigned)(mix64(w) >> 40) & (TABLE_SLOTS - 1);
        for (unsigned probes = 0;; ++probes) {
            unsigned long long prev = s_key[slot];                          // plain load first: a word that is already
            if (prev == EMPTY) prev = atomicCAS(&s_key[slot], EMPTY, w);    // there costs no CAS (N runs repeat one word)
            if (prev == EMPTY) {  // claimed: remember the slot (the distinct words are written out from this list)
                s_ix[slot] = ix;
                s_list[atomicAdd(&s_nd, 1u)] = (unsigned short)slot;
            }
            if (prev == EMPTY || prev == w) {
                atomicAdd(&s_cnt[slot], wt);
                break;
            }
            slot = (slot + 1) & (TABLE_SLOTS - 1);
            if (probes >= TABLE_SLOTS) { s_over = 1u; break; }
        }
    }
    __syncthreads();
    if (s_over) {
        if (tid == 0) atomicExch(overflow, 1u);
        return;
    }
    // the distinct records overwrite the head of the segment (every load above is done: barrier) — unless every record
    // was distinct already: then the segment IS the table (weights and indices untouched) and nothing is written
    const unsigned nd0 = s_nd;
    const bool all_distinct = (unsigned long long)nd0 + (s_ones_cnt ? 1u : 0u) == hi - lo;
    for (unsigned j = tid; j < nd0; j += XT) {
        const unsigned slot = s_list[j];
        const unsigned ix = s_ix[slot];
        if (!all_distinct) {
            const unsigned long long k = s_key[slot];
            reinterpret_cast<uint4*>(recs)[lo + j] = make_uint4((unsigned)k, (unsigned)(k >> 32), s_cnt[slot], ix);
        }
        if (local_uni) atomicAdd(&s_uni[ix - neuron0], 1u);
        else if (uniques) atomicAdd(uniques + ix, 1u);
    }
    if (tid == 0) {
        unsigned nd = nd0;
        if (s_ones_cnt) {
            if (!all_distinct) reinterpret_cast<uint4*>(recs)[lo + nd] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, s_ones_cnt, s_ones_ix);
            if (local_uni) atomicAdd(&s_uni[s_ones_ix - neuron0], 1u);
            else if (uniques) atomicAdd(uniques + s_ones_ix, 1u);
            ++nd;
        }
        distinct[b] = nd;
        if (nd) atomicAdd(n_distinct_total, (unsigned long long)nd);
    }
    if (local_uni) {
        __syncthreads();
        if (tid < bp.neurons_per_bucket && s_uni[tid]) atomicAdd(uniques + neuron0 + tid, s_uni[tid]);
    }
}

// process_sequence: every neuron touched by this sequence gets +1 (local_unique, :204,226,262-264)
__global__ void touched_kernel(const unsigned int* __restrict__ widx, unsigned long long n, unsigned int* flags,
                               unsigned int* uniques, int phase) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned ix = widx[i];
    if (phase == 0) {
        if (atomicExch(flags + ix, 1u) == 0u) atomicAdd(uniques + ix, 1u);
    } else {
        flags[ix] = 0u;
    }
}

// get_count: one warp scans the word's bucket
template <bool POW2>
__global__ void lookup_kernel(const ExactSlot* __restrict__ recs, const unsigned long long* __restrict__ start,
                              const unsigned int* __restrict__ distinct, BucketPlan bp, FastMod fm, RotMul rm,
                              unsigned long long key, unsigned long long* out) {
    const U64 h = siphash13_dev((unsigned)key, (unsigned)(key >> 32), rm);
    const unsigned ix = fastmod_dev<POW2>(h, fm);
    const unsigned long long b = bucket_of(ix, key, bp);
    const unsigned long long lo = start[b];
    const unsigned nd = distinct[b];
    unsigned long long found = 0, cnt = 0;
    for (unsigned j = threadIdx.x; j < nd; j += 32)
        if (recs[lo + j].key == key) { found = 1; cnt = recs[lo + j].count; }
    for (int o = 16; o > 0; o >>= 1) {
        found |= __shfl_down_sync(0xFFFFFFFFu, found, o);
        cnt |= __shfl_down_sync(0xFFFFFFFFu, cnt, o);  // at most one lane holds a non-zero count
    }
    if (threadIdx.x == 0) { out[0] = found; out[1] = cnt; }
}

// dense copy of the table (bucket order): one block per bucket, offsets from a scan of the distinct counts
__global__ void compact_table_kernel(const ExactSlot* __restrict__ recs, const unsigned long long* __restrict__ start,
                                     const unsigned int* __restrict__ distinct, const unsigned long long* __restrict__ dense_start,
                                     unsigned long long* __restrict__ out_keys, unsigned int* __restrict__ out_counts,
                                     ExactSlot* __restrict__ out_recs) {
    const unsigned long long b = blockIdx.x;
    const unsigned long long lo = start[b], d0 = dense_start[b];
    const unsigned nd = distinct[b];
    for (unsigned j = threadIdx.x; j < nd; j += blockDim.x) {
        const ExactSlot r = recs[lo + j];
        if (out_keys) out_keys[d0 + j] = r.key;
        if (out_counts) out_counts[d0 + j] = r.count;
        if (out_recs) out_recs[d0 + j] = r;
    }
}

__global__ void gather_uniques_kernel(const unsigned long long* __restrict__ idx, unsigned long long n,
                                      const unsigned int* __restrict__ uniques, unsigned int* out) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = uniques[idx[i]];
}

__global__ void filter_set_kernel(unsigned int* filter, const unsigned long long* __restrict__ idx, unsigned long long n) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i < n) atomicOr(filter + (idx[i] >> 5), 1u << (idx[i] & 31u));
}

BucketPlan make_plan(unsigned long long n, unsigned long long pool, unsigned extra_split) {
    BucketPlan bp;
    unsigned long long want = (n + BUCKET_TARGET - 1) / BUCKET_TARGET;
    if (want < 1) want = 1;
    want *= extra_split;
    if (want > 0x7FFFFFFFull) want = 0x7FFFFFFFull;   // one CTA per bucket: grid limit
    if (pool >= want) {
        bp.neurons_per_bucket = (unsigned)((pool + want - 1) / want);
        bp.splits = 1;
    } else {
        bp.neurons_per_bucket = 1;
        bp.splits = (unsigned)((want + pool - 1) / pool);
    }
    const unsigned long long groups = (pool + bp.neurons_per_bucket - 1) / bp.neurons_per_bucket;
    bp.nbuckets = groups * bp.splits;
    return bp;
}

unsigned grid_for(unsigned long long n) {
    unsigned long long blocks = (n + XT - 1) / XT;
    if (blocks > 148ull * 16) blocks = 148ull * 16;
    return (unsigned)(blocks ? blocks : 1);
}

cudaError_t ensure_cursor(ExactTable& t, cudaStream_t s) {
    if (!t.cursor) {
        NKX(cudaMalloc(&t.cursor, 8 * sizeof(unsigned long long)));
        NKX(cudaMemsetAsync(t.cursor, 0, 8 * sizeof(unsigned long long), s));
    }
    return cudaSuccess;
}

}  // namespace

cudaError_t exact_reserve_words(ExactTable& t, unsigned long long extra, cudaStream_t s) {
    NKX(ensure_cursor(t, s));
    const unsigned long long need = t.words_bound + extra;
    if (need > t.words_cap) NKX(exact_grow_words(t, need + need / 2 + 1024, t.words_bound, s));
    t.words_bound = need;
    return cudaSuccess;
}

cudaError_t exact_grow_words(ExactTable& t, unsigned long long cap, unsigned long long keep, cudaStream_t s) {
    NKX(ensure_cursor(t, s));
    if (cap <= t.words_cap) return cudaSuccess;
    unsigned long long* nw = nullptr;
    unsigned int* ni = nullptr;
    NKX(cudaMalloc(&nw, cap * sizeof(unsigned long long)));
    cudaError_t e = cudaMalloc(&ni, cap * sizeof(unsigned int));
    if (e != cudaSuccess) { cudaFree(nw); return e; }
    if (t.words) {
        if (keep) {
            NKX(cudaMemcpyAsync(nw, t.words, keep * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
            NKX(cudaMemcpyAsync(ni, t.widx, keep * sizeof(unsigned int), cudaMemcpyDeviceToDevice, s));
        }
        NKX(cudaStreamSynchronize(s));
        cudaFree(t.words);
        cudaFree(t.widx);
    }
    t.words = nw;
    t.widx = ni;
    t.words_cap = cap;
    return cudaSuccess;
}

cudaError_t launch_filter_set(unsigned int* filter, const unsigned long long* idx, unsigned long long n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    filter_set_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(filter, idx, n);
    return cudaGetLastError();
}

cudaError_t exact_clear(ExactTable& t, unsigned long long pool, bool tables_too, cudaStream_t s) {
    t.words_bound = 0;
    if (t.cursor) NKX(cudaMemsetAsync(t.cursor, 0, 4 * sizeof(unsigned long long), s));   // append cursor ... run-of-N count
    if (tables_too) {
        t.n_keys = 0;
        t.valid = false;
        if (t.uniques) NKX(cudaMemsetAsync(t.uniques, 0, pool * sizeof(unsigned int), s));
    }
    return cudaSuccess;
}

// The (word, index) records appended since the last call -> the bucketed table.  merge: ADD them to the table
// that exists (process_sequence: counts accumulate over calls, :218-221) and use its "neurons touched by this
// sequence" rule for the per-neuron column; else the table is replaced (counts.clear(), :157 / :426).
// pool_counts (may be null): per-neuron counts of exactly the appended windows (the call's u64 currents).
// records -> table: bucket sizes, scan, scatter, per-bucket dedup; re-partitions finer if a bucket overflows its
// on-chip table.  Input: the n_new appended windows (weight 1) and n_old weighted records (`old_dense`).
static cudaError_t partition_and_dedup(ExactTable& t, unsigned long long pool, const unsigned long long* pool_counts, bool merge,
                                       unsigned long long n_new, const ExactSlot* old_dense, unsigned long long n_old,
                                       unsigned long long n_zero, unsigned int flo, unsigned int fhi, cudaStream_t s) {
    const unsigned long long n_in = n_new + n_old;
    cudaError_t err = cudaSuccess;
    for (unsigned attempt = 0, split = 1; attempt < 4; ++attempt, split *= 4) {
        const BucketPlan bp = make_plan(n_in, pool, split);
        err = ensure((void**)&t.bucket_count, &t.bucket_count_cap, bp.nbuckets * sizeof(unsigned int));
        if (err == cudaSuccess) err = ensure((void**)&t.bucket_distinct, &t.bucket_distinct_cap, bp.nbuckets * sizeof(unsigned int));
        if (err == cudaSuccess) err = ensure((void**)&t.bucket_start, &t.bucket_start_cap, (bp.nbuckets + 1) * sizeof(unsigned long long));
        if (err == cudaSuccess) err = ensure((void**)&t.bucket_cursor, &t.bucket_cursor_cap, (bp.nbuckets + 1) * sizeof(unsigned long long));
        // (segments sized from the neurons' counts also hold room for the run-of-N windows that arrive as one record)
        if (err == cudaSuccess) err = ensure((void**)&t.recs, &t.recs_cap, (n_in + n_zero) * sizeof(ExactSlot));
        if (err == cudaSuccess) err = cudaMemsetAsync(t.cursor + 1, 0, 2 * sizeof(unsigned long long), s);  // [1] distinct total, [2] overflow
        if (err != cudaSuccess) break;
        if (pool_counts && !merge && bp.splits == 1) {
            bucket_hist_from_pool_kernel<<<(unsigned)((bp.nbuckets + 255) / 256), 256, 0, s>>>(pool_counts, pool, bp, t.bucket_count);
        } else {
            err = cudaMemsetAsync(t.bucket_count, 0, bp.nbuckets * sizeof(unsigned int), s);
            if (err != cudaSuccess) break;
            bucket_hist_kernel<<<grid_for(n_in), XT, 0, s>>>(t.words, t.widx, old_dense, n_old, n_new, bp, flo, fhi, t.bucket_count);
        }
        bucket_scan_kernel<<<1, 1024, 0, s>>>(t.bucket_count, bp.nbuckets, t.bucket_start, t.bucket_cursor);
        bucket_scatter_kernel<<<grid_for(n_in), XT, 0, s>>>(t.words, t.widx, old_dense, n_old, n_new, bp, flo, fhi, t.bucket_cursor, t.recs);
        constexpr int kDedupSmem = TABLE_SLOTS * (8 + 4 + 4 + 2);
        static std::atomic<bool> smem_set[64];
        int dev = 0;
        err = cudaGetDevice(&dev);
        if (err != cudaSuccess) break;
        if (dev >= 0 && dev < 64 && !smem_set[dev].load()) {
            err = cudaFuncSetAttribute(bucket_dedup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDedupSmem);
            if (err != cudaSuccess) break;
            smem_set[dev].store(true);
        }
        bucket_dedup_kernel<<<(unsigned)bp.nbuckets, XT, kDedupSmem, s>>>(t.recs, t.bucket_start, t.bucket_cursor, t.bucket_distinct,
                                                                         merge ? nullptr : t.uniques, t.cursor + 1,
                                                                         reinterpret_cast<unsigned int*>(t.cursor + 2), bp);
        err = cudaGetLastError();
        if (err != cudaSuccess) break;
        unsigned long long res[2] = {0, 0};
        err = cudaMemcpyAsync(res, t.cursor + 1, sizeof res, cudaMemcpyDeviceToHost, s);
        if (err == cudaSuccess) err = cudaStreamSynchronize(s);
        if (err != cudaSuccess) break;
        if ((res[1] & 0xFFFFFFFFull) == 0) {
            t.plan = bp;
            t.n_keys = res[0];
            break;
        }
        // a bucket held more distinct words than the table takes: partition finer.  `uniques` was touched by the
        // buckets that did finish: start over with it
        if (!merge) {
            err = cudaMemsetAsync(t.uniques, 0, pool * sizeof(unsigned int), s);
            if (err != cudaSuccess) break;
        }
        err = cudaErrorInvalidValue;
    }
    return err;
}

cudaError_t exact_finalize(ExactTable& t, const FastMod& fm, unsigned long long pool, const unsigned long long* pool_counts,
                           bool merge, cudaStream_t s) {
    (void)fm;
    NKX(ensure_cursor(t, s));
    unsigned long long cur4[4] = {0, 0, 0, 0};
    NKX(cudaMemcpyAsync(cur4, t.cursor, sizeof cur4, cudaMemcpyDeviceToHost, s));
    NKX(cudaStreamSynchronize(s));
    unsigned long long n_new = cur4[0];
    if (n_new > t.words_cap) n_new = t.words_cap;  // (the uniques pass lets its cursor run past the capacity)
    // windows inside runs of N, summed by the count kernel (nk_count.cu): ONE record {word 0, their number}
    const unsigned long long n_zero = cur4[3];
    ExactSlot* const zero_rec = reinterpret_cast<ExactSlot*>(t.cursor + 4);
    if (n_zero) {
        ExactSlot z;
        z.key = 0ull;
        z.count = (unsigned int)n_zero;
        z.idx = fastmod_u64(siphash13_u64(0ull), fm);
        NKX(cudaMemcpyAsync(zero_rec, &z, sizeof z, cudaMemcpyHostToDevice, s));
        NKX(cudaStreamSynchronize(s));   // `z` is a local
    }
    const unsigned long long n_w = n_zero ? 1ull : 0ull;
    if (!t.uniques) {
        NKX(cudaMalloc(&t.uniques, pool * sizeof(unsigned int)));
        NKX(cudaMemsetAsync(t.uniques, 0, pool * sizeof(unsigned int), s));
    }
    if (!merge) {
        t.n_keys = 0;
        NKX(cudaMemsetAsync(t.uniques, 0, pool * sizeof(unsigned int), s));
    }
    if (merge && n_new + n_w > 0) {
        if (!t.flags) {
            NKX(cudaMalloc(&t.flags, pool * sizeof(unsigned int)));
            NKX(cudaMemsetAsync(t.flags, 0, pool * sizeof(unsigned int), s));
        }
        const unsigned blocks = (unsigned)((n_new + 255) / 256);
        if (n_new) touched_kernel<<<blocks, 256, 0, s>>>(t.widx, n_new, t.flags, t.uniques, 0);
        if (n_w) touched_kernel<<<1, 32, 0, s>>>(&zero_rec->idx, 1, t.flags, t.uniques, 0);
        if (n_new) touched_kernel<<<blocks, 256, 0, s>>>(t.widx, n_new, t.flags, t.uniques, 1);
        if (n_w) touched_kernel<<<1, 32, 0, s>>>(&zero_rec->idx, 1, t.flags, t.uniques, 1);
    }
    // input records of the partition: the new windows (weight 1) and, when merging, the table's records
    const unsigned long long n_tab = merge ? t.n_keys : 0;
    const unsigned long long n_old = n_tab + n_w;          // weighted records: the old table and the run-of-N record
    if (n_new + n_w > 0) {
        ExactSlot* old_dense = nullptr;
        ExactSlot* old_alloc = nullptr;
        if (n_tab) {  // the old table, dense (it is rebuilt together with the new windows)
            NKX(cudaMalloc(&old_alloc, n_old * sizeof(ExactSlot)));
            old_dense = old_alloc;
            cudaError_t e = exact_dense_copy(t, nullptr, nullptr, old_dense, s);
            if (e == cudaSuccess && n_w) e = cudaMemcpyAsync(old_dense + n_tab, zero_rec, sizeof(ExactSlot), cudaMemcpyDeviceToDevice, s);
            if (e != cudaSuccess) { cudaFree(old_alloc); return e; }
        } else if (n_w) {
            old_dense = zero_rec;
        }
        const cudaError_t err = partition_and_dedup(t, pool, pool_counts, merge, n_new, old_dense, n_old, n_zero, 0u, 0xFFFFFFFFu, s);
        if (old_alloc) { cudaStreamSynchronize(s); cudaFree(old_alloc); }
        if (err != cudaSuccess) return err;
    }
    t.valid = true;
    // the words of this call are consumed
    t.words_bound = 0;
    NKX(cudaMemsetAsync(t.cursor, 0, sizeof(unsigned long long), s));
    NKX(cudaMemsetAsync(t.cursor + 3, 0, sizeof(unsigned long long), s));
    return cudaGetLastError();
}

// the table as dense device arrays, bucket order (any of the outputs may be null); needs n_keys entries each
cudaError_t exact_dense_copy(ExactTable& t, unsigned long long* out_keys, unsigned int* out_counts, ExactSlot* out_recs,
                             cudaStream_t s) {
    if (t.n_keys == 0) return cudaSuccess;
    const unsigned long long nb = t.plan.nbuckets;
    NKX(ensure((void**)&t.dense_start, &t.dense_start_cap, (nb + 1) * sizeof(unsigned long long)));
    bucket_scan_kernel<<<1, 1024, 0, s>>>(t.bucket_distinct, nb, t.dense_start, nullptr);
    compact_table_kernel<<<(unsigned)nb, 128, 0, s>>>(t.recs, t.bucket_start, t.bucket_distinct, t.dense_start, out_keys, out_counts, out_recs);
    return cudaGetLastError();
}

// ---- tables of a multi-GPU group: every GPU builds the table of its own share of the windows, then GPU d collects
// the records of ITS neuron slice from all of them and merges equal words (src/spiking_hash.rs:157-172: one map) ----

// The dense copy of `t` (bucket order = neuron order) into out_recs (n_keys records) and, for every neuron range
// [bounds[i], bounds[i+1]) (i < n_ranges), the dense range of the buckets that hold it: first[i] .. last[i]
// (boundary buckets are shared by two ranges: the reader filters by neuron index).
cudaError_t exact_dense_export(ExactTable& t, ExactSlot* out_recs, const unsigned long long* bounds, int n_ranges,
                               unsigned long long* first, unsigned long long* last, cudaStream_t s) {
    for (int i = 0; i < n_ranges; ++i) first[i] = last[i] = 0;
    if (!t.valid || t.n_keys == 0) return cudaSuccess;
    NKX(exact_dense_copy(t, nullptr, nullptr, out_recs, s));
    const unsigned long long npb = t.plan.neurons_per_bucket, sp = t.plan.splits, nb = t.plan.nbuckets;
    for (int i = 0; i < n_ranges; ++i) {
        if (bounds[i + 1] <= bounds[i]) continue;
        unsigned long long b0 = (bounds[i] / npb) * sp, b1 = ((bounds[i + 1] - 1) / npb + 1) * sp;
        if (b0 > nb) b0 = nb;
        if (b1 > nb) b1 = nb;
        NKX(cudaMemcpyAsync(first + i, t.dense_start + b0, 8, cudaMemcpyDeviceToHost, s));
        NKX(cudaMemcpyAsync(last + i, t.dense_start + b1, 8, cudaMemcpyDeviceToHost, s));
    }
    return cudaStreamSynchronize(s);
}

// `t` <- the table of the weighted records `recs` restricted to the neurons [flo, fhi)
cudaError_t exact_build_from_records(ExactTable& t, unsigned long long pool, const ExactSlot* recs, unsigned long long n_recs,
                                     unsigned int flo, unsigned int fhi, cudaStream_t s) {
    NKX(ensure_cursor(t, s));
    if (!t.uniques) NKX(cudaMalloc(&t.uniques, pool * sizeof(unsigned int)));
    NKX(cudaMemsetAsync(t.uniques, 0, pool * sizeof(unsigned int), s));
    t.n_keys = 0;
    t.valid = false;
    if (n_recs) NKX(partition_and_dedup(t, pool, nullptr, false, 0, recs, n_recs, 0, flo, fhi, s));
    t.valid = true;
    return cudaSuccess;
}

cudaError_t exact_lookup(const ExactTable& t, const FastMod& fm, unsigned long long key, unsigned long long* d_out2, cudaStream_t s) {
    if (fm.is_pow2) lookup_kernel<true><<<1, 32, 0, s>>>(t.recs, t.bucket_start, t.bucket_distinct, t.plan, fm, make_rotmul(), key, d_out2);
    else lookup_kernel<false><<<1, 32, 0, s>>>(t.recs, t.bucket_start, t.bucket_distinct, t.plan, fm, make_rotmul(), key, d_out2);
    return cudaGetLastError();
}

cudaError_t exact_gather_uniques(const ExactTable& t, const unsigned long long* idx, unsigned long long n,
                                 unsigned int* out, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    gather_uniques_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(idx, n, t.uniques, out);
    return cudaGetLastError();
}

void exact_free(ExactTable& t) {
    cudaFree(t.words); cudaFree(t.widx); cudaFree(t.cursor); cudaFree(t.recs); cudaFree(t.uniques); cudaFree(t.flags);
    cudaFree(t.bucket_count); cudaFree(t.bucket_distinct); cudaFree(t.bucket_start); cudaFree(t.bucket_cursor); cudaFree(t.dense_start);
    t = ExactTable{};
}

}  // namespace nk
