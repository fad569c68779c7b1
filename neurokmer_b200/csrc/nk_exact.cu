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
//   1. histogram of the windows over BUCKETS of consecutive neurons (~2 K windows per bucket; pools with few
//      neurons split every neuron into sub-buckets by a mix of the word)           bucket_hist_kernel
//   2. one-block exclusive scan of the bucket sizes                                 bucket_scan_kernel
//   3. scatter of the (word, index, weight) records to their bucket                 bucket_scatter_kernel
//   4. one CTA per bucket: a shared-memory hash table (64-bit atomicCAS) merges equal words and counts them,
//      every NEW word adds one to its neuron's `uniques`; the distinct (word, count, index) records are written
//      back IN PLACE at the head of the bucket's own segment                        bucket_dedup_kernel
// get_count(word) = hash -> neuron -> bucket -> one warp scans that bucket's distinct records.
// The table is therefore grouped by neuron range and unordered inside a bucket (like the reference's map);
// nk_copy_exact_table compacts the buckets into dense arrays.  No 2^31 limit: every cursor is 64-bit.
// Memory: 12 B per window appended + 16 B per window while partitioning.  Counts wrap at 2^32 like the
// reference's AtomicU32::fetch_add.
#include <atomic>

#include "nk_kernels.cuh"

namespace nk {

namespace {

constexpr int XT = 256;                       // threads per block
constexpr unsigned TABLE_SLOTS = 4096;        // shared-memory hash table of one bucket (power of two)
constexpr unsigned long long BUCKET_TARGET = 1536;  // windows per bucket aimed at (table load <= ~0.4 when all distinct)
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

__global__ void bucket_hist_kernel(const unsigned long long* __restrict__ words, const unsigned int* __restrict__ widx,
                                   unsigned long long n, BucketPlan bp, unsigned int* __restrict__ bucket_count) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += stride)
        atomicAdd(bucket_count + bucket_of(widx[i], bp.splits > 1 ? words[i] : 0ull, bp), 1u);
}

// one block: bucket_start[b] = sum of bucket_count[< b]; bucket_start[nb] = n; also zeroes the fill cursors
__global__ void __launch_bounds__(1024) bucket_scan_kernel(const unsigned int* __restrict__ count, unsigned long long nb,
                                                            unsigned long long* __restrict__ start, unsigned int* __restrict__ fill) {
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
        fill[i] = 0u;
        s += count[i];
    }
}

__global__ void bucket_scatter_kernel(const unsigned long long* __restrict__ words, const unsigned int* __restrict__ widx,
                                      const unsigned int* __restrict__ weight, unsigned long long n, BucketPlan bp,
                                      const unsigned long long* __restrict__ start, unsigned int* __restrict__ fill,
                                      unsigned long long* __restrict__ out_words, unsigned int* __restrict__ out_idx,
                                      unsigned int* __restrict__ out_weight) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long w = words[i];
        const unsigned ix = widx[i];
        const unsigned long long b = bucket_of(ix, w, bp);
        const unsigned long long pos = start[b] + atomicAdd(fill + b, 1u);
        out_words[pos] = w;
        out_idx[pos] = ix;
        out_weight[pos] = weight ? weight[i] : 1u;
    }
}

// One CTA per bucket.  Equal words are merged in a shared-memory hash table (linear probing, 64-bit atomicCAS
// on the key).  The all-ones word (only pack_kmer of 32 T's, k = 32, non-canonical) is the table's EMPTY marker
// and therefore counted on the side.  The distinct records overwrite the head of the bucket's segment.
// uniques != null: +1 per distinct word on its neuron (kmer_per_neuron).  *overflow is raised if a bucket holds
// more distinct words than the table takes (the caller re-partitions into more buckets).
__global__ void __launch_bounds__(XT) bucket_dedup_kernel(unsigned long long* __restrict__ words, unsigned int* __restrict__ idx,
                                                           unsigned int* __restrict__ weight,
                                                           const unsigned long long* __restrict__ start,
                                                           unsigned int* __restrict__ distinct, unsigned int* __restrict__ uniques,
                                                           unsigned long long* __restrict__ n_distinct_total,
                                                           unsigned int* __restrict__ overflow) {
    extern __shared__ __align__(16) unsigned char dedup_smem[];  // 64 KB: keys, counts, neuron indices
    unsigned long long* s_key = reinterpret_cast<unsigned long long*>(dedup_smem);
    unsigned int* s_cnt = reinterpret_cast<unsigned int*>(s_key + TABLE_SLOTS);
    unsigned int* s_ix = s_cnt + TABLE_SLOTS;
    __shared__ unsigned int s_warp[XT / 32];
    __shared__ unsigned int s_ones_cnt, s_ones_ix, s_over, s_base;
    const unsigned long long b = blockIdx.x;
    const unsigned long long lo = start[b], hi = start[b + 1];
    const unsigned tid = threadIdx.x;
    for (unsigned s = tid; s < TABLE_SLOTS; s += XT) { s_key[s] = EMPTY; s_cnt[s] = 0u; }
    if (tid == 0) { s_ones_cnt = 0u; s_ones_ix = 0u; s_over = 0u; s_base = 0u; }
    __syncthreads();
    for (unsigned long long i = lo + tid; i < hi; i += XT) {
        const unsigned long long w = words[i];
        const unsigned ix = idx[i], wt = weight[i];
        if (w == EMPTY) { atomicAdd(&s_ones_cnt, wt); s_ones_ix = ix; continue; }
        unsigned slot = (unsigned)(mix64(w) >> 40) & (TABLE_SLOTS - 1);
        for (unsigned probes = 0;; ++probes) {
            const unsigned long long prev = atomicCAS(&s_key[slot], EMPTY, w);
            if (prev == EMPTY || prev == w) {
                atomicAdd(&s_cnt[slot], wt);
                if (prev == EMPTY) s_ix[slot] = ix;
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
    // compact the occupied slots to the head of the segment (the loads above are all done: barrier)
    for (unsigned base = 0; base < TABLE_SLOTS; base += XT) {
        const unsigned s = base + tid;
        const bool occ = s_key[s] != EMPTY;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, occ);
        if ((tid & 31) == 0) s_warp[tid >> 5] = __popc(bal);
        __syncthreads();
        unsigned before = s_base;
        for (unsigned w = 0; w < (tid >> 5); ++w) before += s_warp[w];
        const unsigned pos = before + __popc(bal & ((1u << (tid & 31)) - 1u));
        if (occ) {
            words[lo + pos] = s_key[s];
            idx[lo + pos] = s_ix[s];
            weight[lo + pos] = s_cnt[s];
            if (uniques) atomicAdd(uniques + s_ix[s], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned t = 0;
            for (int w = 0; w < XT / 32; ++w) t += s_warp[w];
            s_base += t;
        }
        __syncthreads();
    }
    if (tid == 0) {
        unsigned nd = s_base;
        if (s_ones_cnt) {
            words[lo + nd] = EMPTY;
            idx[lo + nd] = s_ones_ix;
            weight[lo + nd] = s_ones_cnt;
            if (uniques) atomicAdd(uniques + s_ones_ix, 1u);
            ++nd;
        }
        distinct[b] = nd;
        if (nd) atomicAdd(n_distinct_total, (unsigned long long)nd);
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
__global__ void lookup_kernel(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ counts,
                              const unsigned long long* __restrict__ start, const unsigned int* __restrict__ distinct,
                              BucketPlan bp, FastMod fm, RotMul rm, unsigned long long key, unsigned long long* out) {
    const U64 h = siphash13_dev((unsigned)key, (unsigned)(key >> 32), rm);
    const unsigned ix = fastmod_dev<POW2>(h, fm);
    const unsigned long long b = bucket_of(ix, key, bp);
    const unsigned long long lo = start[b];
    const unsigned nd = distinct[b];
    unsigned long long found = 0, cnt = 0;
    for (unsigned j = threadIdx.x; j < nd; j += 32)
        if (keys[lo + j] == key) { found = 1; cnt = counts[lo + j]; }
    for (int o = 16; o > 0; o >>= 1) {
        found |= __shfl_down_sync(0xFFFFFFFFu, found, o);
        cnt |= __shfl_down_sync(0xFFFFFFFFu, cnt, o);  // at most one lane holds a non-zero count
    }
    if (threadIdx.x == 0) { out[0] = found; out[1] = cnt; }
}

// dense copy of the table (bucket order): one block per bucket, offsets from a scan of the distinct counts
__global__ void compact_table_kernel(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ counts,
                                     const unsigned int* __restrict__ idx, const unsigned long long* __restrict__ start,
                                     const unsigned int* __restrict__ distinct, const unsigned long long* __restrict__ dense_start,
                                     unsigned long long* __restrict__ out_keys, unsigned int* __restrict__ out_counts,
                                     unsigned int* __restrict__ out_idx) {
    const unsigned long long b = blockIdx.x;
    const unsigned long long lo = start[b], d0 = dense_start[b];
    const unsigned nd = distinct[b];
    for (unsigned j = threadIdx.x; j < nd; j += blockDim.x) {
        if (out_keys) out_keys[d0 + j] = keys[lo + j];
        if (out_counts) out_counts[d0 + j] = counts[lo + j];
        if (out_idx) out_idx[d0 + j] = idx[lo + j];
    }
}

__global__ void fill_u32_kernel(unsigned int* p, unsigned long long n, unsigned int v) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
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

}  // namespace

cudaError_t exact_reserve_words(ExactTable& t, unsigned long long extra, cudaStream_t s) {
    if (!t.cursor) {
        NKX(cudaMalloc(&t.cursor, 4 * sizeof(unsigned long long)));
        NKX(cudaMemsetAsync(t.cursor, 0, 4 * sizeof(unsigned long long), s));
    }
    const unsigned long long need = t.words_bound + extra;
    if (need > t.words_cap) NKX(exact_grow_words(t, need + need / 2 + 1024, t.words_bound, s));
    t.words_bound = need;
    return cudaSuccess;
}

cudaError_t exact_grow_words(ExactTable& t, unsigned long long cap, unsigned long long keep, cudaStream_t s) {
    if (!t.cursor) {
        NKX(cudaMalloc(&t.cursor, 4 * sizeof(unsigned long long)));
        NKX(cudaMemsetAsync(t.cursor, 0, 4 * sizeof(unsigned long long), s));
    }
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
    if (t.cursor) NKX(cudaMemsetAsync(t.cursor, 0, sizeof(unsigned long long), s));
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
cudaError_t exact_finalize(ExactTable& t, const FastMod& fm, unsigned long long pool, unsigned /*key_bits*/, bool merge,
                           cudaStream_t s) {
    unsigned long long n_new = 0;
    if (t.cursor) {
        NKX(cudaMemcpyAsync(&n_new, t.cursor, sizeof n_new, cudaMemcpyDeviceToHost, s));
        NKX(cudaStreamSynchronize(s));
    }
    if (n_new > t.words_cap) n_new = t.words_cap;  // (the uniques pass lets its cursor run past the capacity)
    if (!t.uniques) {
        NKX(cudaMalloc(&t.uniques, pool * sizeof(unsigned int)));
        NKX(cudaMemsetAsync(t.uniques, 0, pool * sizeof(unsigned int), s));
    }
    if (!merge) {
        t.n_keys = 0;
        NKX(cudaMemsetAsync(t.uniques, 0, pool * sizeof(unsigned int), s));
    }
    if (merge && n_new > 0) {
        if (!t.flags) {
            NKX(cudaMalloc(&t.flags, pool * sizeof(unsigned int)));
            NKX(cudaMemsetAsync(t.flags, 0, pool * sizeof(unsigned int), s));
        }
        const unsigned blocks = (unsigned)((n_new + 255) / 256);
        touched_kernel<<<blocks, 256, 0, s>>>(t.widx, n_new, t.flags, t.uniques, 0);
        touched_kernel<<<blocks, 256, 0, s>>>(t.widx, n_new, t.flags, t.uniques, 1);
    }
    // input records of the partition: the new windows (weight 1) and, when merging, the table's records
    const unsigned long long n_old = merge ? t.n_keys : 0;
    const unsigned long long n = n_new + n_old;
    if (n > 0 && !(merge && n_new == 0)) {
        unsigned long long* in_words = t.words;
        unsigned int* in_idx = t.widx;
        unsigned int* in_weight = nullptr;
        unsigned long long *cat_w = nullptr;
        unsigned int *cat_i = nullptr, *cat_c = nullptr;
        if (n_old) {
            // old records (bucket-sparse -> dense) followed by the new ones
            NKX(cudaMalloc(&cat_w, n * 8)); NKX(cudaMalloc(&cat_i, n * 4)); NKX(cudaMalloc(&cat_c, n * 4));
            NKX(exact_dense_copy(t, cat_w, cat_c, cat_i, s));
            const unsigned long long nk = t.n_keys;
            NKX(cudaMemcpyAsync(cat_w + nk, t.words, n_new * 8, cudaMemcpyDeviceToDevice, s));
            NKX(cudaMemcpyAsync(cat_i + nk, t.widx, n_new * 4, cudaMemcpyDeviceToDevice, s));
            fill_u32_kernel<<<grid_for(n_new), XT, 0, s>>>(cat_c + nk, n_new, 1u);  // every new window weighs 1
            in_words = cat_w; in_idx = cat_i; in_weight = cat_c;
        }
        const unsigned long long n_in = n;
        cudaError_t err = cudaSuccess;
        for (unsigned attempt = 0, split = 1; attempt < 4; ++attempt, split *= 4) {
            const BucketPlan bp = make_plan(n_in, pool, split);
            NKX(ensure((void**)&t.bucket_count, &t.bucket_count_cap, bp.nbuckets * sizeof(unsigned int)));
            NKX(ensure((void**)&t.bucket_fill, &t.bucket_fill_cap, bp.nbuckets * sizeof(unsigned int)));
            NKX(ensure((void**)&t.bucket_distinct, &t.bucket_distinct_cap, bp.nbuckets * sizeof(unsigned int)));
            NKX(ensure((void**)&t.bucket_start, &t.bucket_start_cap, (bp.nbuckets + 1) * sizeof(unsigned long long)));
            NKX(ensure((void**)&t.keys, &t.keys_cap, (n_in ? n_in : 1) * sizeof(unsigned long long)));
            NKX(ensure((void**)&t.counts, &t.counts_cap, (n_in ? n_in : 1) * sizeof(unsigned int)));
            NKX(ensure((void**)&t.kidx, &t.kidx_cap, (n_in ? n_in : 1) * sizeof(unsigned int)));
            NKX(cudaMemsetAsync(t.bucket_count, 0, bp.nbuckets * sizeof(unsigned int), s));
            NKX(cudaMemsetAsync(t.cursor + 1, 0, 2 * sizeof(unsigned long long), s));  // [1] distinct total, [2] overflow
            bucket_hist_kernel<<<grid_for(n_in), XT, 0, s>>>(in_words, in_idx, n_in, bp, t.bucket_count);
            bucket_scan_kernel<<<1, 1024, 0, s>>>(t.bucket_count, bp.nbuckets, t.bucket_start, t.bucket_fill);
            bucket_scatter_kernel<<<grid_for(n_in), XT, 0, s>>>(in_words, in_idx, in_weight, n_in, bp, t.bucket_start, t.bucket_fill,
                                                                t.keys, t.kidx, t.counts);
            constexpr int kDedupSmem = TABLE_SLOTS * (8 + 4 + 4);
            static std::atomic<bool> smem_set[64];
            int dev = 0;
            NKX(cudaGetDevice(&dev));
            if (dev >= 0 && dev < 64 && !smem_set[dev].load()) {
                NKX(cudaFuncSetAttribute(bucket_dedup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDedupSmem));
                smem_set[dev].store(true);
            }
            bucket_dedup_kernel<<<(unsigned)bp.nbuckets, XT, kDedupSmem, s>>>(t.keys, t.kidx, t.counts, t.bucket_start, t.bucket_distinct,
                                                                     merge ? nullptr : t.uniques, t.cursor + 1,
                                                                     reinterpret_cast<unsigned int*>(t.cursor + 2));
            err = cudaGetLastError();
            if (err != cudaSuccess) break;
            unsigned long long res[2] = {0, 0};
            NKX(cudaMemcpyAsync(res, t.cursor + 1, sizeof res, cudaMemcpyDeviceToHost, s));
            NKX(cudaStreamSynchronize(s));
            if ((res[1] & 0xFFFFFFFFull) == 0) {
                t.plan = bp;
                t.n_keys = res[0];
                err = cudaSuccess;
                break;
            }
            // a bucket held more distinct words than the table takes: partition finer.  `uniques` was touched by the
            // buckets that did finish: start over with it
            if (!merge) NKX(cudaMemsetAsync(t.uniques, 0, pool * sizeof(unsigned int), s));
            err = cudaErrorInvalidValue;
        }
        cudaFree(cat_w); cudaFree(cat_i); cudaFree(cat_c);
        if (err != cudaSuccess) return err;
    }
    t.valid = true;
    // the words of this call are consumed
    t.words_bound = 0;
    if (t.cursor) NKX(cudaMemsetAsync(t.cursor, 0, sizeof(unsigned long long), s));
    (void)fm;
    return cudaGetLastError();
}

// the table as dense arrays, bucket order (any of the outputs may be null); needs n_keys entries each
cudaError_t exact_dense_copy(ExactTable& t, unsigned long long* out_keys, unsigned int* out_counts, unsigned int* out_idx,
                             cudaStream_t s) {
    if (t.n_keys == 0) return cudaSuccess;
    const unsigned long long nb = t.plan.nbuckets;
    NKX(ensure((void**)&t.dense_start, &t.dense_start_cap, (nb + 1) * sizeof(unsigned long long)));
    // exclusive scan of the distinct counts (the fill cursors are free now: the scan kernel zeroes them, harmless)
    bucket_scan_kernel<<<1, 1024, 0, s>>>(t.bucket_distinct, nb, t.dense_start, t.bucket_fill);
    compact_table_kernel<<<(unsigned)nb, 128, 0, s>>>(t.keys, t.counts, t.kidx, t.bucket_start, t.bucket_distinct, t.dense_start,
                                                      out_keys, out_counts, out_idx);
    return cudaGetLastError();
}

cudaError_t exact_lookup(const ExactTable& t, const FastMod& fm, unsigned long long key, unsigned long long* d_out2, cudaStream_t s) {
    if (fm.is_pow2) lookup_kernel<true><<<1, 32, 0, s>>>(t.keys, t.counts, t.bucket_start, t.bucket_distinct, t.plan, fm, make_rotmul(), key, d_out2);
    else lookup_kernel<false><<<1, 32, 0, s>>>(t.keys, t.counts, t.bucket_start, t.bucket_distinct, t.plan, fm, make_rotmul(), key, d_out2);
    return cudaGetLastError();
}

cudaError_t exact_gather_uniques(const ExactTable& t, const unsigned long long* idx, unsigned long long n,
                                 unsigned int* out, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    gather_uniques_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(idx, n, t.uniques, out);
    return cudaGetLastError();
}

void exact_free(ExactTable& t) {
    cudaFree(t.words); cudaFree(t.widx); cudaFree(t.cursor);
    cudaFree(t.keys); cudaFree(t.counts); cudaFree(t.kidx); cudaFree(t.uniques); cudaFree(t.flags);
    cudaFree(t.bucket_count); cudaFree(t.bucket_fill); cudaFree(t.bucket_distinct); cudaFree(t.bucket_start); cudaFree(t.dense_start);
    t = ExactTable{};
}

}  // namespace nk
