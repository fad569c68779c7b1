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
// The reference's table is a concurrent hash map keyed by the k-mer word; so is this one: an open-addressing
// table in HBM, 16-byte slots {word u64, count u32, neuron index u32}, linear probing from mix(word) * nslots >> 64,
// slots claimed with a 64-bit atomicCAS on the word (a plain load first: a word that is already there costs one
// atomicAdd on its count).  The count kernel appends (word, neuron index) of every window (mode 2 / 4 — it has
// computed both anyway); at the end of the call one kernel inserts the appended records.  The first insertion of a
// word adds one to its neuron's `uniques` (kmer_per_neuron).  No sort, no partition, no 2^31 limit (64-bit
// cursors).  Round 2 first tried a bucket partition by neuron index + a shared-memory hash table per bucket: the
// scatter of 113 M 16-byte records into 74 K buckets alone took 7.0 ms (L2 transaction bound: an atomic, a load and
// three scattered stores per record; profiles/r02_exact.md) against 12.9 ms for round 1's whole CUB pipeline.
// Memory: 12 B per window appended + 16 B per table slot (1.5 slots per record).  Counts wrap at 2^32 like the
// reference's AtomicU32::fetch_add.
#include <algorithm>

#include "nk_kernels.cuh"

namespace nk {

namespace {

constexpr int XT = 256;                       // threads per block
constexpr unsigned long long EMPTY = ~0ull;   // slot marker.  Only pack_kmer of 32 T's (k = 32, non-canonical) equals it:
                                              // that one word is counted in a side slot (ExactTable::cursor[3..])

#define NKX(expr)                        \
    do {                                 \
        cudaError_t e_ = (expr);         \
        if (e_ != cudaSuccess) return e_; \
    } while (0)

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {  // splitmix64 finaliser
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__device__ __forceinline__ unsigned long long home_slot(unsigned long long word, unsigned long long nslots) {
    return __umul64hi(mix64(word), nslots);
}

__global__ void clear_slots_kernel(ExactSlot* __restrict__ slots, unsigned long long nslots) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const uint4 e = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);  // {EMPTY, count 0, idx 0}
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < nslots; i += stride)
        reinterpret_cast<uint4*>(slots)[i] = e;
}

// one record -> the table.  Returns true if the word was new.
__device__ __forceinline__ bool table_insert(ExactSlot* slots, unsigned long long nslots, unsigned long long word, unsigned idx,
                                             unsigned weight, unsigned long long* side /* [0] count [1] idx+1 of the all-ones word */) {
    if (word == EMPTY) {
        const unsigned long long before = atomicAdd(side, (unsigned long long)weight);
        if (before == 0ull) { side[1] = (unsigned long long)idx + 1ull; return true; }
        return false;
    }
    unsigned long long s = home_slot(word, nslots);
    for (;;) {
        unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(&slots[s].key);
        if (k == EMPTY) {
            k = atomicCAS(&slots[s].key, EMPTY, word);
            if (k == EMPTY) {
                slots[s].idx = idx;
                atomicAdd(&slots[s].count, weight);
                return true;
            }
        }
        if (k == word) {
            atomicAdd(&slots[s].count, weight);
            return false;
        }
        if (++s == nslots) s = 0;
    }
}

// records = (words[i], widx[i], weight 1).  uniques != null: +1 on the neuron of every new word.
__global__ void __launch_bounds__(XT) insert_words_kernel(const unsigned long long* __restrict__ words, const unsigned int* __restrict__ widx,
                                                           unsigned long long n, ExactSlot* slots, unsigned long long nslots,
                                                           unsigned int* uniques, unsigned long long* n_keys, unsigned long long* side) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned mine = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned ix = widx[i];
        if (table_insert(slots, nslots, words[i], ix, 1u, side)) {
            ++mine;
            if (uniques) atomicAdd(uniques + ix, 1u);
        }
    }
    mine = __reduce_add_sync(0xFFFFFFFFu, mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_keys, (unsigned long long)mine);
}

// re-insert the records of an older (smaller) table: table growth
__global__ void __launch_bounds__(XT) reinsert_kernel(const ExactSlot* __restrict__ old_slots, unsigned long long old_n,
                                                       ExactSlot* slots, unsigned long long nslots, unsigned long long* side) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < old_n; i += stride) {
        const ExactSlot o = old_slots[i];
        if (o.key != EMPTY) table_insert(slots, nslots, o.key, o.idx, o.count, side);
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

__global__ void lookup_kernel(const ExactSlot* __restrict__ slots, unsigned long long nslots, const unsigned long long* __restrict__ side,
                              unsigned long long key, unsigned long long* out) {
    out[0] = 0ull;
    out[1] = 0ull;
    if (key == EMPTY) {
        if (side[0]) { out[0] = 1ull; out[1] = side[0] & 0xFFFFFFFFull; }
        return;
    }
    if (nslots == 0) return;
    unsigned long long s = home_slot(key, nslots);
    for (;;) {
        const ExactSlot e = slots[s];
        if (e.key == EMPTY) return;
        if (e.key == key) { out[0] = 1ull; out[1] = e.count; return; }
        if (++s == nslots) s = 0;
    }
}

// occupied slots -> dense arrays (block-wise compaction; the order is the table's, i.e. none)
__global__ void __launch_bounds__(XT) compact_kernel(const ExactSlot* __restrict__ slots, unsigned long long nslots,
                                                      const unsigned long long* __restrict__ side, unsigned long long* cursor,
                                                      unsigned long long* __restrict__ out_keys, unsigned int* __restrict__ out_counts) {
    __shared__ unsigned int s_warp[XT / 32];
    __shared__ unsigned long long s_base;
    const unsigned tid = threadIdx.x;
    const unsigned long long chunks = (nslots + XT - 1) / XT;
    for (unsigned long long c = blockIdx.x; c < chunks; c += gridDim.x) {
        const unsigned long long i = c * XT + tid;
        ExactSlot e{EMPTY, 0u, 0u};
        if (i < nslots) e = slots[i];
        const bool occ = e.key != EMPTY;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, occ);
        if ((tid & 31) == 0) s_warp[tid >> 5] = __popc(bal);
        __syncthreads();
        if (tid == 0) {
            unsigned t = 0;
            for (int w = 0; w < XT / 32; ++w) t += s_warp[w];
            s_base = t ? atomicAdd(cursor, (unsigned long long)t) : 0ull;
        }
        __syncthreads();
        if (occ) {
            unsigned long long pos = s_base + __popc(bal & ((1u << (tid & 31)) - 1u));
            for (unsigned w = 0; w < (tid >> 5); ++w) pos += s_warp[w];
            if (out_keys) out_keys[pos] = e.key;
            if (out_counts) out_counts[pos] = e.count;
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && tid == 0 && side[0]) {  // the all-ones word lives beside the table
        const unsigned long long pos = atomicAdd(cursor, 1ull);
        if (out_keys) out_keys[pos] = EMPTY;
        if (out_counts) out_counts[pos] = (unsigned int)side[0];
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

// a fresh table of at least `want` slots; old != null: its records move over
cudaError_t new_table(ExactTable& t, unsigned long long want, bool keep_old, cudaStream_t s) {
    ExactSlot* old = t.slots;
    const unsigned long long old_n = t.nslots;
    if (!keep_old && want <= t.slots_cap) {
        t.nslots = want;
        clear_slots_kernel<<<grid_for(want), XT, 0, s>>>(t.slots, want);
        return cudaGetLastError();
    }
    ExactSlot* ns = nullptr;
    const unsigned long long cap = want + want / 8;
    NKX(cudaMalloc(&ns, cap * sizeof(ExactSlot)));
    clear_slots_kernel<<<grid_for(want), XT, 0, s>>>(ns, want);
    if (keep_old && old && old_n) {
        // the side slot of the all-ones word stays where it is (cursor[3..4]); only the table proper moves
        unsigned long long* scratch_side = t.cursor + 5;  // reinserted records never hit the side slot
        reinsert_kernel<<<grid_for(old_n), XT, 0, s>>>(old, old_n, ns, want, scratch_side);
    }
    NKX(cudaGetLastError());
    if (old) {
        NKX(cudaStreamSynchronize(s));
        cudaFree(old);
    }
    t.slots = ns;
    t.slots_cap = cap;
    t.nslots = want;
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
    if (t.cursor) NKX(cudaMemsetAsync(t.cursor, 0, sizeof(unsigned long long), s));
    if (tables_too) {
        t.n_keys = 0;
        t.nslots = 0;   // (the allocation is kept: the next build clears what it uses)
        t.valid = false;
        if (t.cursor) NKX(cudaMemsetAsync(t.cursor + 1, 0, 7 * sizeof(unsigned long long), s));
        if (t.uniques) NKX(cudaMemsetAsync(t.uniques, 0, pool * sizeof(unsigned int), s));
    }
    return cudaSuccess;
}

// The (word, index) records appended since the last call -> the table.  merge: ADD them to the table that exists
// (process_sequence: counts accumulate over calls, :218-221) and use its "neurons touched by this sequence" rule
// for the per-neuron column; else the table is replaced (counts.clear(), :157 / :426).
cudaError_t exact_finalize(ExactTable& t, const FastMod& fm, unsigned long long pool, unsigned /*key_bits*/, bool merge,
                           cudaStream_t s) {
    (void)fm;
    NKX(ensure_cursor(t, s));
    unsigned long long n_new = 0;
    NKX(cudaMemcpyAsync(&n_new, t.cursor, sizeof n_new, cudaMemcpyDeviceToHost, s));
    NKX(cudaStreamSynchronize(s));
    if (n_new > t.words_cap) n_new = t.words_cap;  // (the uniques pass lets its cursor run past the capacity)
    if (!t.uniques) {
        NKX(cudaMalloc(&t.uniques, pool * sizeof(unsigned int)));
        NKX(cudaMemsetAsync(t.uniques, 0, pool * sizeof(unsigned int), s));
    }
    if (!merge) {
        t.n_keys = 0;
        t.nslots = 0;
        NKX(cudaMemsetAsync(t.cursor + 1, 0, 7 * sizeof(unsigned long long), s));  // [1] distinct words, [3..4] all-ones word
        NKX(cudaMemsetAsync(t.uniques, 0, pool * sizeof(unsigned int), s));
    }
    if (n_new > 0) {
        if (merge) {
            if (!t.flags) {
                NKX(cudaMalloc(&t.flags, pool * sizeof(unsigned int)));
                NKX(cudaMemsetAsync(t.flags, 0, pool * sizeof(unsigned int), s));
            }
            const unsigned blocks = (unsigned)((n_new + 255) / 256);
            touched_kernel<<<blocks, 256, 0, s>>>(t.widx, n_new, t.flags, t.uniques, 0);
            touched_kernel<<<blocks, 256, 0, s>>>(t.widx, n_new, t.flags, t.uniques, 1);
        }
        // at most n_keys + n_new distinct words afterwards: keep the load factor under 2/3
        const unsigned long long need = (t.n_keys + n_new) * 3 / 2 + 64;
        if (need > t.nslots) NKX(new_table(t, merge && t.nslots ? std::max(need, t.nslots * 2) : need, merge && t.n_keys > 0, s));
        insert_words_kernel<<<grid_for(n_new), XT, 0, s>>>(t.words, t.widx, n_new, t.slots, t.nslots, merge ? nullptr : t.uniques,
                                                           t.cursor + 1, t.cursor + 3);
        NKX(cudaGetLastError());
        unsigned long long nk = 0;
        NKX(cudaMemcpyAsync(&nk, t.cursor + 1, sizeof nk, cudaMemcpyDeviceToHost, s));
        NKX(cudaStreamSynchronize(s));
        t.n_keys = nk;
    }
    t.valid = true;
    // the words of this call are consumed
    t.words_bound = 0;
    NKX(cudaMemsetAsync(t.cursor, 0, sizeof(unsigned long long), s));
    return cudaGetLastError();
}

// the table as dense device arrays (either output may be null; n_keys entries each); the order is the table's
cudaError_t exact_dense_copy(ExactTable& t, unsigned long long* out_keys, unsigned int* out_counts, cudaStream_t s) {
    if (t.n_keys == 0) return cudaSuccess;
    NKX(cudaMemsetAsync(t.cursor + 2, 0, sizeof(unsigned long long), s));
    compact_kernel<<<grid_for(t.nslots), XT, 0, s>>>(t.slots, t.nslots, t.cursor + 3, t.cursor + 2, out_keys, out_counts);
    return cudaGetLastError();
}

cudaError_t exact_lookup(const ExactTable& t, unsigned long long key, unsigned long long* d_out2, cudaStream_t s) {
    lookup_kernel<<<1, 1, 0, s>>>(t.slots, t.nslots, t.cursor + 3, key, d_out2);
    return cudaGetLastError();
}

cudaError_t exact_gather_uniques(const ExactTable& t, const unsigned long long* idx, unsigned long long n,
                                 unsigned int* out, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    gather_uniques_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(idx, n, t.uniques, out);
    return cudaGetLastError();
}

void exact_free(ExactTable& t) {
    cudaFree(t.words); cudaFree(t.widx); cudaFree(t.cursor); cudaFree(t.slots); cudaFree(t.uniques); cudaFree(t.flags);
    t = ExactTable{};
}

}  // namespace nk
